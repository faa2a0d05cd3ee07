"""TEST INFRASTRUCTURE ONLY: CPU oracle for the shlib synthesis/analysis path.

Nothing under ``shxarray_b200/`` may import this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs do.
"""
from .oracle import *  # noqa: F401,F403
