"""shxarray_b200 -- B200-native spherical-harmonic synthesis/analysis engine for shxarray.

Only the hot path is here: hand-written sm_100a CUDA kernels behind a C ABI
(``include/shb200.h``), a numpy/torch-level host API (``engine``) and the xarray plugin class
(``xr_backend.B200Backend``) that registers next to ``engine='shlib'`` / ``'shtns'``.
"""
__version__ = "0.1.0"

from . import capi  # noqa: F401
from .engine import (SHEngine, analysis, analysis_weight_shlib, lonlat_grid, nm_index, synthesis)  # noqa: F401
