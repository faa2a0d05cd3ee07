import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def field_norm_err(a, ref, axis=None):
    """max|a-ref| / max|ref| -- the parity metric of SURVEY.md §8d (per field when ``axis`` selects the field's axes)."""
    a = np.asarray(a)
    ref = np.asarray(ref)
    num = np.abs(a - ref).max(axis=axis)
    den = np.abs(ref).max(axis=axis)
    return np.max(num / den)


def load_group(npz, tag):
    return {k.split("__", 1)[1]: npz[k] for k in npz.files if k.startswith(tag + "__")}


@pytest.fixture(scope="session")
def engine():
    import shxarray_b200 as sb
    return sb.SHEngine(device=0)
