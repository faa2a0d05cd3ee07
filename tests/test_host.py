"""CPU tests (no GPU): C-ABI library loads and exports every declared symbol; host-side logic."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
import shxarray_b200 as sb
from shxarray_b200 import capi


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "shb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(shb200_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTS)
    L = ctypes.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert capi.lib().shb200_version() == 100


def test_plan_desc_struct_matches_header_size():
    # 4 int32, int64, 2 ptr, 2 int32, ptr, 2 int32, ptr, double, ptr, 2 int32 -> 96 bytes on LP64
    assert ctypes.sizeof(capi.PlanDesc) == 96


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.SHB200Error):
        capi.Plan(10, [0.0, 1.0], [0.0, 1.0])


def test_invalid_arguments_raise():
    with pytest.raises(ValueError):
        capi.Plan(-1, [0.0], [0.0])
    with pytest.raises(ValueError):
        sb.lonlat_grid(nmax=10, gtype="nonsense")
    with pytest.raises(ValueError):
        sb.lonlat_grid(nmax=10, lon=[0], lat=[0], gtype="regular")
    with pytest.raises(ValueError):
        sb.lonlat_grid()
    with pytest.raises(RuntimeError):   # analysis.pyx:39-40
        sb.analysis_weight_shlib([0.0, 1.0, 3.0], [0.0, 1.0, 2.0])


def test_nm_index_layout_bit_exact():
    """sh_indexing.py:66: [(n,m) for n in range(nmin,nmax+1) for m in range(-n,n+1)]; position of (n,m) is n^2+n+m."""
    for nmax, nmin in ((0, 0), (5, 0), (12, 3)):
        n, m = sb.nm_index(nmax, nmin)
        ref = [(a, b) for a in range(nmin, nmax + 1) for b in range(-a, a + 1)]
        assert [(int(a), int(b)) for a, b in zip(n, m)] == ref
    n, m = sb.nm_index(30)
    assert np.array_equal(np.arange(len(n)), n.astype(np.int64) ** 2 + n + m)


def test_shlib_grid_reproduced():
    """shlib.pyx:87-96."""
    for nmax, res in ((60, 1.0), (96, 0.5), (256, 0.25), (720, 0.125)):
        g = sb.lonlat_grid(nmax, gtype="regular")
        assert g["lon"][0] == -180 + res / 2 and g["lat"][0] == -90 + res / 2
        assert len(g["lon"]) == round(360 / res) and len(g["lat"]) == round(180 / res)
        g0 = sb.lonlat_grid(nmax, gtype="regular_lon0")
        assert g0["lon"][0] == 0.0 and g0["lon"][-1] == 360 - res


def test_gauss_grid():
    g = sb.lonlat_grid(720, gtype="gauss")
    assert len(g["lat"]) == 724 and len(g["lon"]) == 1536
    assert np.array_equal(g["lat"], -g["lat"][::-1])
    x, w = capi.gauss_nodes(64)
    x2, w2 = np.polynomial.legendre.leggauss(64)
    assert np.abs(x - x2).max() < 1e-15 and np.abs(w - w2).max() < 1e-14
    # exactness: integrates x^(2k) up to degree 2n-1
    for k in (0, 2, 10, 126):
        assert abs(np.sum(w * x ** k) - 2.0 / (k + 1)) < 1e-14


def test_analysis_weight_expression():
    lon = np.arange(0.25, 360, 0.5)
    lat = np.arange(-89.75, 90, 0.5)
    w = sb.analysis_weight_shlib(lon, lat)
    stepx = 0.5 * np.pi / 180
    assert w == abs(stepx * stepx) / (4 * np.pi)


def test_c_abi_from_plain_c(tmp_path):
    """include/shb200.h is valid C and the library is callable from a C program without python/torch."""
    import shutil
    import subprocess
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if cc is None:
        pytest.skip("no C compiler")
    capi.lib()
    exe = str(tmp_path / "cabi_smoke")
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cabi_smoke.c"),
                           "-o", exe, "-L", libdir, "-lshb200", "-lm", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert out.stdout.startswith(("no-gpu", "gpu: ok"))
