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
    assert capi.lib().shb200_version() == capi.VERSION == 200


def test_plan_desc_struct_matches_header_size():
    # 4 int32, int64, 2 ptr, 2 int32, ptr, 2 int32, ptr, double, ptr, 2 int32 -> 96 bytes on LP64
    assert ctypes.sizeof(capi.PlanDesc) == 96
    # shb200_call_opts: 4 int32 + pointer
    assert ctypes.sizeof(capi.CallOpts) == 24


def test_c_struct_sizes_match_the_header(tmp_path):
    """sizeof of every struct of include/shb200.h as the C compiler sees it == the ctypes mirror."""
    import shutil
    import subprocess
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if cc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "shb200.h"\nint main(void){printf("%zu %zu %zu\\n", sizeof(shb200_plan_desc), sizeof(shb200_plan_info), sizeof(shb200_call_opts));return 0;}\n')
    exe = str(tmp_path / "sz")
    subprocess.check_call([cc, "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe])
    sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    assert sizes == [ctypes.sizeof(capi.PlanDesc), ctypes.sizeof(capi.PlanInfo), ctypes.sizeof(capi.CallOpts)]


def test_per_call_options_validate():
    with pytest.raises(ValueError):
        capi._opts(degree_scale=np.ones(5), nmax=10)
    o, keep = capi._opts(degree_scale=np.ones(11), order_shard=(1, 4), nmax=10)
    assert o.struct_size == 24 and o.shard_rank == 1 and o.shard_world == 4 and o.degree_scale == keep.ctypes.data
    assert capi._opts() == (None, None)


def test_digest_distinguishes_large_arrays():
    """plan-cache key of large index arrays: permutations and single-element changes give different keys"""
    from shxarray_b200.engine import _digest
    n, m = sb.nm_index(300)                      # 90601 entries: the checksum branch
    base = _digest(n, m)
    assert base == _digest(n.copy(), m.copy())
    m2 = m.copy(); m2[5000], m2[5001] = m2[5001], m2[5000]
    assert _digest(n, m2) != base
    n3 = n.copy(); n3[70000] += 1
    assert _digest(n3, m) != base
    assert _digest(n[:-1], m[:-1]) != base


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


def _host_copy(a, to_dense=True, dense=None):
    """dense image of the (possibly strided) float64 array ``a`` through shb200_debug_host_copy (or the scatter back)"""
    ext = np.array(a.shape, dtype=np.int64)
    st = np.array([x // 8 for x in a.strides], dtype=np.int64)
    if dense is None:
        dense = np.full(a.size, np.nan)
    info = np.zeros(10, dtype=np.int64)
    capi.check(capi.lib().shb200_debug_host_copy(a.ndim, ext.ctypes.data, st.ctypes.data, a.ctypes.data, dense.ctypes.data, int(to_dense), info.ctypes.data))
    return dense, info


def test_host_layout_gather_scatter_only_touch_owned_elements():
    """The host entry points read / write ONLY the elements a call owns (a span copy overwrote the holes of
    batch-fastest arrays and strided out= views).  Checked here without a GPU through the library's own run logic."""
    rng = np.random.default_rng(3)
    base = rng.standard_normal((7, 12, 20))
    views = {
        "contiguous": base,
        "batch chunk of batch-first": base[2:5],
        "shlib layout chunk [lat, lon, B] -> (b, lat, lon)": np.moveaxis(base, 2, 0)[3:11],      # batch fastest, holes between runs
        "row and column slices": base[:, 1:9, 2:17],
        "every other longitude": base[:, :, ::2],
        "nm-first coefficients": base[0].T,
        "single field": base[3:4],
        "one row": base[:, 5:6, :],
    }
    for name, v in views.items():
        dense, info = _host_copy(v)
        # the dense image follows the HOST axis order (largest stride outermost): compare as sets of runs via dense strides
        ds = info[7:7 + v.ndim]
        ref = np.empty(v.size)
        idx = np.indices(v.shape).reshape(v.ndim, -1)
        ref[(idx * ds[:, None]).sum(axis=0)] = v.reshape(-1)
        assert np.array_equal(dense, ref), name
        assert info[2] == v.size and info[3] == 1, name
        # scatter back into a poisoned copy of the parent: owned elements restored, holes untouched
        parent = np.full(base.shape, -7.0)
        pv = np.lib.stride_tricks.as_strided(parent.reshape(-1)[(v.ctypes.data - base.ctypes.data) // 8:], shape=v.shape, strides=v.strides)
        _host_copy(pv, to_dense=False, dense=dense)
        assert np.array_equal(pv, v), name
        mask = np.zeros(base.shape, dtype=bool)
        mv = np.lib.stride_tricks.as_strided(mask.reshape(-1)[(v.ctypes.data - base.ctypes.data) // 8:], shape=v.shape, strides=tuple(x // 8 for x in v.strides))
        mv[...] = True
        assert np.all(parent[~mask] == -7.0), name
    # layout facts the pipeline relies on
    _, info = _host_copy(base)
    assert info[0] == 1 and info[1] == base.size and info[4] == 1          # one contiguous run, DMA-able
    _, info = _host_copy(np.moveaxis(base, 2, 0)[3:11])
    assert info[0] == 2 and info[1] == 8 and info[6] != 0                  # runs of 8 fields; batch is NOT the outer axis
    _, info = _host_copy(base[:, :, ::2])
    assert info[1] == 1 and info[4] == 0                                   # element gather: never handed to the DMA engine
    # broadcast (stride 0) inputs gather fine but are not a valid output
    b = np.broadcast_to(rng.standard_normal((1, 12, 20)), (4, 12, 20))
    dense, info = _host_copy(b)
    ds = info[7:10]
    idx = np.indices(b.shape).reshape(3, -1)
    assert info[3] == 0 and np.array_equal(dense[(idx * ds[:, None]).sum(axis=0)], b.reshape(-1))


def test_affinity_helpers_parse_and_degrade():
    """shxarray_b200.affinity: cpulist parsing, and binding is a described no-op where the device-local CPU list is unknown
    (this container has no GPU) -- it must never raise or shrink the affinity to nothing."""
    import os
    from shxarray_b200 import affinity
    assert affinity._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert affinity._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    info = affinity.bind_to_device_numa(0)
    assert isinstance(info, dict) and "bound" in info and "reason" in info
    assert os.sched_getaffinity(0) and os.sched_getaffinity(0) <= before
    os.sched_setaffinity(0, before)


def test_plan_cache_digest_memo_sees_in_place_changes():
    """engine._digest memoises small coordinate arrays per object but verifies the contents: an array changed in place, or a new
    array at a recycled id(), must give a new digest (a stale one would silently reuse a plan made for another grid)."""
    import numpy as np
    from shxarray_b200.engine import _digest
    lon = np.arange(0.0, 360.0, 1.0)
    lat = np.arange(-89.5, 90.0, 1.0)
    d0 = _digest(lat, lon, None)
    assert _digest(lat, lon, None) == d0 and _digest(lat.copy(), lon.copy(), None) == d0
    lon[17] += 1e-9
    d1 = _digest(lat, lon, None)
    assert d1 != d0
    lon[17] -= 1e-9
    assert _digest(lat, lon, None) in (d0, d1)                   # back to the old value bit for bit, or not: never a third state
    assert _digest(lat, lon.astype(np.float32), None) != d0       # dtype is part of the key
    big = np.arange(200000, dtype=np.int32)
    e0 = _digest(big)
    big[123456] += 1
    assert _digest(big) != e0


def test_host_pipeline_chunk_plan():
    """Chunk sizes of the host entry points (host_io.cu: host_chunk_plan, through shb200_debug_host_chunks; no GPU): whole
    8-field groups that cover the batch exactly, at most 8 chunks, one chunk when the batch axis is not the slowest, and the
    exposed end of the pipeline (first chunk of a synthesis, last of an analysis, both for a round trip) cut down to 8 fields."""
    def plan(batch, which, chunkable=True):
        sizes = np.zeros(8, dtype=np.int32)
        n = capi.lib().shb200_debug_host_chunks(batch, which, int(chunkable), sizes.ctypes.data)
        assert 1 <= n <= 8, (batch, which, n)
        return [int(v) for v in sizes[:n]]

    for batch in list(range(1, 140)) + [240, 256, 500, 1000]:
        for which in (0, 1, 2):
            s = plan(batch, which)
            assert sum(s) == batch and all(v > 0 for v in s)
            assert all(v % 8 == 0 for v in s[:-1]) or len(s) == 1            # every chunk but the last is whole 8-field groups
            if batch < 32:
                assert s == [batch]                                           # too few fields to overlap copies and kernels
            else:
                if which in (0, 2):
                    assert s[0] == 8                                          # head of the pipeline
                if which in (1, 2):
                    assert s[-1] == (batch % 8 or 8)                          # tail of the pipeline: the remainder, at most 8 fields
            assert plan(batch, which, chunkable=False) == [batch]
    assert plan(64, 0) == [8, 8, 16, 16, 16] and plan(64, 1) == [16, 16, 16, 8, 8] and plan(64, 2) == [8, 8, 16, 16, 8, 8]
    assert plan(128, 2) == [8, 8, 16, 32, 32, 16, 8, 8]
    sizes = np.zeros(8, dtype=np.int32)
    assert capi.lib().shb200_debug_host_chunks(0, 0, 1, sizes.ctypes.data) < 0
    assert capi.lib().shb200_debug_host_chunks(8, 3, 1, sizes.ctypes.data) < 0
