"""Formats either side of the transform: the oracle restatement of the reference's readers against the
reference's own test files and known answers (tests/test_shformats.py:15-46 of the reference), then the device path
(shb200_gfc_parse, shb200_polygon_mask, shb200_scatter_nm) against that oracle -- bit-exact."""
import gzip
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import formats_oracle as FO

FMT = os.path.join(GOLDEN, "formats")


def _lines(path):
    return iter(gzip.open(path, "rb") if path.endswith(".gz") else open(path, "rb"))


def test_oracle_reproduces_the_reference_known_answers():
    """tests/test_shformats.py:15-46 of the reference: nmax, number of records, epoch, selected values."""
    it = _lines(os.path.join(FMT, "icgem_test_sig_ITSG.gfc.gz"))
    attr, time = FO.icgem_header(it)
    nm, c, s = FO.read_records(it, attr["nmax"])
    assert attr["nmax"] == 60 and len(nm) == 599
    assert np.datetime64(time[0]) == np.datetime64("2004-10-15T12:00:00")
    assert dict(zip(nm, c))[(60, 46)] == 1.665788238155e-09
    assert dict(zip(nm, s))[(60, -46)] == 2.363195963172e-12
    it = _lines(os.path.join(FMT, "icgem_test_nosig_ITSG.gfc"))
    attr, time = FO.icgem_header(it)
    nm, c, s = FO.read_records(it, attr["nmax"])
    assert attr["nmax"] == 2 and len(nm) == 9 and s is None
    assert np.datetime64(time[0]) == np.datetime64("2013-11-15T12:00:00")
    d = dict(zip(nm, c))
    assert d[(2, 0)] == -4.841696152100e-04 and d[(1, 1)] == 0.0


def test_oracle_polygon_masks_analytic():
    lon = np.arange(-179.5, 180, 1.0)
    lat = np.arange(-89.5, 90, 1.0)
    box = [np.array([[10.0, -20.0], [50.0, -20.0], [50.0, 30.0], [10.0, 30.0]])]
    hole = [box[0], np.array([[20.0, 0.0], [30.0, 0.0], [30.0, 10.0], [20.0, 10.0]])]
    m = FO.polygon_masks([box, hole], lon, lat)
    X, Y = np.meshgrid(lon, lat)
    inbox = (X > 10) & (X < 50) & (Y > -20) & (Y < 30)
    inhole = (X > 20) & (X < 30) & (Y > 0) & (Y < 10)
    assert np.array_equal(m[0], inbox.astype(float))
    assert np.array_equal(m[1], (inbox & ~inhole).astype(float))
    # a 0..360 grid is wrapped to [-180, 180) before the test (polygons.py:76-81)
    west = [np.array([[-60.0, -10.0], [-20.0, -10.0], [-20.0, 10.0], [-60.0, 10.0]])]
    m2 = FO.polygon_masks([west], np.arange(0.5, 360, 1.0), lat)
    X2, Y2 = np.meshgrid(np.arange(0.5, 360, 1.0), lat)
    assert np.array_equal(m2[0], ((X2 > 300) & (X2 < 340) & (Y2 > -10) & (Y2 < 10)).astype(float))


# ---- device path ---------------------------------------------------------------------------------------------------------
def _bits(a):
    return np.asarray(a, dtype=np.float64).view(np.uint64)


def _synthetic_gfc(nmax, seed, sigma=True, d_exp=False, extras=True):
    """An ICGEM file with the things real files contain: comment lines, a header, gfc and gfct records, D exponents,
    varying widths and digit counts, a record above max_degree, trailing whitespace, no final newline."""
    rng = np.random.default_rng(seed)
    out = ["some free text before the header", "begin_of_head ====", "modelname   synth_2010-03", "max_degree  %d" % nmax,
           "earth_gravity_constant 3.986004415E+14", "radius 6378136.3", "norm fully_normalized", "key L M C S sigma_C sigma_S", "end_of_head ===="]
    fmts = ["%.12e", "%.15E", "%.17e", "%.6e", "%.18e", "%r"]
    for n in range(nmax + 1):
        for m in range(n + 1):
            v = rng.standard_normal(4) * 10.0 ** rng.integers(-25, 3, 4)
            f = fmts[int(rng.integers(len(fmts)))]
            tok = [(f % x) if f != "%r" else repr(float(x)) for x in v]
            if d_exp:
                tok = [t.replace("e", "D").replace("E", "D") for t in tok]
            key = "gfct" if extras and (n + m) % 17 == 5 else "gfc"
            cols = tok if sigma else tok[:2]
            out.append(("%s %4d %4d " % (key, n, m)) + "  ".join(cols) + ("   " if m % 3 == 0 else "") + (" 20100101" if key == "gfct" else ""))
            if extras and key == "gfct":
                out.append("trnd %4d %4d 1.0e-12 1.0e-12 1e-13 1e-13" % (n, m))
    if extras:
        out.append("gfc %d 0 1.0 0.0 0.0 0.0" % (nmax + 1))       # above max_degree: skipped
        out.append("# a comment mentioning gfc inside the line")
    return ("\n".join(out)).encode()


@pytest.mark.gpu
@pytest.mark.parametrize("nmax,sigma,d_exp", [(60, True, False), (95, False, True), (200, True, True)])
def test_gfc_records_parsed_on_device_bit_exact(nmax, sigma, d_exp):
    from shxarray_b200 import formats as F
    data = _synthetic_gfc(nmax, 11 + nmax, sigma, d_exp)
    it = iter(data.split(b"\n"))
    attr, _ = FO.icgem_header(it)
    # the synthetic file carries one record above max_degree: the reference would run off its arrays there (icgem.py:115), the
    # device reader skips and counts it -- the oracle is asked for the same with nmaxstop = max_degree
    nm, c, s = FO.read_records(it, attr["nmax"], nmaxstop=attr["nmax"])
    r = F.read_icgem(data)
    assert r["nmax"] == nmax and r["stats"]["skipped_above_nmax"] == 1
    assert r["stats"]["records"] == (nmax + 1) * (nmax + 2) // 2
    assert np.array_equal(_bits(r["cnm"].cpu().numpy()), _bits(FO.dense(nm, c, nmax)))
    if sigma:
        assert np.array_equal(_bits(r["sigcnm"].cpu().numpy()), _bits(FO.dense(nm, s, nmax)))
    else:
        assert r["sigcnm"] is None
    assert sorted(zip(r["n"].tolist(), r["m"].tolist())) == sorted(nm)
    # nmaxstop: records above it are skipped (icgem.py:105-111)
    r2 = F.read_icgem(data, nmaxstop=20)
    it = iter(data.split(b"\n"))
    FO.icgem_header(it)
    nm2, c2, _ = FO.read_records(it, 20, nmaxstop=20)
    assert r2["nmax"] == 20 and np.array_equal(_bits(r2["cnm"].cpu().numpy()), _bits(FO.dense(nm2, c2, 20)))


@pytest.mark.gpu
def test_reference_icgem_files_on_device():
    """the reference's own test files and known answers (tests/test_shformats.py:15-46) through the device reader"""
    from shxarray_b200 import formats as F
    r = F.read_icgem(os.path.join(FMT, "icgem_test_sig_ITSG.gfc.gz"))
    assert r["nmax"] == 60 and len(r["n"]) == 599 and np.datetime64(r["time"][0]) == np.datetime64("2004-10-15T12:00:00")
    assert r["cnm"][60 * 60 + 60 + 46].item() == 1.665788238155e-09
    assert r["sigcnm"][60 * 60 + 60 - 46].item() == 2.363195963172e-12
    r = F.read_icgem(os.path.join(FMT, "icgem_test_nosig_ITSG.gfc"))
    assert r["nmax"] == 2 and len(r["n"]) == 9 and r["sigcnm"] is None
    assert r["cnm"][2 * 2 + 2].item() == -4.841696152100e-04 and r["cnm"][1 + 1 + 1].item() == 0.0
    # a batch of files of different max_degree
    coef, infos = F.read_icgem_batch([os.path.join(FMT, "icgem_test_sig_ITSG.gfc.gz"), os.path.join(FMT, "icgem_test_nosig_ITSG.gfc")], 60)
    assert coef.shape == (2, 61 * 61) and coef[0, 60 * 60 + 60 + 46].item() == 1.665788238155e-09
    assert coef[1, 6].item() == -4.841696152100e-04 and float(coef[1, 9:].abs().max()) == 0.0


@pytest.mark.gpu
def test_decimal_conversion_matches_float_bit_for_bit():
    """shb200_parse_floats against Python's float() (correctly rounded) on 400 k literals: random digit strings over the whole
    exponent range, halfway cases, subnormals, overflow; tokens outside the grammar must be handed back, never guessed."""
    import torch
    from shxarray_b200 import capi
    rng = np.random.default_rng(3)
    toks = []
    for i in range(400000):
        nd = int(rng.integers(1, 20))
        digs = "".join(rng.choice(list("0123456789"), nd))
        k = int(rng.integers(0, nd + 1))
        mant = (digs[:k] or "0") + "." + digs[k:] if i % 2 else digs
        ex = int(rng.integers(-345, 312)) if i % 3 else int(rng.integers(-30, 30))
        sign = "-" if i % 5 == 0 else ("+" if i % 7 == 0 else "")
        echar = "eED"[i % 3]
        toks.append(f"{sign}{mant}{echar}{ex:+d}" if i % 11 else f"{sign}{mant}")
    # halfway cases between adjacent doubles (exact decimal expansions are long -> must be handed back or exact), powers of two
    toks += ["9007199254740993", "9007199254740992.5", "4.9406564584124654e-324", "2.2250738585072014e-308", "1.7976931348623157e308",
             "1.7976931348623159e308", "2.4703282292062327e-324", "2.4703282292062328e-324", "0.0", "-0.0", "0e0", ".5", "5.", "1e0"]
    bad = ["1.0d0", "abc", "1e", "--1", "1_000", "inf", "nan", "1.2.3", "", "+", "1e+"]
    alltok = toks + [b for b in bad if b]
    text = "\n".join(alltok).encode()
    start = np.cumsum([0] + [len(t) + 1 for t in alltok[:-1]]).astype(np.int64)
    stop = start + np.array([len(t) for t in alltok], dtype=np.int64)
    tt = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    ts, te = torch.from_numpy(start).cuda(), torch.from_numpy(stop).cuda()
    out = torch.empty(len(alltok), dtype=torch.float64, device="cuda")
    ok = torch.empty(len(alltok), dtype=torch.uint8, device="cuda")
    capi.check(capi.lib().shb200_parse_floats(tt.data_ptr(), ts.data_ptr(), te.data_ptr(), len(alltok), out.data_ptr(), ok.data_ptr(), 0))
    torch.cuda.synchronize()
    ok = ok.cpu().numpy().astype(bool)
    out = out.cpu().numpy()
    ref = np.array([float(t.replace("D", "E")) for t in toks])
    nt = len(toks)
    assert ok[:nt].mean() > 0.99                                   # the hand-back path stays the exception
    assert np.array_equal(_bits(out[:nt][ok[:nt]]), _bits(ref[ok[:nt]]))
    assert not ok[nt:].any()                                       # nothing outside the grammar is ever converted here


@pytest.mark.gpu
def test_polygon_masks_and_polygon2sh_vs_oracle():
    import torch
    import shxarray_b200 as sb
    from shxarray_b200 import formats as F
    from oracle import oracle as O
    rng = np.random.default_rng(4)
    nmax = 40
    star = np.stack([35 * (1 + 0.5 * np.sin(5 * np.linspace(0, 2 * np.pi, 400, endpoint=False))) * np.cos(np.linspace(0, 2 * np.pi, 400, endpoint=False)) + 7.123,
                     25 * (1 + 0.5 * np.sin(5 * np.linspace(0, 2 * np.pi, 400, endpoint=False))) * np.sin(np.linspace(0, 2 * np.pi, 400, endpoint=False)) - 3.217], axis=1)
    tri = np.array([[-150.013, 40.07], [-100.31, 75.11], [-60.77, 35.23]])
    hole = np.array([[-110.3, 50.2], [-95.1, 50.4], [-100.7, 60.9]])
    blob = np.cumsum(rng.standard_normal((1500, 2)), axis=0)
    blob = blob / np.abs(blob).max() * np.array([60.0, 40.0]) + np.array([100.03, -30.01])      # self-intersecting: even-odd rule
    polys = [[star], [tri, hole], [blob], [tri, hole, star]]
    eng = sb.SHEngine(device=0)
    for gtype in ("regular", "regular_lon0"):
        g = sb.lonlat_grid(nmax, gtype=gtype)
        got = F.polygon_masks(polys, g["lon"], g["lat"]).cpu().numpy()
        ref = FO.polygon_masks(polys, g["lon"], g["lat"])
        assert np.array_equal(got, ref)
        assert 0.01 < got.mean() < 0.5
    coef, g = F.polygon2sh(eng, polys, nmax, gtype="regular")
    ref = O.analysis(FO.polygon_masks(polys, g["lon"], g["lat"]), nmax, g["lon"], g["lat"], kind=O.best_kind(), nthreads=O.max_threads())
    c = coef.cpu().numpy()
    assert np.max(np.abs(c - ref).max(axis=1) / np.abs(ref).max(axis=1)) <= 1e-12
    # scatter of (n, m, value) records
    n, m = sb.nm_index(nmax)
    perm = rng.permutation(len(n))[: len(n) - 50]
    vals = rng.standard_normal((3, len(perm)))
    dense = F.scatter_nm(n[perm], m[perm], vals, nmax).cpu().numpy()
    refd = np.zeros((3, len(n)))
    refd[:, (n[perm].astype(np.int64) ** 2 + n[perm] + m[perm])] = vals
    assert np.array_equal(dense, refd)
