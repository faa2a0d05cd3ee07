"""CPU tests (no GPU): the oracle restatement against the reference's own known answers and against the
reference's headers compiled in place (oracle/_ref), plus the committed golden vectors."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, field_norm_err, load_group
from oracle import oracle as O
from shxarray_b200 import synthetic as S


def test_port_tables_bit_exact_vs_reference_headers():
    if not O.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for nmax in (1, 2, 20, 200, 720):
        for lat in (0.0, 0.25, 45.0, -31.0, 89.96, -89.5, 90.0):
            c = np.sin(lat * np.pi / 180.0)
            assert np.array_equal(O.legendre(nmax, c, "port"), O.legendre(nmax, c, "ref"))


@pytest.mark.parametrize("kind", ["port", "ref"])
def test_reference_known_answers_synthesis(kind):
    """tests/fixtures.py:66,69-75 through test_synthesis.py:45-50 (point mode), tol 1e-7 relative (:13)."""
    if kind == "ref" and not O.have_ref():
        pytest.skip("oracle/_ref not built")
    ka = json.load(open(os.path.join(GOLDEN, "ref_known_answers.json")))
    for nmax, key in ((200, "nmax200"), (20, "nmax20")):
        v = np.array(ka[key])
        n, m = O.nm_index(nmax)
        out = O.synthesis(np.ones((1, len(n))), n, m, v[:, 0], v[:, 1], grid=False, kind=kind)[:, 0]
        assert np.abs((out - v[:, 2]) / v[:, 2]).max() < 1e-7


def test_reference_known_answers_grid_layouts():
    """test_synthesis.py:24-29: grid lon=-180:60:180, lat=-60:60:60, unit coefficients x (basins,time) scales,
    both nm-last and nm-first memory layouts (fixtures.py:86-108)."""
    ka = json.load(open(os.path.join(GOLDEN, "ref_known_answers.json")))
    v = np.array(ka["nmax200"])
    n, m = O.nm_index(200)
    scales = np.outer(np.arange(5), 0.3 * np.arange(4)).reshape(-1)
    coef = np.ones((len(scales), len(n))) * scales[:, None]
    lon = np.arange(-180, 180.01, 60.0)
    lat = np.arange(-60, 60.01, 60.0)
    for nm_first in (False, True):
        c = np.ascontiguousarray(coef.T) if nm_first else coef
        g = O.synthesis(c, n, m, lon, lat, grid=True, nm_first=nm_first, kind="port")
        for lo, la, val in v:
            got = g[np.where(lat == la)[0][0], np.where(lon == lo)[0][0], :]
            ok = scales != 0
            assert np.abs((got[ok] - val * scales[ok]) / (val * scales[ok])).max() < 1e-7


def test_reference_known_answers_analysis():
    """tests/test_analysis.py:14-25: analysis of the paracap grid vs analytic coefficients, tol 1e-7 absolute."""
    z = np.load(os.path.join(GOLDEN, "paracap_n200.npz"))
    nmax = 100   # same grid, lower truncation keeps the CPU suite short; nmax=200 runs in the GPU suite
    n, m = O.nm_index(200)
    keep = n <= nmax
    out = O.analysis(z["z"], nmax, z["x"], z["y"], kind="port")
    assert np.abs(out - z["cap_coef"][:, keep]).max() < 1e-7


def test_port_matches_ref_transforms():
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    nmax, B = 40, 3
    n, m = O.nm_index(nmax)
    c = S.random_coefficients(nmax, B, kaula=False)
    lon, lat = O.shlib_grid(nmax)
    lon, lat = lon[::7], lat[::5]
    a = O.synthesis(c, n, m, lon, lat, kind="port")
    b = O.synthesis(c, n, m, lon, lat, kind="ref")
    assert field_norm_err(a, b) < 1e-13
    lon, lat = O.shlib_grid(24)
    f = np.random.default_rng(1).standard_normal((2, len(lat), len(lon)))
    a = O.analysis(f, 24, lon, lat, kind="port")
    b = O.analysis(f, 24, lon, lat, kind="ref")
    assert field_norm_err(a, b) < 1e-13


def test_port_reproduces_committed_goldens():
    """The committed vectors were generated with oracle/_ref; the port (which travels to the GPU box) must agree."""
    z = np.load(os.path.join(GOLDEN, "synthesis_shlib_grids.npz"))
    g = load_group(z, "n60_regular")
    nmax, B = int(g["nmax"]), int(g["B"])
    n, m = O.nm_index(nmax)
    c = np.concatenate([S.random_coefficients(nmax, B - B // 2, kaula=True), S.random_coefficients(nmax, B // 2, kaula=False, seed=S.SEED + 1)])
    lon, lat = O.shlib_grid(nmax, str(g["gtype"]))
    out = O.synthesis(c, n, m, lon[g["cols"]], lat[g["rows"]], kind="port")
    assert field_norm_err(out, g["values"]) < 1e-13
    z = np.load(os.path.join(GOLDEN, "synthesis_points_n120.npz"))
    out = O.synthesis(z["coef"], z["n"], z["m"], z["lon"], z["lat"], grid=False, kind="port")
    assert field_norm_err(out, z["values"]) < 1e-13
    z = np.load(os.path.join(GOLDEN, "analysis_shlib_grids.npz"))
    g = load_group(z, "n48_lon0")
    lon, lat = O.shlib_grid(int(g["nmax"]), str(g["gtype"]))
    f = np.random.default_rng(int(g["seed"])).standard_normal((int(g["B"]), len(lat), len(lon)))
    out = O.analysis(f, int(g["nmax"]), lon, lat, kind="port")
    assert field_norm_err(out, g["values"]) < 1e-13
