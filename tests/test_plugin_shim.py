"""The plugin class ``B200Backend`` (shxarray_b200/xr_backend.py) executed end to end, in the shape of the reference's
tests/test_shtns_backend.py:29-185 (Gauss grid, regular_lon0, points, aux dims + nmin > 0, analysis round trips, multiply).

xarray and shxarray are not installable here, so these tests run the plugin against tests/xr_shim.py -- a stand-in for the
few xarray / shxarray.core calls the plugin makes -- and compare the numbers with the oracle (the shlib arithmetic), which
is what ``engine='shlib'`` would return.  Where the real packages exist, tests/test_xr_backend.py exercises the same class
through the real ``.sh`` accessor instead.
"""
import numpy as np
import pytest

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import xr_shim  # noqa: E402

SHIM = xr_shim.install()
if not SHIM:
    pytest.skip("real xarray + shxarray present: tests/test_xr_backend.py covers the plugin", allow_module_level=True)

from oracle import oracle as O  # noqa: E402

TOL = 1e-12


def _backend():
    from shxarray_b200.xr_backend import B200Backend
    return B200Backend(device=0)


def _coeffs(nmax, nmin=0, aux=None, seed=0, nm_first=False):
    aux = aux or {}
    nsh = (nmax + 1) ** 2 - nmin ** 2
    rng = np.random.default_rng(seed)
    shape = tuple(len(v) for v in aux.values()) + (nsh,)
    vals = rng.standard_normal(shape)
    if nm_first:
        vals = np.moveaxis(vals, -1, 0)
    return xr_shim.sh_dataarray(vals, nmax, nmin, aux=aux, nm_first=nm_first)


# ---- host-only behaviour (no kernel launch): grids, errors, registration ----------------------------------------------
def test_shim_stack_unstack_roundtrip():
    da = xr_shim.DataArray(np.arange(24.0).reshape(2, 3, 4), coords={"a": [10, 20], "b": [1, 2, 3], "nm": np.arange(4)}, dims=["a", "nm", "b"][0:1] + ["b", "nm"])
    st = da.stack(auxdim=["a", "b"])
    assert st.dims == ("nm", "auxdim") and st.shape == (4, 6)
    back = st.unstack("auxdim")
    assert back.dims == ("nm", "a", "b")
    assert np.array_equal(back.transpose("a", "b", "nm").data, da.data)
    assert list(back.coords["a"].data) == [10, 20]


def test_lonlat_grid_types_and_attrs():
    be = _backend()
    for gtype, nlat, nlon in (("gauss", 64, 128), (None, 64, 128), ("regular", 180, 360), ("regular_lon0", 180, 360)):
        ds = be.lonlat_grid(nmax=60, gtype=gtype)
        want = "gauss" if gtype is None else gtype                      # generated grids default to Gauss
        assert ds.lon.attrs["shxarray_gtype"] == want and ds.lat.attrs["shxarray_gtype"] == want
        assert ds.lon.attrs["standard_name"] == "longitude" and ds.lat.attrs["units"] == "degrees_north"
        if gtype in ("regular", "regular_lon0"):
            # shlib.pyx:87-96: 1 degree at nmax 60 (BASELINE cfg1: 180 x 360); cell centres, or longitudes starting at 0
            assert len(ds.lat.data) == nlat and len(ds.lon.data) == nlon
            assert ds.lon.data[0] == (-179.5 if gtype == "regular" else 0.0) and ds.lat.data[0] == -89.5
        else:
            assert len(ds.lat.data) >= 61 and len(ds.lon.data) >= 121
    pts = be.lonlat_grid(lon=[0.0, 10.0, 20.0], lat=[-5.0, 0.0, 5.0], gtype="point")
    assert pts.lon.dims == ("nlonlat",) and pts.lat.dims == ("nlonlat",) and pts.lon.attrs["shxarray_gtype"] == "point"
    with pytest.raises(ValueError):
        be.lonlat_grid(nmax=60, gtype="no-such-grid")


def test_errors_and_extras_follow_the_reference_convention():
    be = _backend()
    ds = be.lonlat_grid(nmax=8, gtype="regular")
    with pytest.raises(RuntimeError):                                   # synthesis.pyx:52-58
        be.synthesis(np.zeros(4), ds)
    nosh = xr_shim.DataArray(np.zeros((3, 4)), coords={"t": np.arange(3), "x": np.arange(4)}, dims=["t", "x"])
    with pytest.raises(RuntimeError):
        be.synthesis(nosh, ds)
    with pytest.raises(TypeError):                                      # unknown kwargs are rejected loudly
        be.synthesis(_coeffs(8), ds, no_such_option=1)
    for extra in ("gaunt", "gauntReal", "wigner3j", "p2s"):            # shcomputebase.py:19-20
        with pytest.raises(RuntimeError):
            getattr(be, extra)(1, 2, 3)
    with pytest.raises(RuntimeError):
        be.multiply(_coeffs(4), _coeffs(4), method="spectral")
    pts = be.lonlat_grid(lon=[0.0, 10.0], lat=[-5.0, 5.0], gtype="point")
    pgrid = xr_shim.DataArray(np.zeros(2), coords={"lon": ("nlonlat", pts.lon.data), "lat": ("nlonlat", pts.lat.data)}, dims=["nlonlat"])
    with pytest.raises(RuntimeError):                                   # analysis of a point set
        be.analysis(pgrid, 4)


def test_register_puts_the_engine_into_the_loader_cache():
    from shxarray.core import shxarbase
    from shxarray_b200.xr_backend import register
    eng = register("b200")
    assert shxarbase.loaded_engines["b200"] is eng and hasattr(eng, "synthesis") and hasattr(eng, "analysis")


# ---- GPU: the four cases of test_shtns_backend.py against the shlib arithmetic --------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("gtype", ["regular", "regular_lon0", "gauss"])
def test_synthesis_grid_aux_dim_and_nmin(gtype):
    """aux dim `time` + nmin = 3 (test_shtns_backend.py:100-124): the nm list is a subset, the output has dims time, lat, lon"""
    be = _backend()
    nmax, nmin = 40, 3
    da = _coeffs(nmax, nmin, aux={"time": np.arange(3)}, seed=1)
    da.name = "tws"
    ds = be.lonlat_grid(nmax=nmax, gtype=gtype)
    out = be.synthesis(da, ds)
    assert type(out) is xr_shim.DataArray and out.dims == ("time", "lat", "lon") and out.name == "tws"
    assert np.array_equal(out.lon.data, ds.lon.data) and np.array_equal(out.lat.data, ds.lat.data)
    assert out.lon.attrs["shxarray_gtype"] == gtype and out.lat.attrs["standard_name"] == "latitude"
    assert "b200" in out.attrs["history"] and "libshb200" in out.attrs["comments"]
    n, m = O.nm_index(nmax, nmin)
    rows = np.arange(0, len(ds.lat.data), 7)
    ref = O.synthesis(np.ascontiguousarray(da.data), n, m, ds.lon.data, ds.lat.data[rows], kind="port")     # [rows, lon, time]
    err = np.abs(out.data[:, rows] - np.moveaxis(ref, 2, 0)).max() / np.abs(ref).max()
    assert err <= TOL, err


@pytest.mark.gpu
def test_synthesis_nm_first_two_aux_dims_and_no_aux_dim():
    be = _backend()
    nmax = 24
    ds = be.lonlat_grid(nmax=nmax, gtype="regular")
    da = _coeffs(nmax, aux={"time": np.arange(2), "model": np.array([7, 8, 9])}, seed=2, nm_first=True)   # dims nm, time, model
    out = be.synthesis(da, ds)
    assert out.dims == ("lat", "lon", "time", "model") or set(out.dims) == {"lat", "lon", "time", "model"}
    n, m = O.nm_index(nmax)
    flat = np.moveaxis(da.data, 0, -1).reshape(6, -1)
    ref = O.synthesis(flat, n, m, ds.lon.data, ds.lat.data, kind="port")                                   # [lat, lon, 6]
    got = out.transpose("lat", "lon", "time", "model").data.reshape(len(ds.lat.data), len(ds.lon.data), 6)
    assert np.abs(got - ref).max() / np.abs(ref).max() <= TOL
    single = _coeffs(nmax, seed=3)                                                                         # dims (nm,)
    o1 = be.synthesis(single, ds)
    assert o1.dims == ("lat", "lon")
    r1 = O.synthesis(single.data[None], n, m, ds.lon.data, ds.lat.data, kind="port")[:, :, 0]
    assert np.abs(o1.data - r1).max() / np.abs(r1).max() <= TOL


@pytest.mark.gpu
def test_synthesis_points():
    """lon and lat sharing one dimension -> point synthesis (synthesis.pyx:92-98,186-189; test_shtns_backend.py:73-98)"""
    be = _backend()
    nmax = 30
    rng = np.random.default_rng(4)
    lon, lat = rng.uniform(-180, 180, 50), rng.uniform(-90, 90, 50)
    lat[:2] = (90.0, -90.0)
    pts = be.lonlat_grid(lon=lon, lat=lat, gtype="point")
    da = _coeffs(nmax, aux={"time": np.arange(4)}, seed=5)
    out = be.synthesis(da, pts)
    assert out.dims == ("time", "nlonlat") and out.lon.dims == ("nlonlat",)
    n, m = O.nm_index(nmax)
    ref = O.synthesis(np.ascontiguousarray(da.data), n, m, lon, lat, grid=False, kind="port")            # [npoints, time]
    assert np.abs(out.data - ref.T).max() / np.abs(ref).max() <= TOL


@pytest.mark.gpu
def test_analysis_regular_grid_matches_shlib_and_gauss_roundtrip():
    be = _backend()
    nmax = 20
    da = _coeffs(nmax, aux={"time": np.arange(3)}, seed=6)
    # (1) regular grid: shlib's midpoint weights (analysis.pyx:57-60,129), compared with the shlib arithmetic itself
    ds = be.lonlat_grid(nmax=nmax, gtype="regular")
    grid = be.synthesis(da, ds)
    got = be.analysis(grid, nmax)
    assert got.dims == ("time", "nm") and got.coords["nm"].index.equals(xr_shim.SHindexBase.nm_mi(nmax, 0))
    ref = O.analysis(np.ascontiguousarray(grid.data), nmax, ds.lon.data, ds.lat.data, kind="port")
    assert (np.abs(got.data - ref).max(axis=1) / np.abs(ref).max(axis=1)).max() <= TOL
    # lat/lon first, aux last (shlib's own output layout, analysis.pyx:100-120) gives the same numbers
    got2 = be.analysis(grid.transpose("lat", "lon", "time"), nmax)
    assert np.array_equal(got2.data, got.data)
    # (2) Gauss grid: exact quadrature picked from shxarray_gtype -> the coefficients come back
    dg = be.lonlat_grid(nmax=nmax)
    back = be.analysis(be.synthesis(da, dg), nmax)
    assert np.abs(back.data - da.data).max() <= 1e-12 * np.abs(da.data).max()
    assert "Analysis" in back.attrs["history"]


@pytest.mark.gpu
def test_multiply_runs_on_the_device_and_matches_the_grid_product():
    """shtns.py:394-420 pattern: Gauss grid of nmax1 + nmax2, product, analysis -- compared with the same steps made of oracle
    syntheses and the plugin's own (oracle-checked) Gauss analysis"""
    be = _backend()
    n1, n2 = 12, 9
    a = _coeffs(n1, aux={"time": np.arange(2)}, seed=7)
    b = _coeffs(n2, seed=8)                                             # a single field broadcasts over time
    prod = be.multiply(a, b)
    assert prod.dims == ("time", "nm") and prod.sh.nmax == n1 + n2
    dg = be.lonlat_grid(nmax=n1 + n2)
    na, ma = O.nm_index(n1)
    nb, mb = O.nm_index(n2)
    ga = O.synthesis(np.ascontiguousarray(a.data), na, ma, dg.lon.data, dg.lat.data, kind="port")      # [lat, lon, 2]
    gb = O.synthesis(b.data[None], nb, mb, dg.lon.data, dg.lat.data, kind="port")                       # [lat, lon, 1]
    gp = xr_shim.DataArray(np.moveaxis(ga * gb, 2, 0), coords={"time": np.arange(2), "lat": dg.lat, "lon": dg.lon}, dims=["time", "lat", "lon"])
    ref = be.analysis(gp, n1 + n2)
    assert np.abs(prod.data - ref.data).max() <= 1e-12 * np.abs(ref.data).max()
    # the constant field 1 is the identity of the product
    one = xr_shim.sh_dataarray(np.eye(1, 1).ravel(), 0)
    same = be.multiply(a, one)
    assert np.abs(same.data - a.data).max() <= 1e-12 * np.abs(a.data).max()
