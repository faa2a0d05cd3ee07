"""xarray-facing plugin: ``engine='b200'`` for shxarray's ``.sh`` accessor.

Registers through the reference's own engine plugin interface (entry-point group
``shxarray.computebackends``, pyproject.toml:41-43 of the reference; base class
core/shcomputebase.py:1-20; loader core/shxarbase.py:240-259) as a sibling of ``engine='shlib'``
(src/builtin_backend/shlib.pyx:24-159) and ``engine='shtns'`` (src/shtns_backend/shtns.py:287-420).
Same method names, argument meaning and error behaviour; the accessor calls
``da.sh.synthesis(engine='b200')``, ``da.sh.analysis(nmax, engine='b200')``,
``da.sh.multiply(other, engine='b200')`` and everything layered on those three calls work unchanged.  The shlib-only
extras (gaunt, gauntReal, wigner3j, p2s: shlib.pyx:49-60) raise the base class's "not implemented" RuntimeError
(shcomputebase.py:19-20) -- the accessor defaults to engine="shlib" for them.  Unknown keyword arguments raise TypeError.

Known deliberate difference from shlib (shlib.pyx:101-108): ``lonlat_grid(lon=, lat=, gtype="point")``
returns a true point set (lon/lat sharing the dimension ``nlonlat``) -- shlib overwrites it with a grid dataset, so
equal-length lon/lat always give a grid there; here, as with engine='shtns' (shtns.py:355-363), they give points.

This module needs xarray + shxarray at import time (it is only imported through the entry point or
explicitly); everything below it (``engine.SHEngine``, ``capi``) works with numpy/torch alone.
"""
import numpy as np
import xarray as xr

from shxarray.core.cf import find_lat, find_lon, get_cfatts, get_cfglobal
from shxarray.core.sh_indexing import SHindexBase
from shxarray.core.shcomputebase import SHComputeBackendBase

from . import capi
from .engine import SHEngine, lonlat_grid as _lonlat_grid


def _stack_aux(dain, keep):
    """Stack every dimension not in ``keep`` into one leading batch axis (shtns.py:220-236 does the same, then loops
    over fields in Python -- here the whole batch goes to the device in one call).  Returns (2-D/3-D view helper)."""
    auxdims = [d for d in dain.dims if d not in keep]
    if len(auxdims) > 1:
        dain = dain.stack(auxdim=auxdims)
        auxdim = "auxdim"
    elif len(auxdims) == 0:
        auxdim = "auxsingle"
        dain = dain.expand_dims(auxdim)
    else:
        auxdim = auxdims[0]
    return dain, auxdim


def _unstack_aux(daout, auxdim):
    if auxdim == "auxdim":
        daout = daout.unstack(auxdim)
    elif auxdim == "auxsingle":
        daout = daout.squeeze(auxdim, drop=True)
    return daout


class B200Backend(SHComputeBackendBase):
    _credits = "Used backend: shxarray_b200, hand-written sm_100a CUDA spherical-harmonic transform engine (libshb200)"

    def __init__(self, device=0):
        super().__init__()
        capi.lib()   # fail loudly if the CUDA library is missing: there is no CPU fallback
        self._engine = SHEngine(device=device)

    # ------------------------------------------------------------------------------------------
    def lonlat_grid(self, nmax=None, lon=None, lat=None, gtype=None):
        """Create a lon/lat grid for this backend.  gtype: 'gauss' (default), 'regular', 'regular_lon0' (identical to
        shlib.pyx:87-96) or 'point'.  Sets ``shxarray_gtype`` on both coordinates (shlib.pyx:115-116)."""
        if gtype is None:
            # generated grids default to Gauss-Legendre; user-supplied lon/lat are labelled like shlib does (shlib.pyx:76-77)
            gtype = "gauss" if nmax is not None else "regular"
        if lon is not None and hasattr(lon, "data"):
            lon = lon.data
        if lat is not None and hasattr(lat, "data"):
            lat = lat.data
        g = _lonlat_grid(nmax=nmax, lon=lon, lat=lat, gtype=gtype)
        if gtype == "point":
            dslonlat = xr.Dataset(coords=dict(lon=("nlonlat", g["lon"]), lat=("nlonlat", g["lat"])))
        else:
            dslonlat = xr.Dataset(coords=dict(lon=g["lon"], lat=g["lat"]))
        dslonlat.lon.attrs.update(get_cfatts("longitude"))
        dslonlat.lat.attrs.update(get_cfatts("latitude"))
        dslonlat.lon.attrs["shxarray_gtype"] = gtype
        dslonlat.lat.attrs["shxarray_gtype"] = gtype
        return dslonlat

    # ------------------------------------------------------------------------------------------
    def synthesis(self, dain, dslonlat, **kwargs):
        if type(dain) != xr.DataArray:
            raise RuntimeError("input type should be a xarray.DataArray")
        if "nm" not in dain.indexes:
            raise RuntimeError("Spherical harmonic index not found in input, cannot apply synthesis operator to object")
        loninfo = find_lon(dslonlat.coords)
        latinfo = find_lat(dslonlat.coords)
        lon = np.asarray(loninfo.var.data, dtype=np.float64)
        lat = np.asarray(latinfo.var.data, dtype=np.float64)
        points = latinfo.var.dims[0] == loninfo.var.dims[0]      # synthesis.pyx:92-98
        n = dain.nm.n.data.astype(np.int32)
        m = dain.nm.m.data.astype(np.int32)
        name = dain.name
        dast, auxdim = _stack_aux(dain, ["nm"])
        dast = dast.transpose(auxdim, "nm")
        coef = np.asarray(dast.data, dtype=np.float64)           # [B, nsh], any strides
        # degree_scale=<nmax+1 factors or an IsoKernelBase> fuses an isotropic kernel into the transform
        _reject_unknown(kwargs, ("degree_scale",))
        ds = kwargs.get("degree_scale", None)
        if ds is not None and hasattr(ds, "_dsiso"):
            ds = np.asarray(ds._dsiso.sel(n=np.arange(int(n.max()) + 1)).data, dtype=np.float64)
        out = self._engine.synthesis(coef, n, m, lon, lat, points=points, degree_scale=ds)
        if points:
            pdim = latinfo.var.dims[0]
            coords = {auxdim: dast.coords[auxdim], "lat": (pdim, lat), "lon": (pdim, lon)}
            daout = xr.DataArray(out, coords=coords, dims=[auxdim, pdim])
        else:
            coords = {auxdim: dast.coords[auxdim], "lat": lat, "lon": lon}
            daout = xr.DataArray(out, coords=coords, dims=[auxdim, "lat", "lon"])
        daout = _unstack_aux(daout, auxdim)
        daout.name = name
        daout.lon.attrs = dict(get_cfatts("longitude"))
        daout.lat.attrs = dict(get_cfatts("latitude"))
        for c in ("lon", "lat"):
            if "shxarray_gtype" in dslonlat[c].attrs:
                daout[c].attrs["shxarray_gtype"] = dslonlat[c].attrs["shxarray_gtype"]
        daout.attrs.update(get_cfglobal())
        daout.attrs["history"] = "Synthesis performed using b200 backend"
        daout.attrs["comments"] = self._credits
        return daout

    # ------------------------------------------------------------------------------------------
    def analysis(self, dain, nmax, quadrature=None, **kwargs):
        """quadrature: None -> 'gauss' if the grid carries shxarray_gtype='gauss', else 'shlib' (midpoint weights, the
        values shlib itself produces: analysis.pyx:57-60,129)."""
        if type(dain) != xr.DataArray:
            raise RuntimeError("input type should be a xarray.DataArray")
        loninfo = find_lon(dain.coords)
        latinfo = find_lat(dain.coords)
        gtype = loninfo.var.attrs.get("shxarray_gtype", None)
        if quadrature is None:
            quadrature = "gauss" if gtype == "gauss" else "shlib"
        if quadrature == "shlib" and (loninfo.step is None or latinfo.step is None):
            raise RuntimeError("grid is not equidistant in lon and lat directions")
        if latinfo.var.dims[0] == loninfo.var.dims[0]:
            raise RuntimeError("Analysis is not supported for point grids")
        lonname, latname = loninfo.var.dims[0], latinfo.var.dims[0]
        name = dain.name
        dast, auxdim = _stack_aux(dain, [lonname, latname])
        dast = dast.transpose(auxdim, latname, lonname)          # a view: strides go straight to the C ABI
        grid = np.asarray(dast.data, dtype=np.float64)
        lon = np.asarray(loninfo.var.data, dtype=np.float64)
        lat = np.asarray(latinfo.var.data, dtype=np.float64)
        _reject_unknown(kwargs, ("degree_scale",))
        ds = kwargs.get("degree_scale", None)
        if ds is not None and hasattr(ds, "_dsiso"):
            ds = np.asarray(ds._dsiso.sel(n=np.arange(int(nmax) + 1)).data, dtype=np.float64)
        out = self._engine.analysis(grid, int(nmax), lon, lat, quadrature=quadrature, degree_scale=ds)
        coords = {auxdim: dast.coords[auxdim], SHindexBase.name: SHindexBase.nm_mi(int(nmax), 0)}
        daout = xr.DataArray(out, coords=coords, dims=[auxdim, SHindexBase.name])
        daout = _unstack_aux(daout, auxdim)
        daout.name = name
        daout.attrs.update(get_cfglobal())
        daout.attrs["history"] = "Analysis performed using b200 backend"
        daout.attrs["comments"] = self._credits
        return daout

    # ------------------------------------------------------------------------------------------
    def multiply(self, dash1, dash2, method="spatial", **kwargs):
        """Product of two SH fields through the spatial domain (shtns.py:394-420; shlib.pyx:140-152 method="spatial"): Gauss
        grid of nmax1+nmax2, two syntheses, pointwise product, Gauss-quadrature analysis -- all four steps on the device
        (shb200_multiply), only the coefficients cross PCIe.  method="spectral" (Gaunt coefficients, shlib.pyx:130-139) is
        the shlib engine's own algorithm and is not provided here."""
        if method != "spatial":
            self._notImplemented(f"multiply(method={method!r})")
        _reject_unknown(kwargs, ())
        nmax1, nmax2 = int(dash1.sh.nmax), int(dash2.sh.nmax)
        nmax = nmax1 + nmax2
        name = dash1.name

        def full_triangle(da, nm_max):
            """[B, (nmax+1)^2] in standard nm order (missing coefficients = 0, e.g. nmin > 0), aux dims stacked"""
            dast, auxdim = _stack_aux(da, ["nm"])
            dast = dast.transpose(auxdim, "nm")
            n = dast.nm.n.data.astype(np.int64)
            m = dast.nm.m.data.astype(np.int64)
            full = np.zeros((dast.sizes[auxdim], (nm_max + 1) ** 2))
            full[:, n * n + n + m] = np.asarray(dast.data, dtype=np.float64)
            return full, dast, auxdim

        a, dast1, aux1 = full_triangle(dash1, nmax1)
        b, dast2, aux2 = full_triangle(dash2, nmax2)
        # broadcasting over the auxiliary dimension like the reference's grid product: equal batch sizes or a single field
        if a.shape[0] != b.shape[0]:
            if a.shape[0] == 1:
                a = np.repeat(a, b.shape[0], axis=0)
                dast1, aux1 = dast2, aux2
            elif b.shape[0] == 1:
                b = np.repeat(b, a.shape[0], axis=0)
            else:
                raise RuntimeError("multiply: auxiliary dimensions of the two inputs do not match")
        out = self._engine.multiply(a, nmax1, b, nmax2)
        coords = {aux1: dast1.coords[aux1], SHindexBase.name: SHindexBase.nm_mi(nmax, 0)}
        daout = xr.DataArray(out, coords=coords, dims=[aux1, SHindexBase.name])
        daout = _unstack_aux(daout, aux1)
        daout.name = name
        daout.attrs.update(get_cfglobal())
        daout.attrs["history"] = "Multiplication performed using b200 backend"
        daout.attrs["comments"] = self._credits
        return daout

    # ------------------------------------------------------------------------------------------
    def apply_mask(self, dash, damask, nmax_out=None, quadrature=None, syn_scale=None, ana_scale=None):
        """Grid-space ocean-function operator: ``analysis(damask x synthesis(dash))`` with the grid kept on the device
        (SHEngine.apply_mask).  Replaces ``SpectralSeaLevelSolver.oceanf(load)`` (spectralsealevel.py:97-104: a product-to-sum
        matrix that exists for nmax <= 150 only).  damask: DataArray on a lon/lat grid (e.g. the ocean function rasterised on
        ``lonlat_grid(nmax)``); dash: full-triangle SH DataArray."""
        loninfo = find_lon(damask.coords)
        latinfo = find_lat(damask.coords)
        gtype = loninfo.var.attrs.get("shxarray_gtype", None)
        if quadrature is None:
            quadrature = "gauss" if gtype == "gauss" else "shlib"
        lon = np.asarray(loninfo.var.data, dtype=np.float64)
        lat = np.asarray(latinfo.var.data, dtype=np.float64)
        mask = np.asarray(damask.transpose(latinfo.var.dims[0], loninfo.var.dims[0]).data, dtype=np.float64)
        nmax = int(dash.sh.nmax)
        nmax_out = nmax if nmax_out is None else int(nmax_out)
        dast, auxdim = _stack_aux(dash, ["nm"])
        dast = dast.transpose(auxdim, "nm")
        n = dast.nm.n.data.astype(np.int64)
        m = dast.nm.m.data.astype(np.int64)
        full = np.zeros((dast.sizes[auxdim], (nmax + 1) ** 2))
        full[:, n * n + n + m] = np.asarray(dast.data, dtype=np.float64)
        out = self._engine.apply_mask(full, nmax, mask, lon, lat, nmax_out=nmax_out, quadrature=quadrature, syn_scale=syn_scale, ana_scale=ana_scale)
        coords = {auxdim: dast.coords[auxdim], SHindexBase.name: SHindexBase.nm_mi(nmax_out, 0)}
        daout = _unstack_aux(xr.DataArray(out, coords=coords, dims=[auxdim, SHindexBase.name]), auxdim)
        daout.name = dash.name
        daout.attrs.update(get_cfglobal())
        daout.attrs["history"] = "Grid-space mask operator applied using b200 backend"
        daout.attrs["comments"] = self._credits
        return daout

    # ---- shlib-only extras (xr_accessor.py:36-51,203-207 default to engine="shlib" for these) ----------------------------
    def gaunt(self, *a, **k):
        self._notImplemented("gaunt")

    def gauntReal(self, *a, **k):
        self._notImplemented("gauntReal")

    def wigner3j(self, *a, **k):
        self._notImplemented("wigner3j")

    def p2s(self, *a, **k):
        self._notImplemented("p2s")


def _reject_unknown(kwargs, allowed):
    bad = [k for k in kwargs if k not in allowed]
    if bad:
        raise TypeError(f"b200 backend: unknown keyword argument(s) {bad}; supported: {list(allowed)}")


def register(name="b200", device=0):
    """Make the engine available as ``engine=name`` without installing entry-point metadata (tests, notebooks):
    inserts an instance into shxarray's engine cache (core/shxarbase.py:14,255)."""
    from shxarray.core import shxarbase
    shxarbase.loaded_engines[name] = B200Backend(device=device)
    return shxarbase.loaded_engines[name]
