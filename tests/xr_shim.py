"""Stand-in for the slice of xarray + shxarray.core that ``shxarray_b200/xr_backend.py`` touches.  TEST DOUBLE, not xarray.

This image has neither xarray nor shxarray (no network), so the plugin class ``B200Backend`` could never execute
(VERDICT r01, row b1).  ``install()`` registers minimal modules named ``xarray``, ``shxarray.core.cf``,
``shxarray.core.sh_indexing``, ``shxarray.core.shcomputebase`` and ``shxarray.core.shxarbase`` -- ONLY when the real
packages are absent -- with the semantics the plugin relies on:

* ``DataArray(data, coords=, dims=)`` with named dimensions, 1-D coordinates (plain or pandas.MultiIndex), ``attrs``,
  ``name``, ``stack`` / ``unstack`` / ``expand_dims`` / ``squeeze`` / ``transpose``, ``.indexes``, ``.sizes``,
  attribute and item access to coordinates, and the ``.sh.nmax`` accessor property (xr_accessor / shxarbase.py);
* ``Dataset(coords=)`` as a bag of coordinates;
* ``find_lon`` / ``find_lat`` (cf.py:51-75: name lookup + uniform step detection), ``get_cfatts`` / ``get_cfglobal``;
* ``SHindexBase.nm_mi`` (sh_indexing.py:48-67: [(n, m) for n in nmin..nmax for m in -n..n], levels "n", "m");
* ``SHComputeBackendBase._notImplemented`` (shcomputebase.py:19-20) and ``shxarbase.loaded_engines``.

With real xarray + shxarray installed, tests/test_xr_backend.py runs against them instead and this file is unused.
"""
import sys
import types
from collections import namedtuple

import numpy as np
import pandas as pd


class Coord:
    """1-D coordinate variable: ``data`` (numpy values or a MultiIndex), ``dims`` (1-tuple), ``attrs``."""

    def __init__(self, name, dim, data, attrs=None):
        self.name, self.dims = name, (dim,)
        self.index = data if isinstance(data, pd.MultiIndex) else None
        self._data = None if self.index is not None else np.asarray(data)
        self.attrs = dict(attrs or {})

    @property
    def data(self):
        return self.index.values if self.index is not None else self._data

    values = data

    def __len__(self):
        return len(self.index) if self.index is not None else len(self._data)

    def __getattr__(self, level):               # da.nm.n -> level values of a MultiIndex coordinate
        idx = self.__dict__.get("index")
        if idx is not None and level in idx.names:
            return Coord(level, self.dims[0], np.asarray(idx.get_level_values(level)))
        raise AttributeError(level)

    def take(self, sel):
        return Coord(self.name, self.dims[0], self.index[sel] if self.index is not None else self._data[sel], self.attrs)


class _Coords(dict):
    def __init__(self, owner):
        super().__init__()
        self._owner = owner

    def __missing__(self, key):                 # a dimension without coordinate reads as 0..n-1, like xarray's virtual index
        if key in self._owner.dims:
            return Coord(key, key, np.arange(self._owner.sizes[key]))
        raise KeyError(key)


class _SHAccessor:
    def __init__(self, da):
        self._da = da

    @property
    def nmax(self):
        return int(np.max(self._da.coords["nm"].n.data))

    @property
    def nmin(self):
        return int(np.min(self._da.coords["nm"].n.data))


class DataArray:
    def __init__(self, data, coords=None, dims=None, name=None, attrs=None):
        self.data = np.asarray(data)
        self.dims = tuple(dims) if dims is not None else tuple(f"dim_{i}" for i in range(self.data.ndim))
        assert len(self.dims) == self.data.ndim, (self.dims, self.data.shape)
        self.name, self.attrs = name, dict(attrs or {})
        self.coords = _Coords(self)
        for key, val in (coords or {}).items():
            if isinstance(val, Coord):
                c = Coord(key, val.dims[0], val.index if val.index is not None else val._data, val.attrs)
            elif isinstance(val, tuple):
                c = Coord(key, val[0], val[1])
            else:
                c = Coord(key, key, val)
            assert c.dims[0] in self.dims and len(c) == self.sizes[c.dims[0]], (key, c.dims, len(c), self.sizes)
            self.coords[key] = c

    values = property(lambda self: self.data)
    shape = property(lambda self: self.data.shape)
    sizes = property(lambda self: dict(zip(self.dims, self.data.shape)))
    sh = property(lambda self: _SHAccessor(self))

    @property
    def indexes(self):
        return {k: (c.index if c.index is not None else pd.Index(c.data)) for k, c in self.coords.items() if k == c.dims[0]}

    def __getattr__(self, key):
        coords = self.__dict__.get("coords")
        if coords is not None and (key in coords or key in self.__dict__.get("dims", ())):
            return coords[key]
        raise AttributeError(key)

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.coords[key]
        raise TypeError("the stand-in DataArray only supports coordinate lookup by name")

    def _with(self, data, dims, drop=()):
        keep = {k: c for k, c in self.coords.items() if c.dims[0] in dims and k not in drop}
        return DataArray(data, coords=keep, dims=dims, name=self.name, attrs=self.attrs)

    def transpose(self, *dims):
        assert sorted(dims) == sorted(self.dims), (dims, self.dims)
        return self._with(self.data.transpose([self.dims.index(d) for d in dims]), dims)     # a view, like xarray

    T = property(lambda self: self.transpose(*self.dims[::-1]))

    def expand_dims(self, dim):
        return self._with(self.data[None], (dim,) + self.dims)

    def squeeze(self, dim, drop=True):
        ax = self.dims.index(dim)
        assert self.data.shape[ax] == 1
        return self._with(np.squeeze(self.data, ax), tuple(d for d in self.dims if d != dim), drop=(dim,))

    def stack(self, **kw):
        (new, olds), = kw.items()
        rest = tuple(d for d in self.dims if d not in olds)
        moved = self.transpose(*rest, *olds)
        shape = moved.data.shape[:len(rest)] + (-1,)
        mi = pd.MultiIndex.from_product([np.asarray(self.coords[d].data) for d in olds], names=list(olds))
        out = DataArray(moved.data.reshape(shape), coords={k: c for k, c in self.coords.items() if c.dims[0] in rest},
                        dims=rest + (new,), name=self.name, attrs=self.attrs)
        out.coords[new] = Coord(new, new, mi)
        return out

    def unstack(self, dim):
        mi = self.coords[dim].index
        levels = [np.asarray(lv) for lv in mi.levels]
        assert len(mi) == int(np.prod([len(lv) for lv in levels])), "the stand-in only unstacks full products"
        ax = self.dims.index(dim)
        moved = np.moveaxis(self.data, ax, -1)
        full = np.empty(moved.shape[:-1] + tuple(len(lv) for lv in levels), dtype=self.data.dtype)
        full[(Ellipsis,) + tuple(np.asarray(c) for c in mi.codes)] = moved
        rest = tuple(d for d in self.dims if d != dim)
        out = DataArray(full, coords={k: c for k, c in self.coords.items() if c.dims[0] in rest},
                        dims=rest + tuple(mi.names), name=self.name, attrs=self.attrs)
        for nm_, lv in zip(mi.names, levels):
            out.coords[nm_] = Coord(nm_, nm_, lv)
        return out


class Dataset:
    def __init__(self, coords=None):
        self.coords = {}
        for key, val in (coords or {}).items():
            self.coords[key] = Coord(key, val[0], val[1]) if isinstance(val, tuple) else Coord(key, key, val)

    def __getattr__(self, key):
        coords = self.__dict__.get("coords")
        if coords is not None and key in coords:
            return coords[key]
        raise AttributeError(key)

    def __getitem__(self, key):
        return self.coords[key]

    @property
    def sizes(self):
        return {c.dims[0]: len(c) for c in self.coords.values()}


# ---- shxarray.core pieces -------------------------------------------------------------------------------------------
CoordInfo = namedtuple("CoordInfo", ["min", "max", "step", "direction", "var"])


def _find_coord(coords, names):
    """cf.py:14-47: first coordinate with one of the accepted names; step = the uniform spacing or None"""
    for nm_ in names:
        if nm_ in coords:
            var = coords[nm_]
            v = np.asarray(var.data, dtype=np.float64)
            step = None
            if len(v) > 1:
                d = np.diff(v)
                if np.allclose(d, d[0], rtol=0, atol=1e-9 * max(1.0, abs(d[0]))):
                    step = float(d[0])
            return CoordInfo(float(v.min()), float(v.max()), step, "ascending" if len(v) < 2 or v[1] > v[0] else "descending", var)
    raise KeyError(f"no coordinate named one of {names}")


def find_lon(coords):
    return _find_coord(coords, ["lon", "Longitude", "x", "longitude"])


def find_lat(coords):
    return _find_coord(coords, ["lat", "Latitude", "y", "latitude"])


def get_cfatts(standard_name):
    return {"longitude": dict(units="degrees_east", standard_name="longitude", long_name="Longitude"),
            "latitude": dict(units="degrees_north", standard_name="latitude", long_name="Latitude")}[standard_name]


def get_cfglobal():
    return {"Conventions": "CF-1.8", "source": "stand-in for shxarray.core.cf.get_cfglobal (tests/xr_shim.py)"}


class SHindexBase:
    name = "nm"

    @staticmethod
    def nm_mi(nmax, nmin=0):
        return pd.MultiIndex.from_tuples([(n, m) for n in range(nmin, nmax + 1) for m in range(-n, n + 1)], names=["n", "m"])


class SHComputeBackendBase:
    def __init__(self):
        pass

    def _notImplemented(self, methodname):
        raise RuntimeError(f"Method '{methodname}' is not implemented in backend {self.__class__.__name__}")


def sh_dataarray(values, nmax, nmin=0, aux=None, name="cnm", nm_first=False):
    """Coefficient DataArray the way xr.DataArray.sh.zeros(nmax, auxcoords=...) lays it out: aux dims then nm (or nm first)."""
    aux = aux or {}
    mi = SHindexBase.nm_mi(nmax, nmin)
    coords = {k: np.asarray(v) for k, v in aux.items()}
    coords["nm"] = Coord("nm", "nm", mi)
    dims = (("nm",) + tuple(aux)) if nm_first else (tuple(aux) + ("nm",))
    return DataArray(values, coords=coords, dims=dims, name=name)


def install():
    """Register the stand-in modules unless the real packages import.  Returns True when the stand-ins are in use."""
    try:
        import xarray  # noqa: F401
        import shxarray  # noqa: F401
        return False
    except ImportError:
        pass
    this = sys.modules[__name__]
    xr = types.ModuleType("xarray")
    xr.DataArray, xr.Dataset, xr.__shim__ = DataArray, Dataset, True
    pkgs = {"xarray": xr, "shxarray": types.ModuleType("shxarray"), "shxarray.core": types.ModuleType("shxarray.core")}
    for sub, names in (("cf", ["find_lat", "find_lon", "get_cfatts", "get_cfglobal"]), ("sh_indexing", ["SHindexBase"]),
                       ("shcomputebase", ["SHComputeBackendBase"])):
        mod = types.ModuleType(f"shxarray.core.{sub}")
        for nm_ in names:
            setattr(mod, nm_, getattr(this, nm_))
        pkgs[f"shxarray.core.{sub}"] = mod
    base = types.ModuleType("shxarray.core.shxarbase")
    base.loaded_engines = {}
    pkgs["shxarray.core.shxarbase"] = base
    pkgs["shxarray"].__path__, pkgs["shxarray.core"].__path__ = [], []
    pkgs["shxarray"].core = pkgs["shxarray.core"]
    for k, v in pkgs.items():
        sys.modules[k] = v
        if k.startswith("shxarray.core."):
            setattr(pkgs["shxarray.core"], k.rsplit(".", 1)[1], v)
    return True

