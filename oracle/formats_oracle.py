"""CPU restatement of the reference's coefficient-file readers and polygon masks -- TEST INFRASTRUCTURE ONLY (tests/ and
__graft_entry__.smoke() may import it; the product path in shxarray_b200/ never does).

  read_records  : the record loops of io/icgem.py:94-128 (key b"gfc", D exponents replaced) and io/gsmv6.py:66-90
                  (key b"GRCOF2"), line by line with bytes.split(), int() and float() exactly like the reference; returns the
                  nm list in FILE order plus cnm / sigcnm (what readIcgem / readGSMv6 put into their Dataset).
  icgem_header  : io/icgem.py:37-90 (name/value pairs between begin_of_head and end_of_head, nmax rule, epoch hack).
  polygon_masks : geom/polygons.py:66-93 -- grid points contained in each polygon.  The reference asks geopandas/shapely
                  (absent here); the restatement is the even-odd crossing-number test in planar lon/lat coordinates, which is
                  what "contains" means for points off the boundary.  Pinned by analytic cases in tests/ (rectangles, a polygon
                  with a hole), not by shapely: parity of boundary points is therefore UNPINNED and the tests avoid them.
"""
import re
import sys
from datetime import datetime, timedelta

import numpy as np


def icgem_header(lines, nmaxstop=sys.maxsize):
    """lines: iterator over the file's byte lines; consumes the header.  Returns (attr, time)  [icgem.py:37-90]"""
    hdr = {}
    for ln in lines:
        if b"begin_of_head" in ln:
            continue
        if b"end_of_head" in ln:
            break
        spl = ln.decode("utf-8").split()
        if len(spl) == 2:
            hdr[spl[0]] = spl[1]
    attr = {}
    try:
        nmaxsupp = int(hdr["max_degree"])
        attr["nmaxfile"] = nmaxsupp
        attr["nmax"] = nmaxsupp if nmaxsupp < nmaxstop else nmaxstop
        attr["format"] = hdr["format"] if "format" in hdr else "icgem"
        if "norm" in hdr:
            attr["norm"] = hdr["norm"]
        attr["gm"] = float(hdr["earth_gravity_constant"])
        attr["re"] = float(hdr["radius"])
        attr["modelname"] = hdr["modelname"]
    except KeyError:
        pass
    time = None
    if "modelname" in hdr:
        try:
            time = [datetime.strptime(hdr["modelname"][-7:], "%Y-%m") + timedelta(days=14.5)]
        except ValueError:
            time = None
    return attr, time


def read_records(lines, nmax, key=b"gfc", d_exponent=True, nmaxstop=sys.maxsize, always_sigma=False):
    """The record loop of icgem.py:94-128 / gsmv6.py:66-90.  Returns (nm list in file order, cnm, sigcnm or None)."""
    nsh = (nmax + 1) ** 2
    cnm = np.zeros([max(nsh, 1)])
    sigcnm = np.zeros([max(nsh, 1)])
    ncount = 0
    nm = []
    rx = re.compile(b"^" + key)
    ncolumns = -1
    for ln in lines:
        if rx.match(ln):
            lnspl = (ln.replace(b"D", b"E") if d_exponent else ln).split()
            if ncolumns < 0:
                ncolumns = len(lnspl)
            n = int(lnspl[1])
            if n > nmaxstop:
                if ncount > nsh:
                    break
                continue
            m = int(lnspl[2])
            if ncount >= len(cnm):
                cnm = np.concatenate([cnm, np.zeros_like(cnm)])
                sigcnm = np.concatenate([sigcnm, np.zeros_like(sigcnm)])
            cnm[ncount] = float(lnspl[3])
            if always_sigma or ncolumns >= 6:
                sigcnm[ncount] = float(lnspl[5])
            nm.append((n, m))
            ncount += 1
            if m != 0:
                if ncount >= len(cnm):
                    cnm = np.concatenate([cnm, np.zeros_like(cnm)])
                    sigcnm = np.concatenate([sigcnm, np.zeros_like(sigcnm)])
                cnm[ncount] = float(lnspl[4])
                if always_sigma or ncolumns >= 6:
                    sigcnm[ncount] = float(lnspl[6])
                nm.append((n, -m))
                ncount += 1
    has_sig = always_sigma or ncolumns >= 6
    return nm, cnm[:ncount], (sigcnm[:ncount] if has_sig else None)


def dense(nm, values, nmax):
    """file-order records -> dense vector at position n^2+n+m (sh_indexing.py:66); later records overwrite earlier ones"""
    out = np.zeros((nmax + 1) ** 2)
    for (n, m), v in zip(nm, values):
        if n <= nmax:
            out[n * n + n + m] = v
    return out


def polygon_masks(polygons, lon, lat):
    """polygons: list of polygons, each a list of rings ([nv, 2] arrays of (lon, lat), closed implicitly).
    Returns [npoly, nlat, nlon] of 0/1.  Grid longitudes are mapped to [-180, 180) when the grid starts at >= 0
    (polygons.py:76-81)."""
    lon = np.asarray(lon, dtype=np.float64)
    lat = np.asarray(lat, dtype=np.float64)
    x = lon if lon.min() < 0 else (lon + 180) % 360 - 180
    X, Y = np.meshgrid(x, lat, indexing="xy")            # [nlat, nlon]
    out = np.zeros((len(polygons), len(lat), len(lon)))
    for p, rings in enumerate(polygons):
        inside = np.zeros(X.shape, dtype=bool)
        for ring in rings:
            ring = np.asarray(ring, dtype=np.float64)
            nv = len(ring)
            for a in range(nv):
                x0, y0 = ring[a]
                x1, y1 = ring[(a + 1) % nv]
                cross = (y0 > Y) != (y1 > Y)
                if y1 != y0:
                    xc = (x1 - x0) / (y1 - y0) * (Y - y0) + x0
                    inside ^= cross & (X < xc)
        out[p] = inside
    return out
