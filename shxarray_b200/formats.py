"""Data formats either side of the transform with the heavy part on the device.

* ``read_icgem`` / ``read_gsmv6`` -- the reference's readers (io/icgem.py:28-164, io/gsmv6.py:16-107) walk the record block
  line by line in Python (minutes for a 5' model); here the host only parses the small header, the record block goes to the
  GPU as bytes and ONE kernel (shb200_gfc_parse) turns it into the dense nm vector the engine consumes (position n^2+n+m,
  sh_indexing.py:66), with correctly rounded decimal conversion (bit-identical to float()).  Lines the kernel hands back
  (exotic literals, > 19 significant digits) are converted by the host with the reference's own int()/float() calls.
* ``read_icgem_batch`` -- many files (a time series of monthly solutions) -> one [B, nsh] device array.
* ``polygon_masks`` / ``polygon2sh`` -- geom/polygons.py:14-97: polygons -> 0/1 masks on the engine's grid -> SH analysis,
  rasterised on the device (shb200_polygon_mask) instead of a geopandas spatial-index query per polygon.
* ``scatter_nm`` -- (n, m, value) records in any order -> dense nm vectors.

No CPU fallback: everything here needs libshb200.so and a CUDA device.
"""
import ctypes as C
import gzip
import sys
from datetime import datetime, timedelta

import numpy as np

from . import capi


def _read_bytes(src):
    if isinstance(src, (bytes, bytearray, memoryview)):
        return bytes(src)
    if isinstance(src, str):
        if src.endswith(".gz"):
            with gzip.open(src, "rb") as f:
                return f.read()
        with open(src, "rb") as f:
            return f.read()
    return src.read()


def _icgem_header(data, nmaxstop):
    """io/icgem.py:37-90 on the header bytes.  Returns (attr, time, offset of the first byte after the end_of_head line)."""
    hdr = {}
    pos = 0
    end = len(data)
    body = end
    while pos < end:
        nl = data.find(b"\n", pos)
        nxt = end if nl < 0 else nl + 1
        ln = data[pos:nxt]
        pos = nxt
        if b"begin_of_head" in ln:
            continue
        if b"end_of_head" in ln:
            body = pos
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
        pass          # some values may be absent (icgem.py:78-80)
    time = None
    if "modelname" in hdr:
        try:
            time = [datetime.strptime(hdr["modelname"][-7:], "%Y-%m") + timedelta(days=14.5)]
        except ValueError:
            time = None
    return attr, time, body


def _first_record_columns(data, body, key):
    """number of whitespace-separated columns of the first record line (icgem.py:100-101: decides whether sigmas are read)"""
    pos = body
    while True:
        if data.startswith(key, pos):
            nl = data.find(b"\n", pos)
            return len(data[pos:len(data) if nl < 0 else nl].split())
        nl = data.find(b"\n", pos)
        if nl < 0:
            return -1
        pos = nl + 1


def parse_records(data, nmax, key=b"gfc", d_exponent=True, with_sigma=False, device=0, out=None, sig_out=None, stream=None):
    """Record block ``data`` (bytes) -> (cnm, sigcnm or None, present, stats) as torch CUDA tensors (shb200_gfc_parse).
    present[k] = number of records that wrote position k.  ``out`` / ``sig_out``: optional preallocated [nsh] float64 rows."""
    import torch
    dev = torch.device("cuda", device)
    nsh = (int(nmax) + 1) ** 2
    buf = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev) if len(data) else torch.zeros(1, dtype=torch.uint8, device=dev)
    cnm = torch.empty(nsh, dtype=torch.float64, device=dev) if out is None else out
    sig = (torch.empty(nsh, dtype=torch.float64, device=dev) if sig_out is None else sig_out) if with_sigma else None
    present = torch.empty(nsh, dtype=torch.int32, device=dev)
    stats = torch.empty(4, dtype=torch.int64, device=dev)
    cap = 4096
    fb = torch.empty(cap, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream if stream is None else stream
    capi.check(capi.lib().shb200_gfc_parse(buf.data_ptr(), len(data), key, int(bool(d_exponent)), int(nmax), int(bool(with_sigma)),
                                           cnm.data_ptr(), sig.data_ptr() if with_sigma else None, present.data_ptr(), stats.data_ptr(),
                                           fb.data_ptr(), cap, st))
    s = stats.cpu().numpy()
    nfb = int(s[2])
    if nfb:
        if nfb > cap:
            raise RuntimeError(f"{nfb} record lines could not be converted on the device (more than the hand-back list holds): not a plain ICGEM/GSM record block")
        # the host converts what the kernel handed back, with the reference's own calls (icgem.py:99-127)
        for off in sorted(fb[:nfb].cpu().tolist()):
            nl = data.find(b"\n", off)
            ln = data[off:len(data) if nl < 0 else nl]
            spl = (ln.replace(b"D", b"E") if d_exponent else ln).split()
            n = int(spl[1])
            if n > nmax:
                continue
            m = int(spl[2])
            if m < 0 or m > n:
                raise RuntimeError(f"record with order {m} outside 0..{n}: {ln!r}")
            k0 = n * n + n
            cnm[k0 + m] = float(spl[3])
            present[k0 + m] += 1
            if with_sigma:
                sig[k0 + m] = float(spl[5])
            if m != 0:
                cnm[k0 - m] = float(spl[4])
                present[k0 - m] += 1
                if with_sigma:
                    sig[k0 - m] = float(spl[6])
    return cnm, sig, present, {"records": int(s[0]) + nfb, "skipped_above_nmax": int(s[1]), "host_converted_lines": nfb}


def read_icgem(src, nmaxstop=sys.maxsize, device=0):
    """ICGEM .gfc(.gz) file -> dict(cnm, sigcnm, present: torch CUDA [nsh]; nm: (n, m) of the records present, canonical order;
    attrs; time; stats).  Same header rules, nmax rule and epoch hack as readIcgem (io/icgem.py:28-164)."""
    data = _read_bytes(src)
    attr, time, body = _icgem_header(data, nmaxstop)
    if "nmax" not in attr:
        raise RuntimeError("ICGEM header has no max_degree")
    nmax = attr["nmax"]
    ncol = _first_record_columns(data, body, b"gfc")
    cnm, sig, present, stats = parse_records(data[body:], nmax, key=b"gfc", d_exponent=True, with_sigma=ncol >= 6, device=device)
    return _result(cnm, sig, present, stats, attr, time, nmax)


def read_gsmv6(src, nmaxstop=sys.maxsize, device=0):
    """GRACE GSM RL06 file (YAML header + GRCOF2 records) -> same dict as read_icgem (io/gsmv6.py:16-107)."""
    import yaml
    data = _read_bytes(src)
    mark = data.find(b"# End of YAML header")
    if mark < 0:
        raise RuntimeError("no '# End of YAML header' line")
    nl = data.find(b"\n", mark)
    body = len(data) if nl < 0 else nl + 1
    line0 = data.rfind(b"\n", 0, mark) + 1
    hdr = yaml.safe_load(data[:line0])["header"]
    attr = {"nmaxfile": hdr["dimensions"]["degree"]}
    attr["nmax"] = attr["nmaxfile"]
    attr["tstart"] = hdr["global_attributes"]["time_coverage_start"]
    attr["tend"] = hdr["global_attributes"]["time_coverage_end"]
    nonstand = hdr["non-standard_attributes"]
    attr["gm"] = nonstand["earth_gravity_param"]["value"]
    attr["re"] = nonstand["mean_equator_radius"]["value"]
    time = [attr["tstart"] + (attr["tend"] - attr["tstart"]) / 2]
    nmax = min(attr["nmax"], nmaxstop)
    cnm, sig, present, stats = parse_records(data[body:], nmax, key=b"GRCOF2", d_exponent=False, with_sigma=True, device=device)
    return _result(cnm, sig, present, stats, attr, time, nmax)


def _result(cnm, sig, present, stats, attr, time, nmax):
    pres = present.cpu().numpy()
    if (pres > 1).any():
        k = int(np.argmax(pres > 1))
        n = int(np.sqrt(k))
        raise RuntimeError(f"coefficient (n={n}, m={k - n * n - n}) appears {int(pres[k])} times in the record block")
    from .engine import nm_index
    n, m = nm_index(nmax)
    keep = pres > 0
    return {"cnm": cnm, "sigcnm": sig, "present": present, "n": n[keep], "m": m[keep], "attrs": attr, "time": time, "stats": stats, "nmax": nmax}


def read_icgem_batch(sources, nmax, device=0):
    """A series of ICGEM files (monthly solutions) -> (coef [B, (nmax+1)^2] torch CUDA, list of per-file dicts without arrays).
    Files with a smaller max_degree are zero padded; coefficients above ``nmax`` are dropped (nmaxstop semantics)."""
    import torch
    dev = torch.device("cuda", device)
    nsh = (nmax + 1) ** 2
    coef = torch.zeros((len(sources), nsh), dtype=torch.float64, device=dev)
    infos = []
    for b, src in enumerate(sources):
        data = _read_bytes(src)
        attr, time, body = _icgem_header(data, nmax)
        # one parse into row b: a file with a smaller max_degree only writes its own leading block of the nm vector
        _, _, present, stats = parse_records(data[body:], nmax, key=b"gfc", d_exponent=True, with_sigma=False, device=device, out=coef[b])
        infos.append({"attrs": attr, "time": time, "stats": stats})
    return coef, infos


def scatter_nm(n, m, values, nmax, device=0):
    """(n, m, value) records in any order -> dense [B, (nmax+1)^2] torch CUDA array; values [B, nrec] or [nrec]."""
    import torch
    dev = torch.device("cuda", device)
    tn = torch.as_tensor(np.ascontiguousarray(n, dtype=np.int32), device=dev)
    tm = torch.as_tensor(np.ascontiguousarray(m, dtype=np.int32), device=dev)
    tv = values if hasattr(values, "is_cuda") else torch.as_tensor(np.asarray(values, dtype=np.float64), device=dev)
    squeeze = tv.ndim == 1
    if squeeze:
        tv = tv[None]
    B, nrec = tv.shape
    out = torch.zeros((B, (nmax + 1) ** 2), dtype=torch.float64, device=dev)
    capi.check(capi.lib().shb200_scatter_nm(tn.data_ptr(), tm.data_ptr(), tv.data_ptr(), tv.stride(0), tv.stride(1), nrec, B, int(nmax),
                                            out.data_ptr(), out.stride(0), out.stride(1), torch.cuda.current_stream(dev).cuda_stream))
    return out[0] if squeeze else out


def polygon_masks(polygons, lon, lat, device=0):
    """polygons: list of polygons; a polygon is a list of rings, a ring an [nv, 2] array of (lon, lat) vertices (exterior first,
    holes after it, further parts of a multi-polygon likewise; rings are closed implicitly).  Returns the 0/1 masks
    [npoly, nlat, nlon] as a torch CUDA tensor (geom/polygons.py:66-93: grid points contained in each polygon)."""
    import torch
    dev = torch.device("cuda", device)
    lon = np.ascontiguousarray(lon, dtype=np.float64)
    lat = np.ascontiguousarray(lat, dtype=np.float64)
    vx, vy, ring_start, poly_ring = [], [], [0], [0]
    for rings in polygons:
        for ring in rings:
            r = np.asarray(ring, dtype=np.float64)
            if r.ndim != 2 or r.shape[1] != 2 or len(r) < 3:
                raise ValueError("a ring needs at least 3 (lon, lat) vertices")
            if np.array_equal(r[0], r[-1]):
                r = r[:-1]                      # explicit closing vertex
            vx.append(r[:, 0])
            vy.append(r[:, 1])
            ring_start.append(ring_start[-1] + len(r))
        poly_ring.append(len(ring_start) - 1)
    if not polygons:
        return torch.zeros((0, len(lat), len(lon)), dtype=torch.float64, device=dev)
    tvx = torch.as_tensor(np.concatenate(vx), device=dev)
    tvy = torch.as_tensor(np.concatenate(vy), device=dev)
    trs = torch.as_tensor(np.asarray(ring_start, dtype=np.int32), device=dev)
    tpr = torch.as_tensor(np.asarray(poly_ring, dtype=np.int32), device=dev)
    tlon = torch.as_tensor(lon, device=dev)
    tlat = torch.as_tensor(lat, device=dev)
    wrap = 0 if lon.min() < 0 else 1            # polygons.py:76-81
    mask = torch.empty((len(polygons), len(lat), len(lon)), dtype=torch.float64, device=dev)
    capi.check(capi.lib().shb200_polygon_mask(tvx.data_ptr(), tvy.data_ptr(), trs.data_ptr(), tpr.data_ptr(), len(polygons), tlon.data_ptr(),
                                              tlat.data_ptr(), len(lat), len(lon), wrap, mask.data_ptr(), mask.stride(0), mask.stride(1),
                                              mask.stride(2), torch.cuda.current_stream(dev).cuda_stream))
    return mask


def polygon2sh(engine, polygons, nmax, gtype="regular"):
    """geom/polygons.py:14-97 on the device: masks on ``lonlat_grid(nmax)`` -> SH analysis.  Returns (coef [npoly, nsh] torch
    CUDA, grid dict).  ``engine``: an SHEngine."""
    from .engine import lonlat_grid
    g = lonlat_grid(nmax, gtype=gtype)
    mask = polygon_masks(polygons, g["lon"], g["lat"], device=engine.device)
    q = "gauss" if gtype == "gauss" else "shlib"
    return engine.analysis(mask, nmax, g["lon"], g["lat"], quadrature=q), g
