// Data formats either side of the transform, on the device:
//
//  * shb200_gfc_parse   -- the record block of an ICGEM / GSM coefficient file ("gfc n m C S [sigC sigS]" lines,
//                          io/icgem.py:94-128 of the reference, which walks the file line by line in Python: minutes for a
//                          5' model) parsed by one kernel straight into the dense nm vector the engine consumes
//                          (position n^2+n+m, sh_indexing.py:66).  Decimal -> binary64 conversion is CORRECTLY ROUNDED
//                          (bit-identical to Python's float()): Eisel-Lemire with a 128-bit power-of-five table
//                          (pow5_table.h); anything the fast path cannot decide, or any token outside the plain
//                          [sign] digits [. digits] [E|D|e [sign] digits] grammar, is reported back by byte offset and
//                          converted by the host with the reference's own float().
//  * shb200_polygon_mask -- grid points inside polygons (even-odd rule over all rings of a polygon, planar lon/lat like the
//                          reference's geopandas "contains" query, geom/polygons.py:47-93) -> 0/1 mask fields [npoly][nlat][nlon]
//                          ready for shb200_analysis; longitudes are wrapped to [-180, 180) when the grid is 0..360 (:76-81).
//  * shb200_scatter_nm   -- (n, m, value) triplets in any order -> dense nm vectors (batched: one record table per field).
#include <cstring>

#include "internal.h"
#include "pow5_table.h"

namespace shb {

// ---- decimal -> double ------------------------------------------------------------------------------------------------
// w * 10^q, w < 2^64.  Returns false when the result cannot be decided here (the caller falls back to the host).
__device__ inline bool decimal_to_double(uint64_t w, int q, bool neg, double& out) {
    const uint64_t sign = neg ? 0x8000000000000000ull : 0ull;
    if (w == 0 || q < POW5_QMIN) { out = __longlong_as_double((long long)sign); return true; }
    if (q > POW5_QMAX) { out = __longlong_as_double((long long)(sign | 0x7FF0000000000000ull)); return true; }
    const int lz = __clzll((long long)w);
    w <<= lz;
    const uint64_t t_hi = POW5_128[q - POW5_QMIN][0], t_lo = POW5_128[q - POW5_QMIN][1];
    uint64_t upper = __umul64hi(w, t_hi), lower = w * t_hi;
    if ((upper & 0x1FF) == 0x1FF) {                       // the truncated product is too close to a rounding boundary: refine
        const uint64_t sh = __umul64hi(w, t_lo);
        lower += sh;
        if (sh > lower) ++upper;
    }
    if (lower == 0xFFFFFFFFFFFFFFFFull && !(q >= -27 && q <= 55)) return false;
    const int upperbit = (int)(upper >> 63);
    uint64_t mant = upper >> (upperbit + 9);
    int power2 = (((152170 + 65536) * q) >> 16) + 63 + upperbit - lz + 1023;
    if (power2 <= 0) {                                    // subnormal
        if (-power2 + 1 >= 64) { out = __longlong_as_double((long long)sign); return true; }
        mant >>= -power2 + 1;
        mant += mant & 1;
        mant >>= 1;
        out = __longlong_as_double((long long)(sign | mant));          // mant may carry into the exponent field: that is the value 2^-1022
        return true;
    }
    if (lower <= 1 && q >= -4 && q <= 23 && (mant & 3) == 1) {          // exactly halfway: round to even
        if ((mant << (upperbit + 9)) == upper) mant &= ~1ull;
    }
    mant += mant & 1;
    mant >>= 1;
    if (mant >= (2ull << 52)) { mant = 1ull << 52; ++power2; }
    mant &= ~(1ull << 52);
    if (power2 >= 0x7FF) { out = __longlong_as_double((long long)(sign | 0x7FF0000000000000ull)); return true; }
    out = __longlong_as_double((long long)(sign | ((uint64_t)power2 << 52) | mant));
    return true;
}

__device__ inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\x0b' || c == '\x0c'; }   // bytes.split() minus '\n' (line end)

// token [s, e): plain decimal floating-point literal?  ('D' exponent = the reference's ln.replace(b"D", b"E"))
__device__ inline bool parse_float_token(const char* s, const char* e, double& out, bool allow_d = true) {
    if (s >= e) return false;
    bool neg = false;
    if (*s == '+' || *s == '-') { neg = *s == '-'; ++s; }
    uint64_t w = 0;
    int nd = 0, q = 0, ndig = 0;
    bool seen_point = false;
    for (; s < e; ++s) {
        const char c = *s;
        if (c >= '0' && c <= '9') {
            ++ndig;
            if (nd == 0 && c == '0') { if (seen_point) --q; continue; }      // leading zeros carry no digits
            if (nd >= 19) return false;                                       // more digits than a uint64 holds: host
            w = w * 10 + (uint64_t)(c - '0');
            ++nd;
            if (seen_point) --q;
        } else if (c == '.' && !seen_point) {
            seen_point = true;
        } else {
            break;
        }
    }
    if (ndig == 0) return false;
    if (s < e) {
        if (!(*s == 'E' || *s == 'e' || (allow_d && *s == 'D'))) return false;
        ++s;
        bool eneg = false;
        if (s < e && (*s == '+' || *s == '-')) { eneg = *s == '-'; ++s; }
        if (s >= e) return false;
        int ex = 0;
        for (; s < e; ++s) {
            if (*s < '0' || *s > '9') return false;
            if (ex < 100000) ex = ex * 10 + (*s - '0');
        }
        q += eneg ? -ex : ex;
    }
    return decimal_to_double(w, q, neg, out);
}

__device__ inline bool parse_int_token(const char* s, const char* e, int& out) {
    if (s >= e) return false;
    bool neg = false;
    if (*s == '+' || *s == '-') { neg = *s == '-'; ++s; }
    if (s >= e || e - s > 9) return false;
    int v = 0;
    for (; s < e; ++s) {
        if (*s < '0' || *s > '9') return false;
        v = v * 10 + (*s - '0');
    }
    out = neg ? -v : v;
    return true;
}

struct RecordKey { char c[8]; int len; int allow_d; };

// One thread per byte; the thread that sits on the first byte of a record line parses it (icgem.py:94-128, gsmv6.py:66-90).
__global__ void __launch_bounds__(256) k_gfc_parse(const char* __restrict__ text, int64_t nbytes, RecordKey key, int nmax, int with_sigma, double* __restrict__ cnm,
                                                   double* __restrict__ sig, int32_t* __restrict__ present, unsigned long long* __restrict__ stats,
                                                   int64_t* __restrict__ fallback, int cap) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i + key.len > nbytes) return;
    for (int j = 0; j < key.len; ++j)
        if (text[i + j] != key.c[j]) return;                                          // re.compile(b'^gfc').match(ln)
    if (i > 0 && text[i - 1] != '\n') return;
    const char* p = text + i;
    const char* end = text + nbytes;
    const char* ts[7];
    const char* te[7];
    int nt = 0;
    while (p < end && *p != '\n' && nt < 7) {
        while (p < end && is_ws(*p)) ++p;
        if (p >= end || *p == '\n') break;
        ts[nt] = p;
        while (p < end && *p != '\n' && !is_ws(*p)) ++p;
        te[nt++] = p;
    }
    auto give_up = [&]() {
        const unsigned long long k = atomicAdd(&stats[2], 1ull);
        if (k < (unsigned long long)cap) fallback[k] = i;
    };
    int n, m;
    if (nt < 2 || !parse_int_token(ts[1], te[1], n)) { give_up(); return; }
    if (n > nmax) { atomicAdd(&stats[1], 1ull); return; }                             // icgem.py:105-111
    if (nt < 4 || !parse_int_token(ts[2], te[2], m) || n < 0 || m < 0 || m > n) { give_up(); return; }
    const int need = with_sigma ? (m != 0 ? 7 : 6) : (m != 0 ? 5 : 4);
    if (nt < need) { give_up(); return; }
    double c, s = 0.0, sc = 0.0, ss = 0.0;
    const bool ad = key.allow_d != 0;
    bool ok = parse_float_token(ts[3], te[3], c, ad);
    if (m != 0) ok = ok && parse_float_token(ts[4], te[4], s, ad);
    if (with_sigma) {
        ok = ok && parse_float_token(ts[5], te[5], sc, ad);
        if (m != 0) ok = ok && parse_float_token(ts[6], te[6], ss, ad);
    }
    if (!ok) { give_up(); return; }
    const int64_t k0 = (int64_t)n * n + n;
    cnm[k0 + m] = c;
    if (with_sigma) sig[k0 + m] = sc;
    if (present) atomicAdd(&present[k0 + m], 1);
    if (m != 0) {
        cnm[k0 - m] = s;
        if (with_sigma) sig[k0 - m] = ss;
        if (present) atomicAdd(&present[k0 - m], 1);
    }
    atomicAdd(&stats[0], 1ull);
}

// standalone conversion of newline-separated tokens (tests: the converter against float() on arbitrary literals)
__global__ void __launch_bounds__(256) k_parse_tokens(const char* __restrict__ text, const int64_t* __restrict__ start, const int64_t* __restrict__ stop, int64_t n,
                                                      double* __restrict__ out, uint8_t* __restrict__ ok) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = 0.0;
    ok[i] = parse_float_token(text + start[i], text + stop[i], v) ? 1 : 0;
    out[i] = v;
}

// ---- (n, m, value) triplets -> dense nm vectors -----------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scatter_nm(const int32_t* __restrict__ n, const int32_t* __restrict__ m, const double* __restrict__ val, int64_t vs_b,
                                                    int64_t vs_k, int64_t nrec, int batch, int nmax, double* __restrict__ coef, int64_t cs_b, int64_t cs_nm) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nrec) return;
    const int nn = n[k], mm = m[k];
    if (nn < 0 || nn > nmax || mm > nn || mm < -nn) return;
    const int64_t pos = (int64_t)nn * nn + nn + mm;
    for (int b = 0; b < batch; ++b) coef[b * cs_b + pos * cs_nm] = val[b * vs_b + k * vs_k];
}

// ---- polygons -> grid masks -----------------------------------------------------------------------------------------------
// vertices of all rings of all polygons in one array (x = lon, y = lat), rings closed implicitly; ring_start[r] .. ring_start[r+1]
// are the vertices of ring r, poly_ring[p] .. poly_ring[p+1] the rings of polygon p.  Even-odd rule over every ring of the polygon
// (holes and multi-part polygons).  Block = 256 consecutive grid points of one polygon; edges are staged through shared memory.
constexpr int PM_EDGES = 512;
__global__ void __launch_bounds__(256) k_polygon_mask(const double* __restrict__ vx, const double* __restrict__ vy, const int32_t* __restrict__ ring_start,
                                                      const int32_t* __restrict__ poly_ring, const double* __restrict__ lon, const double* __restrict__ lat,
                                                      int nlat, int nlon, int wrap, double* __restrict__ mask, int64_t ms_p, int64_t ms_lat, int64_t ms_lon) {
    __shared__ double ex0[PM_EDGES], ey0[PM_EDGES], ex1[PM_EDGES], ey1[PM_EDGES];
    const int p = blockIdx.y;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = idx < (int64_t)nlat * nlon;
    const int il = live ? (int)(idx / nlon) : 0, ik = live ? (int)(idx - (int64_t)il * nlon) : 0;
    double x = lon[ik];
    const double y = lat[il];
    if (wrap) x = fmod(x + 180.0, 360.0) - 180.0;          // (lon+180)%360-180  (polygons.py:80)
    bool inside = false;
    for (int r = poly_ring[p]; r < poly_ring[p + 1]; ++r) {
        const int v0 = ring_start[r], nv = ring_start[r + 1] - v0;
        for (int e0 = 0; e0 < nv; e0 += PM_EDGES) {
            const int ne = min(PM_EDGES, nv - e0);
            __syncthreads();
            for (int j = threadIdx.x; j < ne; j += blockDim.x) {
                const int a = v0 + e0 + j, b = v0 + (e0 + j + 1) % nv;
                ex0[j] = vx[a]; ey0[j] = vy[a]; ex1[j] = vx[b]; ey1[j] = vy[b];
            }
            __syncthreads();
            for (int j = 0; j < ne; ++j) {
                const double y0 = ey0[j], y1 = ey1[j];
                if ((y0 > y) != (y1 > y)) {
                    // crossing of the edge with the horizontal through the point, evaluated with explicitly rounded operations
                    const double xc = __dadd_rn(__dmul_rn(__ddiv_rn(__dsub_rn(ex1[j], ex0[j]), __dsub_rn(y1, y0)), __dsub_rn(y, y0)), ex0[j]);
                    if (x < xc) inside = !inside;
                }
            }
        }
    }
    if (live) mask[p * ms_p + il * ms_lat + ik * ms_lon] = inside ? 1.0 : 0.0;
}

}  // namespace shb

using namespace shb;

extern "C" {

int shb200_gfc_parse(const char* text, int64_t nbytes, const char* record_key, int32_t d_exponent, int32_t nmax, int32_t with_sigma, double* cnm, double* sigcnm, int32_t* present,
                     uint64_t* stats, int64_t* fallback_offsets, int32_t fallback_capacity, void* stream) {
    if (!text || nbytes < 0 || nmax < 0 || !cnm || !stats || (with_sigma && !sigcnm) || (fallback_capacity > 0 && !fallback_offsets)) {
        set_error("gfc_parse: invalid argument");
        return SHB200_EINVAL;
    }
    RecordKey key{};
    key.len = record_key ? (int)strnlen(record_key, 9) : 0;
    if (key.len < 1 || key.len > 8) { set_error("gfc_parse: record key must have 1..8 characters"); return SHB200_EINVAL; }
    memcpy(key.c, record_key, key.len);
    key.allow_d = d_exponent;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nsh = (int64_t)(nmax + 1) * (nmax + 1);
    SHB_CUDA(cudaMemsetAsync(cnm, 0, sizeof(double) * nsh, st));
    if (with_sigma) SHB_CUDA(cudaMemsetAsync(sigcnm, 0, sizeof(double) * nsh, st));
    if (present) SHB_CUDA(cudaMemsetAsync(present, 0, sizeof(int32_t) * nsh, st));
    SHB_CUDA(cudaMemsetAsync(stats, 0, sizeof(uint64_t) * 4, st));
    if (nbytes >= key.len)
        k_gfc_parse<<<(unsigned)((nbytes + 255) / 256), 256, 0, st>>>(text, nbytes, key, nmax, with_sigma, cnm, sigcnm, present,
                                                                      reinterpret_cast<unsigned long long*>(stats), fallback_offsets, fallback_capacity);
    SHB_CUDA(cudaGetLastError());
    return 0;
}

int shb200_parse_floats(const char* text, const int64_t* start, const int64_t* stop, int64_t ntokens, double* out, uint8_t* ok, void* stream) {
    if (!text || !start || !stop || !out || !ok || ntokens < 0) { set_error("parse_floats: invalid argument"); return SHB200_EINVAL; }
    if (ntokens) k_parse_tokens<<<(unsigned)((ntokens + 255) / 256), 256, 0, (cudaStream_t)stream>>>(text, start, stop, ntokens, out, ok);
    SHB_CUDA(cudaGetLastError());
    return 0;
}

int shb200_scatter_nm(const int32_t* n, const int32_t* m, const double* values, int64_t vs_b, int64_t vs_k, int64_t nrecords, int32_t batch,
                      int32_t nmax, double* coef, int64_t cs_b, int64_t cs_nm, void* stream) {
    if (!n || !m || !values || !coef || nrecords < 0 || batch < 1 || nmax < 0) { set_error("scatter_nm: invalid argument"); return SHB200_EINVAL; }
    if (nrecords) k_scatter_nm<<<(unsigned)((nrecords + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, m, values, vs_b, vs_k, nrecords, batch, nmax, coef, cs_b, cs_nm);
    SHB_CUDA(cudaGetLastError());
    return 0;
}

int shb200_polygon_mask(const double* vx, const double* vy, const int32_t* ring_start, const int32_t* poly_ring, int32_t npoly,
                        const double* lon, const double* lat, int32_t nlat, int32_t nlon, int32_t wrap_lon,
                        double* mask, int64_t ms_p, int64_t ms_lat, int64_t ms_lon, void* stream) {
    if (!vx || !vy || !ring_start || !poly_ring || !lon || !lat || !mask || npoly < 1 || nlat < 1 || nlon < 1) { set_error("polygon_mask: invalid argument"); return SHB200_EINVAL; }
    dim3 grid((unsigned)(((int64_t)nlat * nlon + 255) / 256), (unsigned)npoly);
    k_polygon_mask<<<grid, 256, 0, (cudaStream_t)stream>>>(vx, vy, ring_start, poly_ring, lon, lat, nlat, nlon, wrap_lon, mask, ms_p, ms_lat, ms_lon);
    SHB_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
