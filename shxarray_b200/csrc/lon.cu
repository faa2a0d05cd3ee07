// Longitude stage, dense path: direct trig sums for arbitrary longitudes and point sets, with trig(m*lonr)
// evaluated like Ynm.hpp:114,127.  The FFT path lives in lon_fft.cu; this file also holds the stage dispatch.
#include "internal.h"

namespace shb {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
// multiply by -i*s  (s=+1: forward sign, s=-1: inverse)
__device__ __forceinline__ double2 mul_mi(double2 a, double s) { return make_double2(s * a.y, -s * a.x); }

// ------------------------------------------------------------------------------------------------
// Dense trig sums.  points != 0: row r has the single longitude lonr[i1[r]] (point mode).
// thread <-> (row, longitude index); one group of GS fields (their cosine and sine columns are adjacent in G: 2*GS contiguous
// doubles per parity) at a time with the sums of the group in registers, orders innermost: sincos(m*lonr) is evaluated once
// per (point, order, group) instead of once per field (r01), in the reference's order of operations (Ynm.hpp:114,127), and
// the order sum of every field runs m = 0, 1, ... exactly as before (bit-identical results).
template <int GS>
__global__ void __launch_bounds__(128) k_lon_syn_dense(DevTables t, int batch, int points, const double* __restrict__ G,
                                                       double* __restrict__ grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon) {
    const int N = t.nmax, R = t.nrows;
    const int nl = points ? 1 : t.nlon;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)R * nl) return;
    const int row = (int)(gid / nl);
    const int j = (int)(gid - (int64_t)row * nl);
    const int i1 = t.row_i1[row], i2 = t.row_i2[row];
    const double lonr = points ? t.lonr[i1] : t.lonr[j];
    const int64_t gstep = (int64_t)R * 2 * t.Gs;
    for (int b0 = 0; b0 < batch; b0 += GS) {
        double s1[GS], s2[GS];
#pragma unroll
        for (int f = 0; f < GS; ++f) { s1[f] = 0.0; s2[f] = 0.0; }
        const double* g = G + ((int64_t)row * 2) * t.Gs + colidx(0, b0, GS);
        for (int m = 0; m <= N; ++m, g += gstep) {
            double sn, cs;
            sincos((double)m * lonr, &sn, &cs);   // Ynm.hpp:127
#pragma unroll
            for (int f = 0; f < GS; ++f) {
                const double Ec = g[f], Es = g[GS + f], Oc = g[t.Gs + f], Os = g[t.Gs + GS + f];
                s1[f] += (Ec + Oc) * cs + (Es + Os) * sn;
                s2[f] += (Ec - Oc) * cs + (Es - Os) * sn;
            }
        }
#pragma unroll
        for (int f = 0; f < GS; ++f) {
            if (b0 + f >= batch) break;
            const int64_t off = (int64_t)(b0 + f) * gs_b + (points ? 0 : (int64_t)j * gs_lon);
            grid[off + (int64_t)i1 * gs_lat] = s1[f];
            if (i2 >= 0) grid[off + (int64_t)i2 * gs_lat] = s2[f];
        }
    }
}

// thread <-> (row, order m); one group of GS fields at a time, longitudes innermost (sequential, deterministic; same sums as r01)
template <int GS>
__global__ void __launch_bounds__(128) k_lon_ana_dense(DevTables t, int batch, const double* __restrict__ grid, int64_t gs_b,
                                                       int64_t gs_lat, int64_t gs_lon, double* __restrict__ G) {
    const int N = t.nmax, R = t.nrows, Bp = t.Bp;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)R * (N + 1)) return;
    const int row = (int)(gid / (N + 1));
    const int m = (int)(gid - (int64_t)row * (N + 1));
    const int i1 = t.row_i1[row], i2 = t.row_i2[row];
    const double w1 = t.w1[row], w2 = t.w2[row];
    for (int b0 = 0; b0 < Bp; b0 += GS) {
        double a[GS], bb[GS], ap[GS], bp[GS];
#pragma unroll
        for (int f = 0; f < GS; ++f) { a[f] = 0.0; bb[f] = 0.0; ap[f] = 0.0; bp[f] = 0.0; }
        const int nf = min(GS, batch - b0);            // live fields of the group (<= 0: padded group, zeros)
        if (nf > 0) {
            const double* r1 = grid + (int64_t)b0 * gs_b + (int64_t)i1 * gs_lat;
            const double* r2 = grid + (int64_t)b0 * gs_b + (int64_t)(i2 >= 0 ? i2 : i1) * gs_lat;
            for (int j = 0; j < t.nlon; ++j) {
                double sn, cs;
                sincos((double)m * t.lonr[j], &sn, &cs);
#pragma unroll
                for (int f = 0; f < GS; ++f) {
                    if (f < nf) {
                        const double f1 = r1[(int64_t)f * gs_b + (int64_t)j * gs_lon];
                        const double f2 = i2 >= 0 ? r2[(int64_t)f * gs_b + (int64_t)j * gs_lon] : 0.0;
                        a[f] += f1 * cs; bb[f] += f1 * sn; ap[f] += f2 * cs; bp[f] += f2 * sn;
                    }
                }
            }
        }
        double* g = G + (((int64_t)m * R + row) * 2) * t.Gs + colidx(0, b0, GS);
#pragma unroll
        for (int f = 0; f < GS; ++f) {
            const double A = a[f] * w1, B = bb[f] * w1, Ap = ap[f] * w2, Bq = bp[f] * w2;
            g[f] = A + Ap; g[GS + f] = B + Bq; g[t.Gs + f] = A - Ap; g[t.Gs + GS + f] = B - Bq;
        }
    }
}

// ------------------------------------------------------------------------------------------------
int launch_lon_syn_fft(const Plan& p, int batch, double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st);
int launch_lon_ana_fft(const Plan& p, int batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st);
int launch_lon_syn_fft3(const Plan& p, int batch, double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st);
int launch_lon_ana_fft3(const Plan& p, int batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st);

int launch_lon_syn(const Plan& p, int batch, double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st) {
    if (p.lon_fft) {
        const int rc3 = launch_lon_syn_fft3(p, batch, grid, gs_b, gs_lat, gs_lon, st);
        return rc3 >= 0 ? rc3 : launch_lon_syn_fft(p, batch, grid, gs_b, gs_lat, gs_lon, st);
    }
    int points = p.mode == SHB200_MODE_POINTS;
    int nl = points ? 1 : p.nlon;
    int64_t tot = (int64_t)p.nrows * nl;
    dim3 g((unsigned)((tot + 127) / 128));
    trace_kernel(p, "k_lon_syn_dense");
    if (p.t.GS == 8) k_lon_syn_dense<8><<<g, 128, 0, st>>>(p.t, batch, points, p.G, grid, gs_b, gs_lat, gs_lon);
    else k_lon_syn_dense<4><<<g, 128, 0, st>>>(p.t, batch, points, p.G, grid, gs_b, gs_lat, gs_lon);
    return 0;
}

int launch_lon_ana(const Plan& p, int batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st) {
    if (p.lon_fft) {
        const int rc3 = launch_lon_ana_fft3(p, batch, grid, gs_b, gs_lat, gs_lon, st);
        return rc3 >= 0 ? rc3 : launch_lon_ana_fft(p, batch, grid, gs_b, gs_lat, gs_lon, st);
    }
    int64_t tot = (int64_t)p.nrows * (p.nmax + 1);
    dim3 g((unsigned)((tot + 127) / 128));
    trace_kernel(p, "k_lon_ana_dense");
    if (p.t.GS == 8) k_lon_ana_dense<8><<<g, 128, 0, st>>>(p.t, batch, grid, gs_b, gs_lat, gs_lon, p.G);
    else k_lon_ana_dense<4><<<g, 128, 0, st>>>(p.t, batch, grid, gs_b, gs_lat, gs_lon, p.G);
    return 0;
}

}  // namespace shb
