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
// thread <-> (row, longitude index); loops orders; fields innermost.
__global__ void __launch_bounds__(128) k_lon_syn_dense(DevTables t, int batch, int points, const double* __restrict__ G,
                                                       double* __restrict__ grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon) {
    const int N = t.nmax, R = t.nrows, NC = t.NC, Bp = t.Bp;
    const int nl = points ? 1 : t.nlon;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)R * nl) return;
    const int row = (int)(gid / nl);
    const int j = (int)(gid - (int64_t)row * nl);
    const int i1 = t.row_i1[row], i2 = t.row_i2[row];
    const double lonr = points ? t.lonr[i1] : t.lonr[j];
    for (int b = 0; b < batch; ++b) {
        double s1 = 0.0, s2 = 0.0;
        for (int m = 0; m <= N; ++m) {
            double sn, cs;
            sincos((double)m * lonr, &sn, &cs);   // Ynm.hpp:127
            const double* g = G + (((int64_t)m * R + row) * 2) * t.Gs;
            const int cc = colidx(0, b, t.GS), sc = colidx(1, b, t.GS);
            double Ec = g[cc], Es = g[sc], Oc = g[t.Gs + cc], Os = g[t.Gs + sc];
            s1 += (Ec + Oc) * cs + (Es + Os) * sn;
            s2 += (Ec - Oc) * cs + (Es - Os) * sn;
        }
        int64_t off = (int64_t)b * gs_b + (points ? 0 : (int64_t)j * gs_lon);
        grid[off + (int64_t)i1 * gs_lat] = s1;
        if (i2 >= 0) grid[off + (int64_t)i2 * gs_lat] = s2;
    }
}

// thread <-> (row, order m); loops longitudes; fields innermost (sequential, deterministic)
__global__ void __launch_bounds__(128) k_lon_ana_dense(DevTables t, int batch, const double* __restrict__ grid, int64_t gs_b,
                                                       int64_t gs_lat, int64_t gs_lon, double* __restrict__ G) {
    const int N = t.nmax, R = t.nrows, NC = t.NC, Bp = t.Bp;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)R * (N + 1)) return;
    const int row = (int)(gid / (N + 1));
    const int m = (int)(gid - (int64_t)row * (N + 1));
    const int i1 = t.row_i1[row], i2 = t.row_i2[row];
    const double w1 = t.w1[row], w2 = t.w2[row];
    for (int b = 0; b < Bp; ++b) {
        double a = 0, bb = 0, ap = 0, bp = 0;
        if (b < batch) {
            for (int j = 0; j < t.nlon; ++j) {
                double sn, cs;
                sincos((double)m * t.lonr[j], &sn, &cs);
                double f1 = grid[(int64_t)b * gs_b + (int64_t)i1 * gs_lat + (int64_t)j * gs_lon];
                double f2 = i2 >= 0 ? grid[(int64_t)b * gs_b + (int64_t)i2 * gs_lat + (int64_t)j * gs_lon] : 0.0;
                a += f1 * cs; bb += f1 * sn; ap += f2 * cs; bp += f2 * sn;
            }
            a *= w1; bb *= w1; ap *= w2; bp *= w2;
        }
        double* g = G + (((int64_t)m * R + row) * 2) * t.Gs;
        const int cc = colidx(0, b, t.GS), sc = colidx(1, b, t.GS);
        g[cc] = a + ap; g[sc] = bb + bp; g[t.Gs + cc] = a - ap; g[t.Gs + sc] = bb - bp;
    }
}

// ------------------------------------------------------------------------------------------------
int launch_lon_syn_fft(const Plan& p, int batch, double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st);
int launch_lon_ana_fft(const Plan& p, int batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st);

int launch_lon_syn(const Plan& p, int batch, double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st) {
    if (p.lon_fft) return launch_lon_syn_fft(p, batch, grid, gs_b, gs_lat, gs_lon, st);
    int points = p.mode == SHB200_MODE_POINTS;
    int nl = points ? 1 : p.nlon;
    int64_t tot = (int64_t)p.nrows * nl;
    dim3 g((unsigned)((tot + 127) / 128));
    trace_kernel(p, "k_lon_syn_dense");
    k_lon_syn_dense<<<g, 128, 0, st>>>(p.t, batch, points, p.G, grid, gs_b, gs_lat, gs_lon);
    return 0;
}

int launch_lon_ana(const Plan& p, int batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st) {
    if (p.lon_fft) return launch_lon_ana_fft(p, batch, grid, gs_b, gs_lat, gs_lon, st);
    int64_t tot = (int64_t)p.nrows * (p.nmax + 1);
    dim3 g((unsigned)((tot + 127) / 128));
    trace_kernel(p, "k_lon_ana_dense");
    k_lon_ana_dense<<<g, 128, 0, st>>>(p.t, batch, grid, gs_b, gs_lat, gs_lon, p.G);
    return 0;
}

}  // namespace shb
