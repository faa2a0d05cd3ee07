// Longitude stage: batched complex FFT (two real rows per transform) and dense trig sums.
//
// What it replaces in the reference: the per-point trig factors cos(m*lon), sin(|m|*lon) of
// Ynm_cpp::set (Ynm.hpp:114,121-131) folded into the dgemv/dger of synthesis.pyx:183 / analysis.pyx:136.
//
// FFT path (periodic, equispaced longitudes lon_j = lon0 + j*360/K):
//   synthesis: rows i (primary) and i' (mirror) of one latitude pair are the real and imaginary part of ONE
//   complex inverse FFT of length K ("two-for-one"); the Hermitian spectra are assembled from
//   G[m][row][E/O][cos/sin][b] with the phase twist e^{i m lon0}; orders m >= K/2 alias onto k = +-m mod K
//   exactly as cos(m*lon_j) does on the grid.
//   analysis: forward FFT of f_i + i f_i', split, de-twist, quadrature weights, E/O combination.
// Stockham autosort, radices {4,2,3,5}, ping-pong in shared memory, several transforms per CTA.
//
// Dense path (any longitudes, point sets): direct sums with trig(m*lonr) evaluated like Ynm.hpp:127.
#include "internal.h"

namespace shb {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
// multiply by -i*s  (s=+1: forward sign, s=-1: inverse)
__device__ __forceinline__ double2 mul_mi(double2 a, double s) { return make_double2(s * a.y, -s * a.x); }

struct FftParams {
    int K;
    int nrad;
    int rad[24];
};

// One Stockham pass of radix R over `nt` transforms held in src[nt][K] -> dst[nt][K].
// sgn = -1 forward (e^{-i}), +1 inverse.  tw[k] = exp(-2 pi i k/K).
template <int R>
__device__ __forceinline__ void stockham_pass(const double2* __restrict__ src, double2* __restrict__ dst, int K, int Ns, int nt, const double2* __restrict__ tw, double sgn) {
    const int per = K / R;
    const int tstride = K / (Ns * R);
    for (int idx = threadIdx.x; idx < nt * per; idx += blockDim.x) {
        int tr = idx / per, j = idx - tr * per;
        const double2* s = src + (size_t)tr * K;
        double2* d = dst + (size_t)tr * K;
        int k = j % Ns;
        double2 v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            v[q] = s[j + q * per];
            if (q > 0 && k > 0) {
                double2 w = tw[(q * k * tstride)];
                if (sgn > 0) w.y = -w.y;
                v[q] = cmul(v[q], w);
            }
        }
        if (R == 2) {
            double2 a = v[0], b = v[1];
            v[0] = cadd(a, b); v[1] = csub(a, b);
        } else if (R == 4) {
            double2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]), c = cadd(v[1], v[3]), e = mul_mi(csub(v[1], v[3]), -sgn);
            // forward: X1 = b - i(v1-v3);  e = (-i*(-sgn))... keep explicit below
            v[0] = cadd(a, c); v[2] = csub(a, c);
            // e = mul_mi(x, -sgn) = (+i*sgn... ) -> for sgn=-1 (forward): mul_mi(x, +1) = -i x
            v[1] = cadd(b, e); v[3] = csub(b, e);
        } else if (R == 3) {
            const double c3 = -0.5, s3 = 0.86602540378443864676;  // cos, sin of 2pi/3
            double2 t1 = cadd(v[1], v[2]);
            double2 t2 = make_double2(v[0].x + c3 * t1.x, v[0].y + c3 * t1.y);
            double2 t3 = csub(v[1], v[2]);
            // forward: X1 = t2 - i s3 t3 ; inverse: X1 = t2 + i s3 t3
            double2 u = make_double2(-sgn * s3 * t3.y, sgn * s3 * t3.x);   // = i*sgn*s3*t3
            v[0] = cadd(v[0], t1);
            v[1] = cadd(t2, u);
            v[2] = csub(t2, u);
        } else if (R == 5) {
            const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;   // cos 2pi/5, cos 4pi/5
            const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;    // sin 2pi/5, sin 4pi/5
            double2 a1 = cadd(v[1], v[4]), b1 = csub(v[1], v[4]);
            double2 a2 = cadd(v[2], v[3]), b2 = csub(v[2], v[3]);
            double2 m1 = make_double2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
            double2 m2 = make_double2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
            double2 n1 = make_double2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
            double2 n2 = make_double2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
            // X_q = m + i*sgn*n  for q=1,2 ; minus for q=4,3
            double2 u1 = make_double2(-sgn * n1.y, sgn * n1.x), u2 = make_double2(-sgn * n2.y, sgn * n2.x);
            v[0] = cadd(v[0], cadd(a1, a2));
            v[1] = cadd(m1, u1); v[4] = csub(m1, u1);
            v[2] = cadd(m2, u2); v[3] = csub(m2, u2);
        }
        int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
        for (int q = 0; q < R; ++q) d[j0 + q * Ns] = v[q];
    }
}

// runs all passes; returns pointer to the buffer holding the result
__device__ double2* run_fft(double2* a, double2* b, const FftParams& fp, int nt, const double2* __restrict__ tw, double sgn) {
    int Ns = 1;
    double2 *src = a, *dst = b;
    for (int s = 0; s < fp.nrad; ++s) {
        int R = fp.rad[s];
        if (R == 4) stockham_pass<4>(src, dst, fp.K, Ns, nt, tw, sgn);
        else if (R == 2) stockham_pass<2>(src, dst, fp.K, Ns, nt, tw, sgn);
        else if (R == 3) stockham_pass<3>(src, dst, fp.K, Ns, nt, tw, sgn);
        else stockham_pass<5>(src, dst, fp.K, Ns, nt, tw, sgn);
        Ns *= R;
        __syncthreads();
        double2* tmp = src; src = dst; dst = tmp;
    }
    return src;
}

// ------------------------------------------------------------------------------------------------
// Synthesis: G -> grid.  CTA = (row, group of nt consecutive fields).
__global__ void __launch_bounds__(256) k_lon_syn_fft(DevTables t, FftParams fp, int nt, int batch, const double* __restrict__ G,
                                                     double* __restrict__ grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon) {
    extern __shared__ double2 smem2[];
    const int K = fp.K, N = t.nmax, R = t.nrows, NC = t.NC, Bp = t.Bp;
    const int row = blockIdx.x;
    const int b0 = blockIdx.y * nt;
    double2* A = smem2;
    double2* B = smem2 + (size_t)nt * K;
    const int i1 = t.row_i1[row], i2 = t.row_i2[row];
    // ---- spectrum assembly: thread <-> (transform, k); sums every order aliasing onto k
    for (int idx = threadIdx.x; idx < nt * K; idx += blockDim.x) {
        int k = idx / nt, tr = idx - k * nt;      // tr fastest: neighbouring threads read neighbouring fields
        int b = b0 + tr;
        double2 W = make_double2(0.0, 0.0);
        if (b < batch) {
            // orders m == k (mod K): + (Z + i Z')/2 ; orders m == -k (mod K): + (conj Z + i conj Z')/2
            for (int pass = 0; pass < 2; ++pass) {
                int mstart = pass == 0 ? k : (k == 0 ? 0 : K - k);
                for (int m = mstart; m <= N; m += K) {
                    const double* g = G + (((int64_t)m * R + row) * 2) * NC + b;
                    double Ec = g[0], Es = g[Bp], Oc = g[NC], Os = g[NC + Bp];
                    double2 ph = t.phase[m];
                    // Z = (A - iB) * ph, primary: A=Ec+Oc, B=Es+Os ; mirror: A'=Ec-Oc, B'=Es-Os
                    double2 Z = cmul(make_double2(Ec + Oc, -(Es + Os)), ph);
                    double2 Zp = i2 >= 0 ? cmul(make_double2(Ec - Oc, -(Es - Os)), ph) : make_double2(0.0, 0.0);
                    if (pass == 0) { W.x += 0.5 * (Z.x - Zp.y); W.y += 0.5 * (Z.y + Zp.x); }
                    else { W.x += 0.5 * (Z.x + Zp.y); W.y += 0.5 * (-Z.y + Zp.x); }
                }
            }
        }
        A[(size_t)tr * K + k] = W;
    }
    __syncthreads();
    double2* res = run_fft(A, B, fp, nt, t.tw, +1.0);
    // ---- store: real part -> primary latitude, imaginary part -> mirror latitude
    for (int idx = threadIdx.x; idx < nt * K; idx += blockDim.x) {
        int tr = idx / K, j = idx - tr * K;
        int b = b0 + tr;
        if (b >= batch) continue;
        double2 v = res[(size_t)tr * K + j];
        grid[(int64_t)b * gs_b + (int64_t)i1 * gs_lat + (int64_t)j * gs_lon] = v.x;
        if (i2 >= 0) grid[(int64_t)b * gs_b + (int64_t)i2 * gs_lat + (int64_t)j * gs_lon] = v.y;
    }
}

// Analysis: grid -> G (quadrature weights applied).
__global__ void __launch_bounds__(256) k_lon_ana_fft(DevTables t, FftParams fp, int nt, int batch, const double* __restrict__ grid,
                                                     int64_t gs_b, int64_t gs_lat, int64_t gs_lon, double* __restrict__ G) {
    extern __shared__ double2 smem2[];
    const int K = fp.K, N = t.nmax, R = t.nrows, NC = t.NC, Bp = t.Bp;
    const int row = blockIdx.x;
    const int b0 = blockIdx.y * nt;
    double2* A = smem2;
    double2* B = smem2 + (size_t)nt * K;
    const int i1 = t.row_i1[row], i2 = t.row_i2[row];
    const double w1 = t.w1[row], w2 = t.w2[row];
    for (int idx = threadIdx.x; idx < nt * K; idx += blockDim.x) {
        int tr = idx / K, j = idx - tr * K;
        int b = b0 + tr;
        double2 v = make_double2(0.0, 0.0);
        if (b < batch) {
            v.x = grid[(int64_t)b * gs_b + (int64_t)i1 * gs_lat + (int64_t)j * gs_lon];
            if (i2 >= 0) v.y = grid[(int64_t)b * gs_b + (int64_t)i2 * gs_lat + (int64_t)j * gs_lon];
        }
        A[(size_t)tr * K + j] = v;
    }
    __syncthreads();
    double2* res = run_fft(A, B, fp, nt, t.tw, -1.0);
    for (int idx = threadIdx.x; idx < nt * (N + 1); idx += blockDim.x) {
        int m = idx / nt, tr = idx - m * nt;
        int b = b0 + tr;
        if (b0 + tr >= Bp) continue;
        int k = m % K, kc = (K - k) % K;
        double2 Wk = res[(size_t)tr * K + k], Wc = cconj(res[(size_t)tr * K + kc]);
        // spectra of the two real rows
        double2 F = make_double2(0.5 * (Wk.x + Wc.x), 0.5 * (Wk.y + Wc.y));
        double2 D = csub(Wk, Wc);
        double2 Fp = make_double2(0.5 * D.y, -0.5 * D.x);          // (Wk - conj(W_{K-k})) / (2i)
        double2 phc = cconj(t.phase[m]);
        double2 Z = cmul(F, phc), Zp = cmul(Fp, phc);              // a - i b
        double a = w1 * Z.x, bb = -w1 * Z.y, ap = w2 * Zp.x, bp = -w2 * Zp.y;
        double* g = G + (((int64_t)m * R + row) * 2) * NC + b;
        bool livecol = b < batch;
        g[0] = livecol ? a + ap : 0.0; g[Bp] = livecol ? bb + bp : 0.0;
        g[NC] = livecol ? a - ap : 0.0; g[NC + Bp] = livecol ? bb - bp : 0.0;
    }
}

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
            const double* g = G + (((int64_t)m * R + row) * 2) * NC + b;
            double Ec = g[0], Es = g[Bp], Oc = g[NC], Os = g[NC + Bp];
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
        double* g = G + (((int64_t)m * R + row) * 2) * NC + b;
        g[0] = a + ap; g[Bp] = bb + bp; g[NC] = a - ap; g[NC + Bp] = bb - bp;
    }
}

// ------------------------------------------------------------------------------------------------
int lon_fft_smem_bytes(int K, int ntrans) { return (int)(2 * sizeof(double2) * (size_t)K * ntrans); }

int lon_fft_pick_ntrans(int K, int Bp) {
    const int limit = 200 * 1024;
    int nt = 0;
    for (int c = 1; c <= 8 && c <= Bp; c *= 2)
        if (lon_fft_smem_bytes(K, c) <= limit && (c == 1 || (size_t)c * K <= 8192)) nt = c;
    return nt;   // 0: K too large for the shared-memory FFT
}

static FftParams make_params(const Plan& p) {
    FftParams fp;
    fp.K = p.K; fp.nrad = (int)p.radices.size();
    for (int i = 0; i < fp.nrad; ++i) fp.rad[i] = p.radices[i];
    return fp;
}

int launch_lon_syn(const Plan& p, int batch, double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st) {
    if (p.lon_fft) {
        int nt = p.fft_ntrans;
        int smem = lon_fft_smem_bytes(p.K, nt);
        static int attr = 0;
        if (attr < smem) { SHB_CUDA(cudaFuncSetAttribute(k_lon_syn_fft, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr = 227 * 1024; }
        dim3 g(p.nrows, (batch + nt - 1) / nt);
        k_lon_syn_fft<<<g, 256, smem, st>>>(p.t, make_params(p), nt, batch, p.G, grid, gs_b, gs_lat, gs_lon);
    } else {
        int points = p.mode == SHB200_MODE_POINTS;
        int nl = points ? 1 : p.nlon;
        int64_t tot = (int64_t)p.nrows * nl;
        dim3 g((unsigned)((tot + 127) / 128));
        k_lon_syn_dense<<<g, 128, 0, st>>>(p.t, batch, points, p.G, grid, gs_b, gs_lat, gs_lon);
    }
    return 0;
}

int launch_lon_ana(const Plan& p, int batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st) {
    if (p.lon_fft) {
        int nt = p.fft_ntrans;
        int smem = lon_fft_smem_bytes(p.K, nt);
        static int attr = 0;
        if (attr < smem) { SHB_CUDA(cudaFuncSetAttribute(k_lon_ana_fft, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr = 227 * 1024; }
        dim3 g(p.nrows, (p.Bp + nt - 1) / nt);
        k_lon_ana_fft<<<g, 256, smem, st>>>(p.t, make_params(p), nt, batch, grid, gs_b, gs_lat, gs_lon, p.G);
    } else {
        int64_t tot = (int64_t)p.nrows * (p.nmax + 1);
        dim3 g((unsigned)((tot + 127) / 128));
        k_lon_ana_dense<<<g, 128, 0, st>>>(p.t, batch, grid, gs_b, gs_lat, gs_lon, p.G);
    }
    return 0;
}

}  // namespace shb
