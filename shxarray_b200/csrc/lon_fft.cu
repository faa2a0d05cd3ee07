// Longitude stage, FFT path: batched complex FFTs in shared memory, two real latitude rows per transform.
//
// What it replaces in the reference: the trig factors cos(m*lon), sin(|m|*lon) of Ynm_cpp::set
// (Ynm.hpp:114,121-131) that shlib folds into its per-point dgemv / dger (synthesis.pyx:183, analysis.pyx:136).
//
// Periodic equispaced longitudes lon_j = lon0 + j*360/K:
//   synthesis: rows i (primary) and i' (mirror) of one latitude pair are the real and imaginary part of ONE complex
//   inverse FFT ("two-for-one"): the Hermitian spectra are assembled from G[m][row][E/O][cos/sin][b] (E+O -> primary,
//   E-O -> mirror) with the phase twist e^{i m lon0}; orders m >= K/2 alias onto k = +-m mod K exactly as cos(m*lon_j)
//   does on the grid.
//   analysis: forward FFT of f_i + i f_i', split into the two real spectra, de-twist, quadrature weights, E/O.
// Algorithm: in-place mixed-radix (8,4,2,3,5) Cooley-Tukey in shared memory, one buffer per transform (16 B/point,
// index skewed by i>>3 against bank conflicts): decimation in time for synthesis (spectrum written straight to its
// digit-reversed slot, natural-order output -> coalesced row stores), decimation in frequency for analysis (natural
// input, spectrum picked up from the digit-reversed slots).  NF consecutive fields of one latitude pair per CTA so that
// every G access is a full 64 B segment.
#include <algorithm>
#include <cstdlib>
#include "internal.h"
#include "fft_common.cuh"

namespace shb {
namespace fft {

constexpr int MAXTHREADS = 512;
constexpr int MAXRAD = 16;

struct Params {
    int K, nrad, NF, TPT, stride, nthreads;
    int rad[MAXRAD];   // DIT order (first pass has sub-transform length 1)
    int ns[MAXRAD];    // product of the previous radices
    unsigned magic[MAXRAD];   // ceil(2^32 / ns): idx / ns == __umulhi(idx, magic) for idx * ns < 2^32 (ns >= 2)
    static constexpr bool is_static = false;
    __host__ __device__ int K_() const { return K; }
    __host__ __device__ int nrad_() const { return nrad; }
    __host__ __device__ int NF_() const { return NF; }
    __host__ __device__ int TPT_() const { return TPT; }
    __host__ __device__ int stride_() const { return stride; }
};

// The same parameters as compile-time constants of K (for plans of >= 8 fields): the kernels below are instantiated for the
// longitude counts of the usual grids (launch_lon_*_fft), which turns every stride, trip count and division of the
// passes into immediates and drops the radices a length does not use (r01: -7 % / -11 % on the cfg3 longitude stages).
template <int KK>
struct Factors {
    int n = 0;
    int rad[MAXRAD] = {};
    int ns[MAXRAD] = {};
    unsigned magic[MAXRAD] = {};
    constexpr Factors() {
        int k = KK;
        while (k % 8 == 0 && n < MAXRAD) { rad[n++] = 8; k /= 8; }
        while (k % 4 == 0 && n < MAXRAD) { rad[n++] = 4; k /= 4; }
        while (k % 2 == 0 && n < MAXRAD) { rad[n++] = 2; k /= 2; }
        while (k % 3 == 0 && n < MAXRAD) { rad[n++] = 3; k /= 3; }
        while (k % 5 == 0 && n < MAXRAD) { rad[n++] = 5; k /= 5; }
        int m = 1;
        for (int i = 0; i < n; ++i) { ns[i] = m; magic[i] = m >= 2 ? (unsigned)((0x100000000ULL + m - 1) / m) : 0u; m *= rad[i]; }
        if (k != 1) n = 0;
    }
};
template <int KK>
struct StaticParams {
    static constexpr bool is_static = true;
    static constexpr Factors<KK> F{};
    static constexpr int K = KK, NRAD = F.n;
    static constexpr int stride = KK + (KK >> 3) + 1;
    static constexpr size_t per = (size_t)stride * 16;
    static constexpr int NF = 8 * per <= 220 * 1024 ? 8 : (4 * per <= 220 * 1024 ? 4 : (2 * per <= 220 * 1024 ? 2 : 1));
    static constexpr int TPT = (NF == 8 && 2 * (8 * per + 2048) <= 220 * 1024) ? 32 : MAXTHREADS / NF;
    template <int s> static constexpr int rad_v = F.rad[s];
    template <int s> static constexpr int ns_v = F.ns[s];
    template <int s> static constexpr unsigned magic_v = F.magic[s];
    __host__ __device__ static constexpr int K_() { return K; }
    __host__ __device__ static constexpr int nrad_() { return NRAD; }
    __host__ __device__ static constexpr int NF_() { return NF; }
    __host__ __device__ static constexpr int TPT_() { return TPT; }
    __host__ __device__ static constexpr int stride_() { return stride; }
};

__device__ __forceinline__ int padi(int i) { return i + (i >> 3); }
// one in-place pass over one transform (buffer x, skewed indexing).  DIT: twiddle inputs; DIF: twiddle outputs.
// Butterflies are processed in chunks of UB with their twiddle loads (global, L2 latency) issued together up front.
// IO = 1: the last DIT pass (sub-transform length Ns*R = K) sends its outputs straight to the two grid rows (real part ->
// primary latitude, imaginary part -> mirror) instead of back to shared memory; IO = 2: the first DIF pass takes its inputs
// straight from the grid rows.  Element k + q*Ns of thread k: consecutive threads touch consecutive longitudes.
template <int R, bool DIT, int IO>
__device__ __forceinline__ void pass(double2* __restrict__ x, int K, int Ns, unsigned magic, int tpt, int lt, const double2* __restrict__ tw, double sgn, const RowIO& io) {
    // IO = 2 keeps the row values of the whole chunk in registers: at most 16 complex values per thread
    constexpr int UB = IO == 2 ? (R >= 8 ? 2 : (R >= 4 ? 3 : 4)) : (R >= 8 ? 3 : 4);
    const int M = Ns * R, nb = K / R, tws = K / M;
    const bool pow2 = (Ns & (Ns - 1)) == 0;            // true for all passes before the first radix 3/5
    const int sh = 31 - __clz(Ns);
    for (int i0 = lt; i0 < nb; i0 += tpt * UB) {
        double2 w1[UB];
        int base[UB], kk[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int idx = i0 + u * tpt;
            const int g = pow2 ? idx >> sh : (int)__umulhi((unsigned)idx, magic), k = idx - g * Ns;
            base[u] = idx < nb ? g * M + k : -1;
            kk[u] = k;
            w1[u] = (idx < nb && k > 0) ? tw[k * tws] : make_double2(1.0, 0.0);
        }
        double2 vin[IO == 2 ? UB : 1][IO == 2 ? R : 1];
        if (IO == 2) {
            // all row loads of the chunk in flight before the first butterfly
#pragma unroll
            for (int u = 0; u < UB; ++u)
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    double2 v = make_double2(0.0, 0.0);
                    if (base[u] >= 0 && io.live) {
                        const int64_t j = (int64_t)(base[u] + q * Ns) * io.gs_lon;
                        v.x = io.o1[j];
                        if (io.mirror) v.y = io.o2[j];
                    }
                    vin[IO == 2 ? u : 0][IO == 2 ? q : 0] = v;
                }
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            if (base[u] < 0) continue;
            double2 v[R], w[R];
            const bool tw_on = kk[u] > 0;
            if (tw_on) {
                double2 ww = w1[u];
                if (sgn > 0) ww.y = -ww.y;
                twiddle_powers<R>(ww, w);
            }
#pragma unroll
            for (int q = 0; q < R; ++q) v[q] = IO == 2 ? vin[IO == 2 ? u : 0][IO == 2 ? q : 0] : x[padi(base[u] + q * Ns)];
            if (DIT && tw_on) {
#pragma unroll
                for (int q = 1; q < R; ++q) v[q] = cmul(v[q], w[q]);
            }
            dft<R>(v, sgn);
            if (!DIT && tw_on) {
#pragma unroll
                for (int q = 1; q < R; ++q) v[q] = cmul(v[q], w[q]);
            }
            if (IO == 1) {
                if (io.live) {
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        const int64_t j = (int64_t)(base[u] + q * Ns) * io.gs_lon;
                        io.o1[j] = v[q].x;
                        if (io.mirror) io.o2[j] = v[q].y;
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < R; ++q) x[padi(base[u] + q * Ns)] = v[q];
            }
        }
    }
}

template <bool DIT, int IO>
__device__ __forceinline__ void run_pass(const Params& fp, int s, double2* x, int lt, const double2* tw, double sgn, const RowIO& io) {
    const int R = fp.rad[s], K = fp.K, Ns = fp.ns[s], tpt = fp.TPT_();
    const unsigned magic = fp.magic[s];
    if (R == 8) pass<8, DIT, IO>(x, K, Ns, magic, tpt, lt, tw, sgn, io);
    else if (R == 4) pass<4, DIT, IO>(x, K, Ns, magic, tpt, lt, tw, sgn, io);
    else if (R == 2) pass<2, DIT, IO>(x, K, Ns, magic, tpt, lt, tw, sgn, io);
    else if (R == 3) pass<3, DIT, IO>(x, K, Ns, magic, tpt, lt, tw, sgn, io);
    else pass<5, DIT, IO>(x, K, Ns, magic, tpt, lt, tw, sgn, io);
}
template <class FP, bool DIT, int IO, int s>
__device__ __forceinline__ void run_pass_c(double2* x, int lt, const double2* tw, double sgn, const RowIO& io) {
    constexpr int R = FP::template rad_v<s>, Ns = FP::template ns_v<s>;
    constexpr unsigned magic = FP::template magic_v<s>;
    pass<R, DIT, IO>(x, FP::K, Ns, magic, FP::TPT, lt, tw, sgn, io);
}

// all passes of one transform.  Synthesis (DIT): passes 0 .. nrad-2 in place, the last one straight to the grid rows;
// analysis (DIF): the first one (s = nrad-1) straight from the grid rows, then nrad-2 .. 0 in place.
// SM = true: the boundary pass stays in shared memory too (batch-fastest grids, [lat][lon][B] as shlib lays them out:
// the CTA then moves the 8 fields of a longitude together, 64 contiguous bytes, instead of one strided row per transform).
template <class FP, bool SM, int s>
__device__ __forceinline__ void syn_passes_c(double2* x, int tr, int lt, const double2* tw, const RowIO& io) {
    if constexpr (s + 1 < FP::NRAD) {
        run_pass_c<FP, true, 0, s>(x, lt, tw, +1.0, io);
        group_sync(1 + tr, FP::TPT);
        syn_passes_c<FP, SM, s + 1>(x, tr, lt, tw, io);
    } else {
        run_pass_c<FP, true, SM ? 0 : 1, s>(x, lt, tw, +1.0, io);
    }
}
template <class FP, bool SM, int s, class Hook>
__device__ __forceinline__ void ana_passes_c(double2* x, int tr, int lt, const double2* tw, const RowIO& io, Hook&& after_first) {
    if constexpr (s == FP::NRAD - 1) { run_pass_c<FP, false, SM ? 0 : 2, s>(x, lt, tw, -1.0, io); if (!SM) after_first(); }
    else run_pass_c<FP, false, 0, s>(x, lt, tw, -1.0, io);
    if constexpr (s > 0) {
        group_sync(1 + tr, FP::TPT);
        ana_passes_c<FP, SM, s - 1>(x, tr, lt, tw, io, after_first);
    }
}
template <class FP, bool SM>
__device__ __forceinline__ void syn_passes(const FP& fp, double2* x, int tr, int lt, const double2* tw, const RowIO& io) {
    if constexpr (FP::is_static) {
        syn_passes_c<FP, SM, 0>(x, tr, lt, tw, io);
    } else {
        for (int s = 0; s + 1 < fp.nrad; ++s) {
            run_pass<true, 0>(fp, s, x, lt, tw, +1.0, io);
            group_sync(1 + tr, fp.TPT_());
        }
        run_pass<true, SM ? 0 : 1>(fp, fp.nrad - 1, x, lt, tw, +1.0, io);
    }
}
template <class FP, bool SM, class Hook>
__device__ __forceinline__ void ana_passes(const FP& fp, double2* x, int tr, int lt, const double2* tw, const RowIO& io, Hook&& after_first) {
    if constexpr (FP::is_static) {
        ana_passes_c<FP, SM, FP::NRAD - 1>(x, tr, lt, tw, io, after_first);
    } else {
        run_pass<false, SM ? 0 : 2>(fp, fp.nrad - 1, x, lt, tw, -1.0, io);
        if (!SM) after_first();
        if (fp.nrad > 1) group_sync(1 + tr, fp.TPT_());
        for (int s = fp.nrad - 2; s >= 0; --s) {
            run_pass<false, 0>(fp, s, x, lt, tw, -1.0, io);
            if (s > 0) group_sync(1 + tr, fp.TPT_());
        }
    }
}

// ------------------------------------------------------------------------------------------------
template <class FP>
__global__ void __launch_bounds__(MAXTHREADS) k_lon_syn_fft(DevTables t, FP fp, const int* __restrict__ perm, int batch, const double* __restrict__ G,
                                                             double* __restrict__ grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, int pf_ahead) {
    extern __shared__ double2 smem2[];
    const int K = fp.K_(), N = t.nmax, R = t.nrows, NF = fp.NF_();
    const int row = blockIdx.y, b0 = blockIdx.x * NF;
    const int tid = threadIdx.x, NTHREADS = blockDim.x;
    const int i1 = t.row_i1[row], i2 = t.row_i2[row];
    const bool mirror = i2 >= 0;
    // ---- spectrum assembly: thread <-> (field f = tid % NF, orders m = tid / NF + i * NTHREADS / NF); fields fastest so
    // that a warp reads full 64/128-byte groups of G.  NF is a power of two.
    const bool simple = 2 * N <= K;                    // no aliasing: every slot is written at most once (m = K/2 folds onto itself like m = 0)
    const int fsh = 31 - __clz(NF);
    const int f = tid & (NF - 1), mstep = NTHREADS >> fsh;
    double2* xf = smem2 + (size_t)f * fp.stride_();
    const bool flive = b0 + f < batch;
    const int64_t gstep = (int64_t)R * 2 * t.Gs;       // G stride between consecutive orders
    const double* gcol = G + ((int64_t)row * 2) * t.Gs + colidx(0, b0 + f, t.GS);
    if (simple) {
        // slots N < k < K-N receive nothing: zero them (and everything for a padded field)
        for (int k = N + 1 + (tid >> fsh); k < K - N; k += mstep) xf[padi(perm[k])] = make_double2(0.0, 0.0);
        if (!flive)
            for (int k = tid >> fsh; k < K; k += mstep) xf[padi(k)] = make_double2(0.0, 0.0);
    } else {
        for (int i = tid; i < NF * fp.stride_(); i += NTHREADS) smem2[i] = make_double2(0.0, 0.0);
        __syncthreads();
    }
    const int nrounds = simple ? 1 : (N + K) / K;
    for (int rnd = 0; rnd < nrounds; ++rnd) {
        const int mlo = rnd * K, mhi = min(N, mlo + K - 1);
        for (int sub = 0; sub < (simple ? 1 : 2); ++sub) {
            constexpr int UN = 4;                       // independent orders in flight per thread (6 spills at the 128-register cap and is slower)
            for (int m0 = mlo + (tid >> fsh); m0 <= mhi; m0 += mstep * UN) {
                double Ec[UN], Es[UN], Oc[UN], Os[UN];
                double2 ph[UN];
                int pk[UN], pkc[UN];
                bool ok[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const int m = m0 + u * mstep;
                    ok[u] = flive && m <= mhi;
                    if (ok[u]) {
                        const double* g = gcol + (int64_t)m * gstep;
                        Ec[u] = g[0]; Es[u] = g[t.GS]; Oc[u] = g[t.Gs]; Os[u] = g[t.Gs + t.GS];
                        ph[u] = t.phase[m];
                        const int mm = m - mlo;
                        pk[u] = perm[mm]; pkc[u] = perm[mm ? K - mm : 0];
                    }
                }
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    if (!ok[u]) continue;
                    // primary row: A = Ec+Oc, B = Es+Os; mirror row: A' = Ec-Oc, B' = Es-Os (zero without a mirror row)
                    // Wa = (Z + iZ')/2 = [(A+B') + i(A'-B)] ph/2 -> k = m mod K ; Wb = (conj Z + i conj Z')/2 = [(A-B') + i(A'+B)] conj(ph)/2 -> k = -m mod K
                    const double A = Ec[u] + Oc[u], B = Es[u] + Os[u];
                    const double Ap = mirror ? Ec[u] - Oc[u] : 0.0, Bp_ = mirror ? Es[u] - Os[u] : 0.0;
                    const double2 hp = make_double2(0.5 * ph[u].x, 0.5 * ph[u].y);
                    const double2 Wa = cmul(make_double2(A + Bp_, Ap - B), hp);
                    const double2 Wb = cmul(make_double2(A - Bp_, Ap + B), cconj(hp));
                    if (simple) {
                        if (m0 + u * mstep == 0 || 2 * (m0 + u * mstep) == K) xf[padi(pk[u])] = cadd(Wa, Wb);
                        else { xf[padi(pk[u])] = Wa; xf[padi(pkc[u])] = Wb; }
                    } else if (sub == 0) {
                        double2* p = &xf[padi(pk[u])]; *p = cadd(*p, Wa);
                    } else {
                        double2* p = &xf[padi(pkc[u])]; *p = cadd(*p, Wb);
                    }
                }
            }
            __syncthreads();
        }
    }
    // ---- L2 prefetch for the CTA that will take this one's place: CTAs start in linear order, `ahead` of them are resident
    // at once, so CTA L + ahead begins about when this one ends; its slab of G (one 128-byte line per order and parity at 8
    // fields) then waits in L2 instead of DRAM when its assembly starts (ncu r01: 16 % of this kernel's samples were the
    // first use of those loads)
    {
        const int64_t L = (int64_t)blockIdx.y * gridDim.x + blockIdx.x + pf_ahead;
        const int prow = (int)(L / gridDim.x), pgrp = (int)(L - (int64_t)prow * gridDim.x);
        if (pf_ahead > 0 && prow < R) {
            const char* base = reinterpret_cast<const char*>(G + ((int64_t)prow * 2) * t.Gs + colidx(0, pgrp * NF, t.GS));
            const int64_t ostep = gstep * 8, estep = (int64_t)t.Gs * 8;       // bytes between orders / parities
            const int span = NF * 16;                                           // bytes of the group's cos+sin columns
            for (int i = tid; i < 2 * (N + 1); i += NTHREADS) {
                const char* a = base + (int64_t)(i >> 1) * ostep + (i & 1) * estep;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
                if (span > 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + 128));
            }
        }
    }
    // ---- inverse FFT (decimation in time)
    const int tr = tid / fp.TPT_(), lt = tid - tr * fp.TPT_();
    double2* x = smem2 + (size_t)tr * fp.stride_();
    // the transforms of a CTA are independent from here on: per-transform named barriers let their passes and row
    // stores drift apart and overlap instead of marching in lock step
    const int b = b0 + tr;
    RowIO io;
    io.live = b < batch; io.mirror = mirror; io.gs_lon = gs_lon;
    io.o1 = grid + (int64_t)b * gs_b + (int64_t)i1 * gs_lat;
    io.o2 = grid + (int64_t)b * gs_b + (int64_t)(mirror ? i2 : 0) * gs_lat;
    // ---- passes; the last one stores: real part -> primary latitude, imaginary part -> mirror latitude
    if (gs_b == 1 && NF == 8) {
        // batch-fastest grid: finish in shared memory, then thread <-> (longitude j, field f): a warp writes 4 longitudes x 64 B
        syn_passes<FP, true>(fp, x, tr, lt, t.tw, io);
        __syncthreads();
        const int ff = tid & 7;
        if (b0 + ff < batch) {
            const double2* xs = smem2 + (size_t)ff * fp.stride_();
            double* q1 = grid + (b0 + ff) + (int64_t)i1 * gs_lat;
            double* q2 = grid + (b0 + ff) + (int64_t)(mirror ? i2 : 0) * gs_lat;
            for (int j = tid >> 3; j < K; j += NTHREADS >> 3) {
                const double2 v = xs[padi(j)];
                q1[(int64_t)j * gs_lon] = v.x;
                if (mirror) q2[(int64_t)j * gs_lon] = v.y;
            }
        }
    } else {
        syn_passes<FP, false>(fp, x, tr, lt, t.tw, io);
    }
}

template <class FP>
__global__ void __launch_bounds__(MAXTHREADS) k_lon_ana_fft(DevTables t, FP fp, const int* __restrict__ perm, int batch, const double* __restrict__ grid,
                                                             int64_t gs_b, int64_t gs_lat, int64_t gs_lon, double* __restrict__ G, int pf_ahead) {
    extern __shared__ double2 smem2[];
    const int K = fp.K_(), N = t.nmax, R = t.nrows, Bp = t.Bp, NF = fp.NF_();
    const int row = blockIdx.y, b0 = blockIdx.x * NF;
    const int tid = threadIdx.x, NTHREADS = blockDim.x;
    const int i1 = t.row_i1[row], i2 = t.row_i2[row];
    const bool mirror = i2 >= 0;
    const double w1 = t.w1[row], w2 = t.w2[row];
    const int tr = tid / fp.TPT_(), lt = tid - tr * fp.TPT_();
    double2* x = smem2 + (size_t)tr * fp.stride_();
    {
        // ---- first pass straight from the grid rows (primary latitude -> real part, mirror -> imaginary part); fields past
        // the batch enter as zeros
        const int b = b0 + tr;
        RowIO io;
        io.live = b < batch; io.mirror = mirror; io.gs_lon = gs_lon;
        io.o1 = const_cast<double*>(grid) + (int64_t)b * gs_b + (int64_t)i1 * gs_lat;
        io.o2 = const_cast<double*>(grid) + (int64_t)b * gs_b + (int64_t)(mirror ? i2 : 0) * gs_lat;
        // L2 prefetch of the grid rows of the CTA that will take this one's place (see k_lon_syn_fft), issued once this
        // transform's own row loads are done (before them it costs more than it saves: 0.40 -> 0.46 ms at cfg3)
        auto prefetch_next = [&]() {
            const int64_t L = (int64_t)blockIdx.y * gridDim.x + blockIdx.x + pf_ahead;
            const int prow = (int)(L / gridDim.x), pb = (int)(L - (int64_t)prow * gridDim.x) * NF + tr;
            if (pf_ahead > 0 && prow < R && pb < batch && gs_lon == 1) {
                const int p1 = t.row_i1[prow], p2 = t.row_i2[prow];
                const char* r1 = reinterpret_cast<const char*>(grid + (int64_t)pb * gs_b + (int64_t)p1 * gs_lat);
                const char* r2 = reinterpret_cast<const char*>(grid + (int64_t)pb * gs_b + (int64_t)(p2 >= 0 ? p2 : p1) * gs_lat);
                for (int i = lt * 128; i < K * 8; i += fp.TPT_() * 128) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(r1 + i));
                    if (p2 >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(r2 + i));
                }
            }
        };
        if (gs_b == 1 && NF == 8) {
            // batch-fastest grid: thread <-> (longitude j, field f) loads 64 contiguous bytes per longitude into the transform
            // buffers, then every pass runs in shared memory
            const int ff = tid & 7;
            const bool flive = b0 + ff < batch;
            double2* xs = smem2 + (size_t)ff * fp.stride_();
            const double* q1 = grid + (b0 + ff) + (int64_t)i1 * gs_lat;
            const double* q2 = grid + (b0 + ff) + (int64_t)(mirror ? i2 : 0) * gs_lat;
#pragma unroll 8
            for (int j = tid >> 3; j < K; j += NTHREADS >> 3) {
                double2 v = make_double2(0.0, 0.0);
                if (flive) { v.x = q1[(int64_t)j * gs_lon]; if (mirror) v.y = q2[(int64_t)j * gs_lon]; }
                xs[padi(j)] = v;
            }
            {   // successor prefetch for this layout: one 64-byte piece (8 fields) per longitude and row
                const int64_t L = (int64_t)blockIdx.y * gridDim.x + blockIdx.x + pf_ahead;
                const int prow = (int)(L / gridDim.x), pb0 = (int)(L - (int64_t)prow * gridDim.x) * NF;
                if (pf_ahead > 0 && prow < R && pb0 < batch) {
                    const int p1 = t.row_i1[prow], p2 = t.row_i2[prow];
                    const double* r1 = grid + pb0 + (int64_t)p1 * gs_lat;
                    const double* r2 = grid + pb0 + (int64_t)(p2 >= 0 ? p2 : p1) * gs_lat;
                    for (int j = tid; j < K; j += NTHREADS) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(r1 + (int64_t)j * gs_lon));
                        if (p2 >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(r2 + (int64_t)j * gs_lon));
                    }
                }
            }
            __syncthreads();
            ana_passes<FP, true>(fp, x, tr, lt, t.tw, io, prefetch_next);
        } else {
            ana_passes<FP, false>(fp, x, tr, lt, t.tw, io, prefetch_next);
        }
    }
    __syncthreads();
    {
        // thread <-> (field f = tid % NF, orders m = tid / NF + i * NTHREADS / NF)
        const int fsh = 31 - __clz(NF);
        const int f = tid & (NF - 1), mstep = NTHREADS >> fsh;
        const int b = b0 + f;
        if (b < Bp) {
            const double2* xx = smem2 + (size_t)f * fp.stride_();
            const bool live = b < batch;
            const int64_t gstep = (int64_t)R * 2 * t.Gs;
            double* gcol = G + ((int64_t)row * 2) * t.Gs + colidx(0, b, t.GS);
            // UN orders per thread at a time: their table look-ups (two digit-reversal entries and the phase factor, all from
            // global memory) are issued together before the first dependent shared-memory read; a plainly unrolled loop waited
            // for each order's look-up in turn (ncu r01: 11 % of the kernel's samples)
            constexpr int UN = 4;
            const int kstep = mstep % K;
            int k = (tid >> fsh) % K;
            if (NF < 8) {
                // long transforms (2 or 4 fields per CTA, many orders per thread): the plain loop is faster (measured,
                // K = 4320: 0.20 vs 0.23 ms for one field)
#pragma unroll 4
                for (int m = tid >> fsh; m <= N; m += mstep) {
                    const int kc = k ? K - k : 0;
                    const int kcur = k;
                    k += kstep; if (k >= K) k -= K;
                    if (t.mstride > 1 && (m < t.m0 || (m - t.m0) % t.mstride)) continue;
                    const double2 Wk = xx[padi(perm[kcur])], Wc = cconj(xx[padi(perm[kc])]);
                    const double2 F = make_double2(0.5 * (Wk.x + Wc.x), 0.5 * (Wk.y + Wc.y));
                    const double2 D = csub(Wk, Wc);
                    const double2 Fp = make_double2(0.5 * D.y, -0.5 * D.x);
                    const double2 phc = cconj(t.phase[m]);
                    const double2 Z = cmul(F, phc), Zp = cmul(Fp, phc);
                    const double a = w1 * Z.x, bb = -w1 * Z.y, ap = w2 * Zp.x, bp = -w2 * Zp.y;
                    double* g = gcol + (int64_t)m * gstep;
                    g[0] = live ? a + ap : 0.0; g[t.GS] = live ? bb + bp : 0.0;
                    g[t.Gs] = live ? a - ap : 0.0; g[t.Gs + t.GS] = live ? bb - bp : 0.0;
                }
            } else
            for (int m0 = tid >> fsh; m0 <= N; m0 += mstep * UN) {
                int pk[UN], pkc[UN];
                double2 ph[UN];
                bool ok[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const int m = m0 + u * mstep;
                    const int kc = k ? K - k : 0;
                    // an order integrated on another rank (m-sharding) is skipped
                    ok[u] = m <= N && !(t.mstride > 1 && (m < t.m0 || (m - t.m0) % t.mstride));
                    if (ok[u]) { pk[u] = perm[k]; pkc[u] = perm[kc]; ph[u] = t.phase[m]; }
                    k += kstep; if (k >= K) k -= K;
                }
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    if (!ok[u]) continue;
                    const int m = m0 + u * mstep;
                    const double2 Wk = xx[padi(pk[u])], Wc = cconj(xx[padi(pkc[u])]);
                    const double2 F = make_double2(0.5 * (Wk.x + Wc.x), 0.5 * (Wk.y + Wc.y));      // spectrum of the primary row
                    const double2 D = csub(Wk, Wc);
                    const double2 Fp = make_double2(0.5 * D.y, -0.5 * D.x);                        // (Wk - conj W_{K-k}) / (2i): mirror row
                    const double2 phc = cconj(ph[u]);
                    const double2 Z = cmul(F, phc), Zp = cmul(Fp, phc);                            // a_m - i b_m
                    const double a = w1 * Z.x, bb = -w1 * Z.y, ap = w2 * Zp.x, bp = -w2 * Zp.y;
                    double* g = gcol + (int64_t)m * gstep;
                    g[0] = live ? a + ap : 0.0; g[t.GS] = live ? bb + bp : 0.0;
                    g[t.Gs] = live ? a - ap : 0.0; g[t.Gs + t.GS] = live ? bb - bp : 0.0;
                }
            }
        }
    }
}

}  // namespace fft

// ---- host side ---------------------------------------------------------------------------------
static bool fft_params(int K, int Bp, fft::Params& fp) {
    int k = K, n = 0;
    int rad[fft::MAXRAD];
    while (k % 8 == 0 && n < fft::MAXRAD) { rad[n++] = 8; k /= 8; }
    while (k % 4 == 0 && n < fft::MAXRAD) { rad[n++] = 4; k /= 4; }
    while (k % 2 == 0 && n < fft::MAXRAD) { rad[n++] = 2; k /= 2; }
    while (k % 3 == 0 && n < fft::MAXRAD) { rad[n++] = 3; k /= 3; }
    while (k % 5 == 0 && n < fft::MAXRAD) { rad[n++] = 5; k /= 5; }
    if (k != 1) return false;
    fp.K = K; fp.nrad = n;
    int ns = 1;
    for (int i = 0; i < n; ++i) { fp.rad[i] = rad[i]; fp.ns[i] = ns; fp.magic[i] = ns >= 2 ? (unsigned)((0x100000000ULL + ns - 1) / ns) : 0u; ns *= rad[i]; }
    fp.stride = K + (K >> 3) + 1;
    // fields per CTA: as many as fit (8 = one full 128-byte line of G per order and parity); two CTAs of 4 fields per SM
    // measured no faster (r01: 0.61 vs 0.57 ms synthesis at cfg3)
    const size_t per = (size_t)fp.stride_() * sizeof(double2);
    int nf = 0;
    for (int c = 1; c <= 8 && c <= Bp; c *= 2)
        if (c * per <= 220 * 1024) nf = c;
    if (nf == 0) return false;
    fp.NF = nf;
    fp.TPT = fft::MAXTHREADS / nf;                 // a full 512-thread CTA: long transforms get more threads each
    // short transforms: two 256-thread CTAs per SM (the 128-register kernels fill the register file with 512 threads) overlap
    // one CTA's DRAM phases with the other's passes; one warp per transform
    if (nf == 8 && 2 * (nf * per + 2048) <= 220 * 1024) fp.TPT = 32;
    fp.nthreads = fp.NF * fp.TPT;
    return true;
}

bool lon_fft_supported(int K, int Bp) {
    fft::Params fp;
    return fft_params(K, Bp, fp);
}

// digit-reversal table matching the pass order of fft_params (position of spectral index k in the buffer)
void lon_fft_perm(int K, int Bp, std::vector<int>& perm) {
    fft::Params fp;
    perm.assign(K, 0);
    if (!fft_params(K, Bp, fp)) return;
    for (int k = 0; k < K; ++k) {
        int n = k, pos = 0;
        for (int s = fp.nrad - 1; s >= 0; --s) { int q = n % fp.rad[s]; n /= fp.rad[s]; pos += q * fp.ns[s]; }
        perm[k] = pos;
    }
}

// longitude counts with compile-time kernels: regular grids of 1 / 0.5 / 0.25 / 0.125 degree and 5', Gauss grids of
// nmax 96 ... 2160 (engine.py: _nice_nlon)
#define SHB_FFT_STATIC_LENGTHS(X) X(240) X(360) X(384) X(480) X(576) X(720) X(768) X(1152) X(1440) X(1536) X(2304) X(2880) X(3072) X(4320) X(4608) X(5760) X(11520)

template <class FP>
static int launch_syn_fp(const Plan& p, const FP& fp, int NF, int nthreads, size_t smem, int batch, double* grid, int64_t gs_b, int64_t gs_lat,
                         int64_t gs_lon, cudaStream_t st) {
    static PerDeviceOnce attr;
    if (attr.first()) { SHB_CUDA(cudaFuncSetAttribute(fft::k_lon_syn_fft<FP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); }
    dim3 g((batch + NF - 1) / NF, p.nrows);
    // CTAs resident at once = the prefetch distance
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fft::k_lon_syn_fft<FP>, nthreads, smem);
    static const int pf_env = getenv("SHB200_FFT_PREFETCH") ? atoi(getenv("SHB200_FFT_PREFETCH")) : 1;
    const int ahead = pf_env ? std::max(1, per_sm) * device_sm_count() : 0;
    trace_kernel(p, FP::is_static ? (NF == 8 ? "k_lon_syn_fft<static/8>" : "k_lon_syn_fft<static>") : "k_lon_syn_fft<runtime>");
    fft::k_lon_syn_fft<FP><<<g, nthreads, smem, st>>>(p.t, fp, p.fft_perm, batch, p.G, grid, gs_b, gs_lat, gs_lon, ahead);
    return 0;
}
template <class FP>
static int launch_ana_fp(const Plan& p, const FP& fp, int NF, int nthreads, size_t smem, int batch, const double* grid, int64_t gs_b, int64_t gs_lat,
                         int64_t gs_lon, cudaStream_t st) {
    static PerDeviceOnce attr;
    if (attr.first()) { SHB_CUDA(cudaFuncSetAttribute(fft::k_lon_ana_fft<FP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); }
    // field groups made of padding only are skipped: their G columns keep whatever they held (zeros after plan creation),
    // and the Legendre stage's results for padded fields are never unpacked
    dim3 g((batch + NF - 1) / NF, p.nrows);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fft::k_lon_ana_fft<FP>, nthreads, smem);
    static const int pf_env = getenv("SHB200_FFT_PREFETCH") ? atoi(getenv("SHB200_FFT_PREFETCH")) : 1;
    const int ahead = pf_env ? std::max(1, per_sm) * device_sm_count() : 0;
    trace_kernel(p, FP::is_static ? (NF == 8 ? "k_lon_ana_fft<static/8>" : "k_lon_ana_fft<static>") : "k_lon_ana_fft<runtime>");
    fft::k_lon_ana_fft<FP><<<g, nthreads, smem, st>>>(p.t, fp, p.fft_perm, batch, grid, gs_b, gs_lat, gs_lon, p.G, ahead);
    return 0;
}

int launch_lon_syn_fft(const Plan& p, int batch, double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st) {
    fft::Params fp;
    fft_params(p.K, p.Bp, fp);
    const size_t smem = (size_t)fp.NF * fp.stride * sizeof(double2);
    switch (p.K) {
#define X(KK)                                                                                                              \
    case KK:                                                                                                               \
        if (fft::StaticParams<KK>::NF == fp.NF && fft::StaticParams<KK>::TPT == fp.TPT)                                    \
            return launch_syn_fp(p, fft::StaticParams<KK>{}, fp.NF, fp.nthreads, smem, batch, grid, gs_b, gs_lat, gs_lon, st); \
        break;
        SHB_FFT_STATIC_LENGTHS(X)
#undef X
        default: break;
    }
    return launch_syn_fp(p, fp, fp.NF, fp.nthreads, smem, batch, grid, gs_b, gs_lat, gs_lon, st);
}

int launch_lon_ana_fft(const Plan& p, int batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st) {
    fft::Params fp;
    fft_params(p.K, p.Bp, fp);
    const size_t smem = (size_t)fp.NF * fp.stride * sizeof(double2);
    switch (p.K) {
#define X(KK)                                                                                                              \
    case KK:                                                                                                               \
        if (fft::StaticParams<KK>::NF == fp.NF && fft::StaticParams<KK>::TPT == fp.TPT)                                    \
            return launch_ana_fp(p, fft::StaticParams<KK>{}, fp.NF, fp.nthreads, smem, batch, grid, gs_b, gs_lat, gs_lon, st); \
        break;
        SHB_FFT_STATIC_LENGTHS(X)
#undef X
        default: break;
    }
    return launch_ana_fp(p, fp, fp.NF, fp.nthreads, smem, batch, grid, gs_b, gs_lat, gs_lon, st);
}

}  // namespace shb
