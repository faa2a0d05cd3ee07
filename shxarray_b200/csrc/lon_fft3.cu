// Longitude stage, FFT path for the usual grid lengths: three-pass (two-pass for short rows) composite-radix FFT.
//
// Same mathematics and the same spectrum assembly / pick-up as lon_fft.cu (which stays the path for any other 2,3,5-smooth
// length, for batch-fastest grids and for plans of fewer than 8 fields); what changes is the transform itself:
//   * K = R0 * R1 * R2 with R0 in {8, 16} and composite radices up to 20 done entirely in registers (a radix-16 butterfly is
//     4 x 4 with constant inner twiddles), so K = 1536 = 16*16*6 makes three trips through shared memory instead of the four
//     of 8*8*8*3, and K = 720 = 8*9*10 three instead of five;
//   * no padding: the buffers are XOR-swizzled (unit e of a transform lives at e ^ ((e >> log2 R0) & 7)), which frees the
//     shared memory for the twiddle tables of the passes ([q][k] order: consecutive threads read consecutive entries), so no
//     pass waits for L2;
//   * every index, trip count and division is a compile-time constant of K.
// It replaces the trig factors of Ynm_cpp::set (Ynm.hpp:114,121-131) exactly like lon_fft.cu does.
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <utility>
#include "internal.h"
#include "fft_common.cuh"
#include "fft_roots.h"

namespace shb {
namespace fft3 {
using namespace fft;

// ---- compile-time loops ----------------------------------------------------------------------------------------
template <class F, int... I>
__device__ __forceinline__ void sfor_impl(F&& f, std::integer_sequence<int, I...>) { (f(std::integral_constant<int, I>{}), ...); }
template <int N, class F>
__device__ __forceinline__ void sfor(F&& f) { sfor_impl(f, std::make_integer_sequence<int, N>{}); }

// ---- composite butterflies ---------------------------------------------------------------------------------------
// R = A * B, n = n1*B + n2, k = k1 + A*k2:  X[k] = sum_n2 W_B^(n2 k2) [ W_R^(n2 k1) sum_n1 W_A^(n1 k1) x[n1*B + n2] ]
template <int R> struct Split { static constexpr int A = 0, B = 0; };
template <> struct Split<6> { static constexpr int A = 2, B = 3; };
template <> struct Split<9> { static constexpr int A = 3, B = 3; };
template <> struct Split<10> { static constexpr int A = 2, B = 5; };
template <> struct Split<12> { static constexpr int A = 4, B = 3; };
template <> struct Split<15> { static constexpr int A = 3, B = 5; };
template <> struct Split<16> { static constexpr int A = 4, B = 4; };
template <> struct Split<18> { static constexpr int A = 2, B = 9; };
template <> struct Split<20> { static constexpr int A = 4, B = 5; };

// a * exp(SGN * 2 pi i J / R) with the factor folded at compile time
template <int R, int J, int SGN>
__device__ __forceinline__ double2 mulroot(double2 a) {
    if constexpr (J == 0) return a;
    else if constexpr (2 * J == R) return make_double2(-a.x, -a.y);
    else if constexpr (4 * J == R) return SGN > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
    else if constexpr (4 * J == 3 * R) return SGN > 0 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x);
    else if constexpr ((8 * J) % R == 0) {
        constexpr double h = 0.70710678118654752440;
        constexpr int o = 8 * J / R;                                  // odd octant 1, 3, 5, 7
        constexpr double sc = (o == 1 || o == 7) ? 1.0 : -1.0;        // sign of cos
        constexpr double ss = ((o == 1 || o == 3) ? 1.0 : -1.0) * SGN;  // sign of SGN * sin
        return make_double2(h * (sc * a.x - ss * a.y), h * (sc * a.y + ss * a.x));
    } else {
        constexpr double c = fftroots::rc(R, J), s = SGN * fftroots::rs(R, J);
        return make_double2(c * a.x - s * a.y, c * a.y + s * a.x);
    }
}

template <int R, int SGN>
struct DftN {
    static __device__ __forceinline__ void run(double2* v) {
        if constexpr (Split<R>::A == 0) {
            dft<R>(v, (double)SGN);
        } else {
            constexpr int A = Split<R>::A, B = Split<R>::B;
            double2 y[R];
            sfor<B>([&](auto n2c) {
                constexpr int n2 = decltype(n2c)::value;
                double2 t[A];
                sfor<A>([&](auto n1c) { constexpr int n1 = decltype(n1c)::value; t[n1] = v[n1 * B + n2]; });
                DftN<A, SGN>::run(t);
                sfor<A>([&](auto k1c) { constexpr int k1 = decltype(k1c)::value; y[k1 * B + n2] = mulroot<R, (n2 * k1) % R, SGN>(t[k1]); });
            });
            sfor<A>([&](auto k1c) {
                constexpr int k1 = decltype(k1c)::value;
                double2 t[B];
                sfor<B>([&](auto n2c) { constexpr int n2 = decltype(n2c)::value; t[n2] = y[k1 * B + n2]; });
                DftN<B, SGN>::run(t);
                sfor<B>([&](auto k2c) { constexpr int k2 = decltype(k2c)::value; v[k1 + A * k2] = t[k2]; });
            });
        }
    }
};

// ---- configuration of one length -----------------------------------------------------------------------------------
template <int K_, int R0_, int R1_, int R2_, int NF_, int TPT_>
struct Cfg {
    static constexpr int K = K_, R0 = R0_, R1 = R1_, R2 = R2_, NF = NF_, TPT = TPT_, NT = NF_ * TPT_;
    static constexpr int NP = R2_ > 1 ? 3 : 2;
    static constexpr int MINB = NF_ * TPT_ <= 256 ? 2 : 1;      // CTAs per SM the register budget must allow
    static constexpr int S = R0_ == 16 ? 4 : 3;              // swizzle shift
    static constexpr int BUF = K_ + 1;                        // 16-byte units between the transforms of a CTA (odd: bank skew)
    static constexpr int T1N = (R1_ - 1) * R0_;               // middle pass: full table [q-1][k], k < R0
    static constexpr int T2N = R2_ > 1 ? R0_ * R1_ : 0;       // last pass: w^k, k < R0*R1; the powers are formed in registers
    static constexpr size_t SMEM = ((size_t)NF_ * BUF + T1N + T2N) * sizeof(double2);
    static_assert(K_ == R0_ * R1_ * R2_, "factorisation");
    static_assert(R0_ == 8 || R0_ == 16, "first radix");
    template <int P> static constexpr int rad_v = P == 0 ? R0_ : (P == 1 ? R1_ : R2_);
    template <int P> static constexpr int ns_v = P == 0 ? 1 : (P == 1 ? R0_ : R0_ * R1_);
};

// cycle counters of shb200_debug_fft_cycles
__device__ int g_prof_on = 0;
__device__ unsigned long long g_prof[16];
struct Prof {
    long long t;
    bool on;
    __device__ __forceinline__ void start(int tid) { on = tid == 0 && g_prof_on; if (on) { t = clock64(); } }
    __device__ __forceinline__ void mark(int slot) {
        if (on) { const long long n = clock64(); atomicAdd(&g_prof[slot], (unsigned long long)(n - t)); t = n; }
    }
    __device__ __forceinline__ void count(int slot) { if (on) atomicAdd(&g_prof[slot], 1ULL); }
};

template <int S>
__device__ __forceinline__ int swz(int e) { return e ^ ((e >> S) & 7); }

template <int TPT>
__device__ __forceinline__ void tsync(int tr) {
    if constexpr (TPT <= 32) __syncwarp();
    else group_sync(1 + tr, TPT);
}

// w^1 .. w^(R-1) from w^1, product depth ceil(log2 R)
template <int R>
__device__ __forceinline__ void powers(double2 w1, double2* w) {
    w[1] = w1;
    sfor<R - 2>([&](auto qc) { constexpr int q = decltype(qc)::value + 2; w[q] = cmul(w[q / 2], w[q - q / 2]); });
}

// One pass over one transform.  DIT (synthesis, SGN = +1): inputs twiddled, pass P = 0 first; DIF (analysis, SGN = -1):
// outputs twiddled, pass NP-1 first.  IO = 1: outputs go straight to the grid rows (last DIT pass); IO = 2: inputs come
// straight from the grid rows (first DIF pass).
template <class C, int P, bool DIT, int IO>
__device__ __forceinline__ void pass3(double2* __restrict__ x, const double2* __restrict__ T1, const double2* __restrict__ T2, int lt, const RowIO& io) {
    constexpr int R = C::template rad_v<P>, Ns = C::template ns_v<P>, NB = C::K / R, TPT = C::TPT, SGN = DIT ? 1 : -1;
    constexpr int NCH = (NB + TPT - 1) / TPT;
    constexpr bool w1mode = P == 2;                        // only the third pass forms its twiddle powers in registers
    constexpr int UB = IO == 2 ? (R <= 4 ? 4 : (R <= 5 ? 3 : (R <= 8 ? 2 : 1))) : 1;      // chunks whose row loads are issued together
#pragma unroll 1
    for (int c0 = 0; c0 < NCH; c0 += UB) {
        double2 v[UB][R];
        int base[UB], kk[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int idx = lt + (c0 + u) * TPT;
            const int g = Ns == 1 ? idx : idx / Ns, k = Ns == 1 ? 0 : idx - g * Ns;
            kk[u] = k;
            base[u] = (c0 + u < NCH && idx < NB) ? g * (Ns * R) + k : -1;
        }
        if constexpr (IO == 2) {
#pragma unroll
            for (int u = 0; u < UB; ++u)
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    double2 a = make_double2(0.0, 0.0);
                    if (base[u] >= 0 && io.live) {
                        const int64_t j = (int64_t)(base[u] + q * Ns) * io.gs_lon;
                        a.x = io.o1[j];
                        if (io.mirror) a.y = io.o2[j];
                    }
                    v[u][q] = a;
                }
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            if (base[u] < 0) continue;
            const int k = kk[u];
            if constexpr (IO != 2) {
#pragma unroll
                for (int q = 0; q < R; ++q) v[u][q] = x[swz<C::S>(base[u] + q * Ns)];
            }
            if constexpr (DIT && P > 0) {
                if constexpr (w1mode) {
                    double2 w[R];
                    powers<R>(T2[k], w);
#pragma unroll
                    for (int q = 1; q < R; ++q) v[u][q] = cmul(v[u][q], w[q]);
                } else {
#pragma unroll
                    for (int q = 1; q < R; ++q) v[u][q] = cmul(v[u][q], T1[(q - 1) * Ns + k]);
                }
            }
            DftN<R, SGN>::run(v[u]);
            if constexpr (!DIT && P > 0) {
                if constexpr (w1mode) {
                    double2 w[R];
                    powers<R>(T2[k], w);
#pragma unroll
                    for (int q = 1; q < R; ++q) v[u][q] = cmul(v[u][q], w[q]);
                } else {
#pragma unroll
                    for (int q = 1; q < R; ++q) v[u][q] = cmul(v[u][q], T1[(q - 1) * Ns + k]);
                }
            }
            if constexpr (IO == 1) {
                if (io.live) {
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        const int64_t j = (int64_t)(base[u] + q * Ns) * io.gs_lon;
                        io.o1[j] = v[u][q].x;
                        if (io.mirror) io.o2[j] = v[u][q].y;
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < R; ++q) x[swz<C::S>(base[u] + q * Ns)] = v[u][q];
            }
        }
    }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// twiddle tables of this CTA: T1[(q-1)*R0 + k] = W^(k q K/(R0 R1)), T2[k] = W^(k K/(R0 R1 R2)); W = exp(-2 pi i/K) forward,
// conjugated for the inverse transform
template <class C, bool INV, int TWS = 1>                         // TWS: the plan's root table belongs to the length TWS * C::K
__device__ __forceinline__ void fill_tables(const double2* __restrict__ tw, double2* T1, double2* T2, int tid) {
    for (int i = tid; i < C::T1N; i += C::NT) {
        const int q = i / C::R0 + 1, k = i - (q - 1) * C::R0;
        double2 w = tw[k * q * (C::K / (C::R0 * C::R1)) * TWS];
        if (INV) w.y = -w.y;
        T1[i] = w;
    }
    for (int i = tid; i < C::T2N; i += C::NT) {
        double2 w = tw[i * TWS];                                // K / (R0 R1 R2) = 1
        if (INV) w.y = -w.y;
        T2[i] = w;
    }
}

// ------------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(C::NT, C::MINB) k_lon_syn_fft3(DevTables t, const int* __restrict__ perm, int batch, const double* __restrict__ G,
                                                        double* __restrict__ grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, int pf_ahead) {
    extern __shared__ double2 smem2[];
    constexpr int K = C::K, NF = C::NF, NT = C::NT, S = C::S;
    double2* T1 = smem2 + (size_t)NF * C::BUF;
    double2* T2 = T1 + C::T1N;
    const int N = t.nmax, R = t.nrows;
    const int row = blockIdx.y, b0 = blockIdx.x * NF;
    const int tid = threadIdx.x;
    const int i1 = t.row_i1[row], i2 = t.row_i2[row];
    const bool mirror = i2 >= 0;
    Prof prof;
    prof.start(tid);
    prof.count(0);
    fill_tables<C, true>(t.tw, T1, T2, tid);
    // ---- spectrum assembly: thread <-> (field f = tid % NF, orders m = tid / NF + i * NT / NF); fields fastest: a warp reads
    // whole groups of G (NF * 16 bytes per order and parity)
    const bool simple = 2 * N <= K;
    constexpr int fsh = NF == 8 ? 3 : (NF == 4 ? 2 : (NF == 2 ? 1 : 0));
    constexpr int mstep = NT >> fsh;
    const int f = tid & (NF - 1);
    double2* xf = smem2 + (size_t)f * C::BUF;
    const bool flive = b0 + f < batch;
    const int64_t gstep = (int64_t)R * 2 * t.Gs;
    const double* gcol = G + ((int64_t)row * 2) * t.Gs + colidx(0, b0 + f, t.GS);
    if (simple) {
        for (int k = N + 1 + (tid >> fsh); k < K - N; k += mstep) xf[swz<S>(perm[k])] = make_double2(0.0, 0.0);
        if (!flive)
            for (int k = tid >> fsh; k < K; k += mstep) xf[k] = make_double2(0.0, 0.0);
    } else {
        for (int i = tid; i < NF * C::BUF; i += NT) smem2[i] = make_double2(0.0, 0.0);
        __syncthreads();
    }
    const int nrounds = simple ? 1 : (N + K) / K;
    for (int rnd = 0; rnd < nrounds; ++rnd) {
        const int mlo = rnd * K, mhi = min(N, mlo + K - 1);
        for (int sub = 0; sub < (simple ? 1 : 2); ++sub) {
            constexpr int UN = 4;
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
                    // see lon_fft.cu: Wa -> k = m mod K, Wb -> k = -m mod K
                    const double A = Ec[u] + Oc[u], B = Es[u] + Os[u];
                    const double Ap = mirror ? Ec[u] - Oc[u] : 0.0, Bp_ = mirror ? Es[u] - Os[u] : 0.0;
                    const double2 hp = make_double2(0.5 * ph[u].x, 0.5 * ph[u].y);
                    const double2 Wa = cmul(make_double2(A + Bp_, Ap - B), hp);
                    const double2 Wb = cmul(make_double2(A - Bp_, Ap + B), cconj(hp));
                    if (simple) {
                        if (m0 + u * mstep == 0 || 2 * (m0 + u * mstep) == K) xf[swz<S>(pk[u])] = cadd(Wa, Wb);
                        else { xf[swz<S>(pk[u])] = Wa; xf[swz<S>(pkc[u])] = Wb; }
                    } else if (sub == 0) {
                        double2* p = &xf[swz<S>(pk[u])]; *p = cadd(*p, Wa);
                    } else {
                        double2* p = &xf[swz<S>(pkc[u])]; *p = cadd(*p, Wb);
                    }
                }
            }
            __syncthreads();
        }
    }
    prof.mark(1);
    // ---- L2 prefetch of the G slab of the CTA that will take this one's place (lon_fft.cu)
    {
        const int64_t L = (int64_t)blockIdx.y * gridDim.x + blockIdx.x + pf_ahead;
        const int prow = (int)(L / gridDim.x), pgrp = (int)(L - (int64_t)prow * gridDim.x);
        if (pf_ahead > 0 && prow < R) {
            const char* base = reinterpret_cast<const char*>(G + ((int64_t)prow * 2) * t.Gs + colidx(0, pgrp * NF, t.GS));
            const int64_t ostep = gstep * 8, estep = (int64_t)t.Gs * 8;
            for (int i = tid; i < 2 * (N + 1); i += NT) {
                const char* a = base + (int64_t)(i >> 1) * ostep + (i & 1) * estep;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
            }
        }
    }
    // ---- inverse FFT (decimation in time); the transforms of a CTA are independent from here on
    const int tr = tid / C::TPT, lt = tid - tr * C::TPT;
    double2* x = smem2 + (size_t)tr * C::BUF;
    const int b = b0 + tr;
    RowIO io;
    io.live = b < batch; io.mirror = mirror; io.gs_lon = gs_lon;
    io.o1 = grid + (int64_t)b * gs_b + (int64_t)i1 * gs_lat;
    io.o2 = grid + (int64_t)b * gs_b + (int64_t)(mirror ? i2 : 0) * gs_lat;
    pass3<C, 0, true, 0>(x, T1, T2, lt, io);
    tsync<C::TPT>(tr);
    prof.mark(2);
    if constexpr (C::NP == 3) {
        pass3<C, 1, true, 0>(x, T1, T2, lt, io);
        tsync<C::TPT>(tr);
        prof.mark(3);
        pass3<C, 2, true, 1>(x, T1, T2, lt, io);
    } else {
        pass3<C, 1, true, 1>(x, T1, T2, lt, io);
    }
    prof.mark(4);
}

template <class C>
__global__ void __launch_bounds__(C::NT, C::MINB) k_lon_ana_fft3(DevTables t, const int* __restrict__ perm, int batch, const double* __restrict__ grid,
                                                        int64_t gs_b, int64_t gs_lat, int64_t gs_lon, double* __restrict__ G, int pf_ahead) {
    extern __shared__ double2 smem2[];
    constexpr int K = C::K, NF = C::NF, NT = C::NT, S = C::S;
    double2* T1 = smem2 + (size_t)NF * C::BUF;
    double2* T2 = T1 + C::T1N;
    const int N = t.nmax, R = t.nrows, Bp = t.Bp;
    const int row = blockIdx.y, b0 = blockIdx.x * NF;
    const int tid = threadIdx.x;
    const int i1 = t.row_i1[row], i2 = t.row_i2[row];
    const bool mirror = i2 >= 0;
    const double w1 = t.w1[row], w2 = t.w2[row];
    const int tr = tid / C::TPT, lt = tid - tr * C::TPT;
    double2* x = smem2 + (size_t)tr * C::BUF;
    Prof prof;
    prof.start(tid);
    prof.count(8);
    fill_tables<C, false>(t.tw, T1, T2, tid);
    {
        const int b = b0 + tr;
        RowIO io;
        io.live = b < batch; io.mirror = mirror; io.gs_lon = gs_lon;
        io.o1 = const_cast<double*>(grid) + (int64_t)b * gs_b + (int64_t)i1 * gs_lat;
        io.o2 = const_cast<double*>(grid) + (int64_t)b * gs_b + (int64_t)(mirror ? i2 : 0) * gs_lat;
        // the first pass is twiddled (DIF twiddles outputs), so the tables must be complete first
        __syncthreads();
        prof.mark(9);
        if constexpr (C::NP == 3) pass3<C, 2, false, 2>(x, T1, T2, lt, io);
        else pass3<C, 1, false, 2>(x, T1, T2, lt, io);
        {   // L2 prefetch of the grid rows of the CTA that will take this one's place
            const int64_t L = (int64_t)blockIdx.y * gridDim.x + blockIdx.x + pf_ahead;
            const int prow = (int)(L / gridDim.x), pb = (int)(L - (int64_t)prow * gridDim.x) * NF + tr;
            if (pf_ahead > 0 && prow < R && pb < batch && gs_lon == 1) {
                const int p1 = t.row_i1[prow], p2 = t.row_i2[prow];
                const char* r1 = reinterpret_cast<const char*>(grid + (int64_t)pb * gs_b + (int64_t)p1 * gs_lat);
                const char* r2 = reinterpret_cast<const char*>(grid + (int64_t)pb * gs_b + (int64_t)(p2 >= 0 ? p2 : p1) * gs_lat);
                for (int i = lt * 128; i < K * 8; i += C::TPT * 128) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(r1 + i));
                    if (p2 >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(r2 + i));
                }
            }
        }
        tsync<C::TPT>(tr);
        prof.mark(10);
        if constexpr (C::NP == 3) {
            pass3<C, 1, false, 0>(x, T1, T2, lt, io);
            tsync<C::TPT>(tr);
            prof.mark(11);
        }
        pass3<C, 0, false, 0>(x, T1, T2, lt, io);
    }
    __syncthreads();
    prof.mark(12);
    {
        // spectrum pick-up: thread <-> (field f = tid % NF, orders m = tid / NF + i * NT / NF)
        constexpr int fsh = NF == 8 ? 3 : (NF == 4 ? 2 : (NF == 2 ? 1 : 0));
        constexpr int mstep = NT >> fsh;
        const int f = tid & (NF - 1);
        const int b = b0 + f;
        if (b < Bp) {
            const double2* xx = smem2 + (size_t)f * C::BUF;
            const bool live = b < batch;
            const int64_t gstep = (int64_t)R * 2 * t.Gs;
            double* gcol = G + ((int64_t)row * 2) * t.Gs + colidx(0, b, t.GS);
            constexpr int UN = 4;
            constexpr int kstep = mstep % K;
            int k = (tid >> fsh) % K;
            for (int m0 = tid >> fsh; m0 <= N; m0 += mstep * UN) {
                int pk[UN], pkc[UN];
                double2 ph[UN];
                bool ok[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const int m = m0 + u * mstep;
                    const int kc = k ? K - k : 0;
                    ok[u] = m <= N && !(t.mstride > 1 && (m < t.m0 || (m - t.m0) % t.mstride));      // m-sharding: foreign orders are skipped
                    if (ok[u]) { pk[u] = perm[k]; pkc[u] = perm[kc]; ph[u] = t.phase[m]; }
                    k += kstep; if (k >= K) k -= K;
                }
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    if (!ok[u]) continue;
                    const int m = m0 + u * mstep;
                    const double2 Wk = xx[swz<S>(pk[u])], Wc = cconj(xx[swz<S>(pkc[u])]);
                    const double2 F = make_double2(0.5 * (Wk.x + Wc.x), 0.5 * (Wk.y + Wc.y));      // spectrum of the primary row
                    const double2 D = csub(Wk, Wc);
                    const double2 Fp = make_double2(0.5 * D.y, -0.5 * D.x);                        // mirror row
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
    prof.mark(13);
}

// ---- rows of length 2 * C::K (shlib's own nmax-2160 grid: 11520 = 2 x 5760) ---------------------------------------------------
// A transform of that length does not fit the three-pass scheme (radices <= 20) and, alone in the 227 KB of a CTA, ran as six
// in-place passes (lon_fft.cu: 7x its HBM time).  One radix-2 step by hand instead, two three-pass transforms of length C::K per
// (latitude pair, field) in one CTA:
//   synthesis: x[2j] = IDFT_K(X[k] + X[k+K]), x[2j+1] = IDFT_K((X[k] - X[k+K]) w^k), w = exp(2 pi i / 2K).  With 2 nmax < K the
//     only non-zero bins are k = m and 2K - m, so the even half gets (Wa, Wb) at (m, K - m) like a length-K row and the odd half
//     (Wa w^m, Wb conj(w^m)); the halves write the even / odd longitudes of the grid row (stride 2 gs_lon);
//   analysis: E = DFT_K(x[2j]), O = DFT_K(x[2j+1]);  X[m] = E[m] + tw[m] O[m],  X[2K - m] = E[K - m] + conj(tw[m]) O[K - m].
template <class C>
__global__ void __launch_bounds__(C::NT, 1) k_lon_syn_fft3s2(DevTables t, const int* __restrict__ perm, int batch, const double* __restrict__ G,
                                                             double* __restrict__ grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon) {
    extern __shared__ double2 smem2[];
    constexpr int K = C::K, NT = C::NT, S = C::S;
    static_assert(C::NF == 2 && C::NP == 3, "two halves per CTA");
    double2* T1 = smem2 + (size_t)2 * C::BUF;
    double2* T2 = T1 + C::T1N;
    const int N = t.nmax, R = t.nrows;
    const int row = blockIdx.y, b = blockIdx.x;
    const int tid = threadIdx.x;
    const int i1 = t.row_i1[row], i2 = t.row_i2[row];
    const bool mirror = i2 >= 0;
    fill_tables<C, true, 2>(t.tw, T1, T2, tid);
    double2* xe = smem2;
    double2* xo = smem2 + C::BUF;
    for (int k = N + 1 + tid; k < K - N; k += NT) { const int q = swz<S>(perm[k]); xe[q] = make_double2(0.0, 0.0); xo[q] = make_double2(0.0, 0.0); }
    const int64_t gstep = (int64_t)R * 2 * t.Gs;
    const double* gcol = G + ((int64_t)row * 2) * t.Gs + colidx(0, b, t.GS);
    constexpr int UN = 4;
    for (int m0 = tid; m0 <= N; m0 += NT * UN) {
        double Ec[UN], Es[UN], Oc[UN], Os[UN];
        double2 ph[UN], wm[UN];
        int pk[UN], pkc[UN];
        bool ok[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int m = m0 + u * NT;
            ok[u] = m <= N;
            if (ok[u]) {
                const double* g = gcol + (int64_t)m * gstep;
                Ec[u] = g[0]; Es[u] = g[t.GS]; Oc[u] = g[t.Gs]; Os[u] = g[t.Gs + t.GS];
                ph[u] = t.phase[m];
                wm[u] = cconj(t.tw[m]);                                   // exp(+2 pi i m / 2K)
                pk[u] = perm[m]; pkc[u] = perm[m ? K - m : 0];
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            if (!ok[u]) continue;
            const double A = Ec[u] + Oc[u], B = Es[u] + Os[u];
            const double Ap = mirror ? Ec[u] - Oc[u] : 0.0, Bp_ = mirror ? Es[u] - Os[u] : 0.0;
            const double2 hp = make_double2(0.5 * ph[u].x, 0.5 * ph[u].y);
            const double2 Wa = cmul(make_double2(A + Bp_, Ap - B), hp);
            const double2 Wb = cmul(make_double2(A - Bp_, Ap + B), cconj(hp));
            if (m0 + u * NT == 0) {
                const double2 s0 = cadd(Wa, Wb);
                xe[swz<S>(pk[u])] = s0; xo[swz<S>(pk[u])] = s0;
            } else {
                xe[swz<S>(pk[u])] = Wa; xe[swz<S>(pkc[u])] = Wb;
                xo[swz<S>(pk[u])] = cmul(Wa, wm[u]); xo[swz<S>(pkc[u])] = cmul(Wb, cconj(wm[u]));
            }
        }
    }
    __syncthreads();
    const int tr = tid / C::TPT, lt = tid - tr * C::TPT;               // 0: even longitudes, 1: odd longitudes
    double2* x = smem2 + (size_t)tr * C::BUF;
    RowIO io;
    io.live = b < batch; io.mirror = mirror; io.gs_lon = 2 * gs_lon;
    io.o1 = grid + (int64_t)b * gs_b + (int64_t)i1 * gs_lat + tr * gs_lon;
    io.o2 = grid + (int64_t)b * gs_b + (int64_t)(mirror ? i2 : 0) * gs_lat + tr * gs_lon;
    pass3<C, 0, true, 0>(x, T1, T2, lt, io);
    tsync<C::TPT>(tr);
    pass3<C, 1, true, 0>(x, T1, T2, lt, io);
    tsync<C::TPT>(tr);
    pass3<C, 2, true, 1>(x, T1, T2, lt, io);
}

template <class C>
__global__ void __launch_bounds__(C::NT, 1) k_lon_ana_fft3s2(DevTables t, const int* __restrict__ perm, int batch, const double* __restrict__ grid,
                                                             int64_t gs_b, int64_t gs_lat, int64_t gs_lon, double* __restrict__ G) {
    extern __shared__ double2 smem2[];
    constexpr int K = C::K, NT = C::NT, S = C::S;
    static_assert(C::NF == 2 && C::NP == 3, "two halves per CTA");
    double2* T1 = smem2 + (size_t)2 * C::BUF;
    double2* T2 = T1 + C::T1N;
    const int N = t.nmax, R = t.nrows;
    const int row = blockIdx.y, b = blockIdx.x;
    const int tid = threadIdx.x;
    const int i1 = t.row_i1[row], i2 = t.row_i2[row];
    const bool mirror = i2 >= 0;
    const double w1 = t.w1[row], w2 = t.w2[row];
    const int tr = tid / C::TPT, lt = tid - tr * C::TPT;
    double2* x = smem2 + (size_t)tr * C::BUF;
    fill_tables<C, false, 2>(t.tw, T1, T2, tid);
    {
        RowIO io;
        io.live = b < batch; io.mirror = mirror; io.gs_lon = 2 * gs_lon;
        io.o1 = const_cast<double*>(grid) + (int64_t)b * gs_b + (int64_t)i1 * gs_lat + tr * gs_lon;
        io.o2 = const_cast<double*>(grid) + (int64_t)b * gs_b + (int64_t)(mirror ? i2 : 0) * gs_lat + tr * gs_lon;
        __syncthreads();
        pass3<C, 2, false, 2>(x, T1, T2, lt, io);
        tsync<C::TPT>(tr);
        pass3<C, 1, false, 0>(x, T1, T2, lt, io);
        tsync<C::TPT>(tr);
        pass3<C, 0, false, 0>(x, T1, T2, lt, io);
    }
    __syncthreads();
    if (b < t.Bp) {
        const double2* xe = smem2;
        const double2* xo = smem2 + C::BUF;
        const bool live = b < batch;
        const int64_t gstep = (int64_t)R * 2 * t.Gs;
        double* gcol = G + ((int64_t)row * 2) * t.Gs + colidx(0, b, t.GS);
        for (int m = tid; m <= N; m += NT) {
            if (t.mstride > 1 && (m < t.m0 || (m - t.m0) % t.mstride)) continue;       // m-sharding: foreign orders are skipped
            const int pk = swz<S>(perm[m]), pkc = swz<S>(perm[m ? K - m : 0]);
            const double2 wm = t.tw[m];                                                // exp(-2 pi i m / 2K)
            const double2 Wk = cadd(xe[pk], cmul(wm, xo[pk]));
            const double2 Wc = cconj(cadd(xe[pkc], cmul(cconj(wm), xo[pkc])));
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
    }
}

// ---- persistent synthesis kernel: the next work item's G slab streams into shared memory while this one is transformed ------
// One CTA per SM walks the (latitude pair, field group) items.  What the two-CTA kernel above cannot hide at the long lengths
// is its input phase (r02 cycle counters at cfg3: 53 % of a CTA's time is the spectrum assembly waiting for G -- registers
// bound the bytes in flight).  Here the slab of item i+1 is fetched with cp.async (LDGSTS.128, no registers, 92 KB in flight
// per SM) into a staging area as soon as the assembly of item i has consumed its own, i.e. during the passes and row stores of
// item i, and the assembly reads shared memory.  Measured (profiles/r02_fft_variants.md): cfg3 synthesis 0.349 -> 0.305 ms;
// the same idea for analysis (rows staged, first pass from shared memory) was slower than the two-CTA kernel (0.356 vs
// 0.311 ms: the scattered G stores of the pick-up keep the LSU busy when the next rows should be issued) and is not kept;
// at K <= 768 sixteen one-warp transforms per SM (two CTAs) beat eight two-warp ones, so the short lengths stay there too.
template <class C>
struct PCfg {
    static constexpr int MAXORD = C::K / 2 + 1;                                   // orders without aliasing
    static constexpr int ROWD = 4 * C::NF;                                        // doubles per order in the staged G slab: 4 chunks (E/O x cos/sin) of NF fields
    static constexpr size_t XB = ((size_t)C::NF * C::BUF + C::T1N + C::T2N) * sizeof(double2);
    static constexpr size_t STG_SYN = (size_t)MAXORD * ROWD * sizeof(double);
    static constexpr size_t PERM = ((size_t)C::K * sizeof(unsigned short) + 15) / 16 * 16;
    static constexpr size_t SMEM_SYN = XB + STG_SYN + PERM;
};

template <class C>
__global__ void __launch_bounds__(C::NT, 1) k_lon_syn_fft3p(DevTables t, const int* __restrict__ perm, int batch, const double* __restrict__ G,
                                                            double* __restrict__ grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, int ngroups, int nitems) {
    extern __shared__ double2 smem2[];
    using P = PCfg<C>;
    constexpr int K = C::K, NF = C::NF, NT = C::NT, S = C::S;
    double2* T1 = smem2 + (size_t)NF * C::BUF;
    double2* T2 = T1 + C::T1N;
    double* stg = reinterpret_cast<double*>(reinterpret_cast<char*>(smem2) + P::XB);
    unsigned short* sperm = reinterpret_cast<unsigned short*>(reinterpret_cast<char*>(stg) + P::STG_SYN);
    const int N = t.nmax, R = t.nrows;
    const int tid = threadIdx.x;
    Prof prof;
    prof.start(tid);
    fill_tables<C, true>(t.tw, T1, T2, tid);
    for (int k = tid; k < K; k += NT) sperm[k] = (unsigned short)swz<S>(perm[k]);
    const int64_t gstep = (int64_t)R * 2 * t.Gs;
    // the G slab of one item: per order 4 chunks (parity, cos/sin) of NF doubles; chunk c of order m is staged at slot c ^ (m & 3)
    // so that the assembly's reads (NF fields x 32/NF consecutive orders per warp) spread over the banks
    constexpr int PPC = NF / 2, PPO = 4 * PPC;             // 16-byte pieces per chunk / per order
    auto issue_stage = [&](int item) {
        const int row = item / ngroups, b0 = (item - row * ngroups) * NF;
        const double* gbase = G + ((int64_t)row * 2) * t.Gs + colidx(0, b0, t.GS);
        for (int p = tid; p < (N + 1) * PPO; p += NT) {
            const int m = p / PPO, r = p - m * PPO, c = r / PPC, sub = r - c * PPC;
            const double* src = gbase + (int64_t)m * gstep + (c >> 1) * t.Gs + (c & 1) * t.GS + sub * 2;
            cp_async16(stg + m * P::ROWD + ((c ^ (m & 3)) * NF) + sub * 2, src);
        }
        cp_async_commit();
    };
    int item = blockIdx.x;
    if (item < nitems) issue_stage(item);
    constexpr int fsh = NF == 8 ? 3 : (NF == 4 ? 2 : (NF == 2 ? 1 : 0));
    constexpr int mstep = NT >> fsh;
    const int f = tid & (NF - 1);
    double2* xf = smem2 + (size_t)f * C::BUF;
    const int tr = tid / C::TPT, lt = tid - tr * C::TPT;
    double2* x = smem2 + (size_t)tr * C::BUF;
    const bool unit_phase = t.phase_unit != 0;
    for (; item < nitems; item += gridDim.x) {
        const int row = item / ngroups, b0 = (item - row * ngroups) * NF;
        const int i1 = t.row_i1[row], i2 = t.row_i2[row];
        const bool mirror = i2 >= 0;
        const bool flive = b0 + f < batch;
        prof.count(0);
        cp_async_wait_all();
        __syncthreads();                                   // slab landed; every thread is done with the previous item's buffers
        prof.mark(5);
        // ---- spectrum assembly from the staged slab: thread <-> (field f, orders m = tid / NF + i * NT / NF)
        for (int k = N + 1 + (tid >> fsh); k < K - N; k += mstep) xf[sperm[k]] = make_double2(0.0, 0.0);
        if (!flive) {
            for (int k = tid >> fsh; k < K; k += mstep) xf[k] = make_double2(0.0, 0.0);
        } else {
#pragma unroll 2
            for (int m = tid >> fsh; m <= N; m += mstep) {
                const double* sp = stg + m * P::ROWD + f;
                const int sw = m & 3;
                const double Ec = sp[(0 ^ sw) * NF], Es = sp[(1 ^ sw) * NF], Oc = sp[(2 ^ sw) * NF], Os = sp[(3 ^ sw) * NF];
                // see lon_fft.cu: Wa -> k = m mod K, Wb -> k = -m mod K
                const double A = Ec + Oc, B = Es + Os;
                const double Ap = mirror ? Ec - Oc : 0.0, Bp_ = mirror ? Es - Os : 0.0;
                double2 Wa = make_double2(0.5 * (A + Bp_), 0.5 * (Ap - B)), Wb = make_double2(0.5 * (A - Bp_), 0.5 * (Ap + B));
                if (!unit_phase) {
                    const double2 ph = t.phase[m];
                    Wa = cmul(Wa, ph);
                    Wb = cmul(Wb, cconj(ph));
                }
                if (m == 0 || 2 * m == K) xf[sperm[m]] = cadd(Wa, Wb);
                else { xf[sperm[m]] = Wa; xf[sperm[K - m]] = Wb; }
            }
        }
        __syncthreads();                                   // spectra complete, staging area free
        prof.mark(1);
        if (item + (int)gridDim.x < nitems) issue_stage(item + gridDim.x);
        // ---- inverse FFT (decimation in time), one group of TPT threads per transform
        const int b = b0 + tr;
        RowIO io;
        io.live = b < batch; io.mirror = mirror; io.gs_lon = gs_lon;
        io.o1 = grid + (int64_t)b * gs_b + (int64_t)i1 * gs_lat;
        io.o2 = grid + (int64_t)b * gs_b + (int64_t)(mirror ? i2 : 0) * gs_lat;
        if (io.live) {
            pass3<C, 0, true, 0>(x, T1, T2, lt, io);
            tsync<C::TPT>(tr);
            prof.mark(2);
            if constexpr (C::NP == 3) {
                pass3<C, 1, true, 0>(x, T1, T2, lt, io);
                tsync<C::TPT>(tr);
                prof.mark(3);
                pass3<C, 2, true, 1>(x, T1, T2, lt, io);
            } else {
                pass3<C, 1, true, 1>(x, T1, T2, lt, io);
            }
            prof.mark(4);
        }
    }
    cp_async_wait_all();
}

}  // namespace fft3

// ---- host side ---------------------------------------------------------------------------------
// length, R0, R1, R2, fields per CTA, threads per transform.  Up to K = 1536 two CTAs share an SM (4 transforms each, or 8
// short ones); beyond, one CTA of 512 threads.
#define SHB_FFT3_CONFIGS(X)                                                                                              \
    X(240, 16, 15, 1, 8, 32) X(360, 8, 9, 5, 8, 32) X(384, 8, 8, 6, 8, 32) X(480, 8, 10, 6, 8, 32) X(576, 8, 9, 8, 8, 32)  \
    X(720, 8, 10, 9, 8, 32) X(768, 8, 12, 8, 8, 32) X(1152, 8, 12, 12, 4, 64) X(1440, 16, 10, 9, 4, 64)                    \
    X(1536, 16, 16, 6, 4, 64) X(2304, 16, 12, 12, 4, 128) X(2880, 16, 15, 12, 4, 128) X(3072, 16, 16, 12, 4, 128)          \
    X(4320, 16, 18, 15, 2, 256) X(4608, 16, 18, 16, 2, 256) X(5760, 16, 18, 20, 2, 256)

struct Fft3Info { int K, R0, R1, R2, NF, TPT; };
static const Fft3Info* fft3_find(int K) {
    static const Fft3Info tab[] = {
#define X(KK, A, B, C, NF, TPT) {KK, A, B, C, NF, TPT},
        SHB_FFT3_CONFIGS(X)
#undef X
    };
    static const int off = getenv("SHB200_FFT3") ? atoi(getenv("SHB200_FFT3")) : 1;
    if (!off) return nullptr;
    for (const auto& e : tab)
        if (e.K == K) return &e;
    return nullptr;
}

// rows of twice a listed two-transforms-per-CTA length (8640, 9216, 11520): k_lon_*_fft3s2, any number of fields
static const Fft3Info* fft3_find_split(int K) {
    static const int on = getenv("SHB200_FFT3S2") ? atoi(getenv("SHB200_FFT3S2")) : 1;
    if (!on || K % 2 || fft3_find(K)) return nullptr;
    const Fft3Info* e = fft3_find(K / 2);
    return e && e->NF == 2 && e->R2 > 1 ? e : nullptr;
}

// the three-pass kernels serve plans of >= 8 fields on the listed lengths (their G column groups are NF fields wide)
bool lon_fft3_plan(int K, int Bp, int* group_size) {
    if (fft3_find_split(K)) { if (group_size) *group_size = 0; return true; }      // group size: the plan's default
    const Fft3Info* e = fft3_find(K);
    if (!e || Bp < 8) return false;
    if (group_size) *group_size = e->NF >= 4 ? e->NF : 4;
    return true;
}

void lon_fft3_perm(int K, std::vector<int>& perm) {
    const Fft3Info* e = fft3_find(K);
    if (!e && (e = fft3_find_split(K))) K /= 2;                                    // table of the half length
    perm.assign(K, 0);
    if (!e) return;
    const int rad[3] = {e->R0, e->R1, e->R2}, ns[3] = {1, e->R0, e->R0 * e->R1};
    const int np = e->R2 > 1 ? 3 : 2;
    for (int k = 0; k < K; ++k) {
        int n = k, pos = 0;
        for (int s = np - 1; s >= 0; --s) { int q = n % rad[s]; n /= rad[s]; pos += q * ns[s]; }
        perm[k] = pos;
    }
}

template <class C>
static int launch_syn3(const Plan& p, int batch, double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st) {
    static PerDeviceOnce attr;
    if (attr.first()) { SHB_CUDA(cudaFuncSetAttribute(fft3::k_lon_syn_fft3<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM)); }
    dim3 g((batch + C::NF - 1) / C::NF, p.nrows);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fft3::k_lon_syn_fft3<C>, C::NT, C::SMEM);
    static const int pf_env = getenv("SHB200_FFT_PREFETCH") ? atoi(getenv("SHB200_FFT_PREFETCH")) : 1;
    const int ahead = pf_env ? std::max(1, per_sm) * device_sm_count() : 0;
    trace_kernel(p, "k_lon_syn_fft3");
    fft3::k_lon_syn_fft3<C><<<g, C::NT, C::SMEM, st>>>(p.t, p.fft3_perm, batch, p.G, grid, gs_b, gs_lat, gs_lon, ahead);
    return 0;
}
template <class C>
static int launch_ana3(const Plan& p, int batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st) {
    static PerDeviceOnce attr;
    if (attr.first()) { SHB_CUDA(cudaFuncSetAttribute(fft3::k_lon_ana_fft3<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM)); }
    dim3 g((batch + C::NF - 1) / C::NF, p.nrows);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fft3::k_lon_ana_fft3<C>, C::NT, C::SMEM);
    static const int pf_env = getenv("SHB200_FFT_PREFETCH") ? atoi(getenv("SHB200_FFT_PREFETCH")) : 1;
    const int ahead = pf_env ? std::max(1, per_sm) * device_sm_count() : 0;
    trace_kernel(p, "k_lon_ana_fft3");
    fft3::k_lon_ana_fft3<C><<<g, C::NT, C::SMEM, st>>>(p.t, p.fft3_perm, batch, grid, gs_b, gs_lat, gs_lon, p.G, ahead);
    return 0;
}


// persistent synthesis: lengths 1152 ... 1536, no aliasing (2 nmax <= K), contiguous rows
template <class C>
static int launch_syn3p(const Plan& p, int batch, double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st) {
    using P = fft3::PCfg<C>;
    static PerDeviceOnce attr;
    if (attr.first()) { SHB_CUDA(cudaFuncSetAttribute(fft3::k_lon_syn_fft3p<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::SMEM_SYN)); }
    const int ngroups = (batch + C::NF - 1) / C::NF, nitems = ngroups * p.nrows;
    trace_kernel(p, "k_lon_syn_fft3p");
    fft3::k_lon_syn_fft3p<C><<<std::min(nitems, device_sm_count()), C::NT, P::SMEM_SYN, st>>>(p.t, p.fft3_perm, batch, p.G, grid, gs_b, gs_lat, gs_lon, ngroups, nitems);
    return 0;
}
// length, R0, R1, R2, fields per CTA, threads per transform of the persistent synthesis kernel (512 threads, one CTA per SM)
// Three warps per transform: the 96 radix-16 butterflies of a K = 1536 pass (96 / 144 radix-12 / -8 ones at 1152, 90 / 144 / 160 at
// 1440) then occupy every thread; with four warps a quarter of the threads idled through two of the three passes and ptxas
// was held to 128 registers (cfg3 synthesis stage 0.324 -> 0.297 ms at 384 threads x 162 registers).  SHB200_FFT3P_TPT=128: r02a shape.
#define SHB_FFT3P_CONFIGS(X) X(1152, 8, 12, 12, 4, 96) X(1440, 16, 10, 9, 4, 96) X(1536, 16, 16, 6, 4, 96)
#define SHB_FFT3P_CONFIGS_128(X) X(1152, 8, 12, 12, 4, 128) X(1440, 16, 10, 9, 4, 128) X(1536, 16, 16, 6, 4, 128)

static bool fft3p_on() {
    static const int v = getenv("SHB200_FFT3P") ? atoi(getenv("SHB200_FFT3P")) : 1;
    return v != 0;
}

#define SHB_FFT3S2_CONFIGS(X) X(4320, 16, 18, 15, 2, 256) X(4608, 16, 18, 16, 2, 256) X(5760, 16, 18, 20, 2, 256)

template <class C>
static int launch_syn3s2(const Plan& p, int batch, double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st) {
    static PerDeviceOnce attr;
    if (attr.first()) { SHB_CUDA(cudaFuncSetAttribute(fft3::k_lon_syn_fft3s2<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM)); }
    trace_kernel(p, "k_lon_syn_fft3s2");
    fft3::k_lon_syn_fft3s2<C><<<dim3(batch, p.nrows), C::NT, C::SMEM, st>>>(p.t, p.fft3_perm, batch, p.G, grid, gs_b, gs_lat, gs_lon);
    return 0;
}
template <class C>
static int launch_ana3s2(const Plan& p, int batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st) {
    static PerDeviceOnce attr;
    if (attr.first()) { SHB_CUDA(cudaFuncSetAttribute(fft3::k_lon_ana_fft3s2<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM)); }
    trace_kernel(p, "k_lon_ana_fft3s2");
    // every field of the plan's width is written (zeros beyond the batch): the Legendre stage reads whole column groups
    fft3::k_lon_ana_fft3s2<C><<<dim3(p.Bp, p.nrows), C::NT, C::SMEM, st>>>(p.t, p.fft3_perm, batch, grid, gs_b, gs_lat, gs_lon, p.G);
    return 0;
}

// returns -1 when this path does not apply to the call (the caller then uses lon_fft.cu)
int launch_lon_syn_fft3(const Plan& p, int batch, double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st) {
    if (!p.fft3_perm || gs_b == 1) return -1;
    if (fft3_find_split(p.K)) {
        if (4 * p.nmax >= p.K) return -1;                     // aliasing into the other half's bins: the in-place kernels
        switch (p.K / 2) {
#define X(KK, A, B, C, NF, TPT) case KK: return launch_syn3s2<fft3::Cfg<KK, A, B, C, NF, TPT>>(p, batch, grid, gs_b, gs_lat, gs_lon, st);
            SHB_FFT3S2_CONFIGS(X)
#undef X
            default: return -1;
        }
    }
    if (fft3p_on() && 2 * p.nmax <= p.K && gs_lon == 1) {
        static const int tpt = getenv("SHB200_FFT3P_TPT") ? atoi(getenv("SHB200_FFT3P_TPT")) : 96;
        if (tpt == 128) {
            switch (p.K) {
#define X(KK, A, B, C, NF, TPT) case KK: if (p.t.GS == NF) return launch_syn3p<fft3::Cfg<KK, A, B, C, NF, TPT>>(p, batch, grid, gs_b, gs_lat, gs_lon, st); break;
                SHB_FFT3P_CONFIGS_128(X)
#undef X
                default: break;
            }
        }
        switch (p.K) {
#define X(KK, A, B, C, NF, TPT) case KK: if (p.t.GS == NF) return launch_syn3p<fft3::Cfg<KK, A, B, C, NF, TPT>>(p, batch, grid, gs_b, gs_lat, gs_lon, st); break;
            SHB_FFT3P_CONFIGS(X)
#undef X
            default: break;
        }
    }
    switch (p.K) {
#define X(KK, A, B, C, NF, TPT) case KK: return launch_syn3<fft3::Cfg<KK, A, B, C, NF, TPT>>(p, batch, grid, gs_b, gs_lat, gs_lon, st);
        SHB_FFT3_CONFIGS(X)
#undef X
        default: return -1;
    }
}
int launch_lon_ana_fft3(const Plan& p, int batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st) {
    if (!p.fft3_perm || gs_b == 1) return -1;
    if (fft3_find_split(p.K)) {
        if (4 * p.nmax >= p.K) return -1;
        switch (p.K / 2) {
#define X(KK, A, B, C, NF, TPT) case KK: return launch_ana3s2<fft3::Cfg<KK, A, B, C, NF, TPT>>(p, batch, grid, gs_b, gs_lat, gs_lon, st);
            SHB_FFT3S2_CONFIGS(X)
#undef X
            default: return -1;
        }
    }
    // (one CTA of 4 x 96 threads per SM instead of two of 4 x 64, the shape that helped the persistent synthesis kernel, is slower
    // here: 0.311 -> 0.367 ms at cfg3 -- the two CTAs hide each other's row loads and G stores)
    switch (p.K) {
#define X(KK, A, B, C, NF, TPT) case KK: return launch_ana3<fft3::Cfg<KK, A, B, C, NF, TPT>>(p, batch, grid, gs_b, gs_lat, gs_lon, st);
        SHB_FFT3_CONFIGS(X)
#undef X
        default: return -1;
    }
}

}  // namespace shb

extern "C" int shb200_debug_fft_cycles(int32_t enable, int64_t* out) {
    unsigned long long h[16] = {};
    if (out) {
        SHB_CUDA(cudaDeviceSynchronize());
        SHB_CUDA(cudaMemcpyFromSymbol(h, shb::fft3::g_prof, sizeof h));
        for (int i = 0; i < 16; ++i) out[i] = (int64_t)h[i];
    }
    memset(h, 0, sizeof h);
    SHB_CUDA(cudaMemcpyToSymbol(shb::fft3::g_prof, h, sizeof h));
    const int on = enable ? 1 : 0;
    SHB_CUDA(cudaMemcpyToSymbol(shb::fft3::g_prof_on, &on, sizeof on));
    return 0;
}
