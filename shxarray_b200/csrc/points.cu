// Point-mode synthesis in ONE kernel: f(b, point) = sum_m [ cos(m lon) sum_n P_nm(lat) C_nm(b) + sin(m lon) sum_n P_nm(lat) S_nm(b) ]
//
// This is the reference's per-point loop (synthesis.pyx:186-189: Ynm.set(lon, lat) followed by one dgemv per point) with the
// roles turned round: thread <-> point, the recursion of P_nm runs in registers (bit-faithful, recursion.cuh), the
// coefficients of 32 degrees at a time are staged in shared memory and read back as warp-wide broadcasts, and the trig
// factor of an order is applied once per (point, order) exactly as Ynm.hpp:114,127 evaluates it (sincos(m * lonr)).
// The generic path (Legendre stage with one "latitude row" per point + dense longitude stage) materialises
// G[m][point][E/O][column]: 15.4 GB for 1e5 points at nmax 200 x 20 fields, and runs the tensor-core kernels with their mirror
// (E/O) machinery on rows that have no mirror (r02: 37 ms for that case).  Points have no structure to share, so the work is
// one FMA per (n, m, point, column) plus 7 FP64 operations per (n, m, point) of recursion per column block: FP64-pipe bound.
#include <cstdlib>

#include "internal.h"
#include "recursion.cuh"

namespace shb {

constexpr int PS_CH = 32;          // degrees per staged chunk
constexpr int PS_NT = 128;         // points per CTA

// NCOL consecutive columns of Cm per CTA (column = [field group][cos | sin][field in group], groups of GS fields); PPT points per
// thread: every coefficient read from shared memory (a warp-wide broadcast, one LSU wavefront) then feeds PPT x 2 DFMAs.  With
// one point per thread and 32 columns the kernel issued 16 LDS.128 per 39 FP64 instructions and the LSU was as busy as the FP64
// pipe (ncu r02: fp64 pipe 51 %); two points x 16 columns read 8 per 46.
template <int NCOL, int GS, int MINB, int PPT>
__global__ void __launch_bounds__(PS_NT, MINB) k_syn_points(DevTables t, int batch, int colbase, const double* __restrict__ Cm, double* __restrict__ out,
                                                            int64_t gs_b, int64_t gs_pt) {
    static_assert(NCOL % (2 * GS) == 0, "whole field groups per column block");
    constexpr int NF = NCOL / 2;                       // fields of the block
    constexpr int PER = PS_CH * NCOL / PS_NT;          // staged coefficients per thread and chunk
    __shared__ __align__(16) double sc[2][PS_CH][NCOL];
    __shared__ double sw[2][PS_CH + 1], sr[2][PS_CH + 1];   // wnm(m + k, m), RN(1 / wnm) for k = k0 - 1 .. k0 + 31
    const int N = t.nmax, R = t.nrows;
    const int tid = threadIdx.x;
    const int col0 = colbase + blockIdx.y * NCOL;
    bool live[PPT];
    int r[PPT], ipt[PPT];
    double cost[PPT], lonr[PPT], f[PPT][NF];
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        const int row = (blockIdx.x * PPT + q) * PS_NT + tid;
        live[q] = row < R;
        r[q] = live[q] ? row : R - 1;
        cost[q] = t.cost[r[q]];
        ipt[q] = t.row_i1[r[q]];
        lonr[q] = t.lonr[ipt[q]];
#pragma unroll
        for (int j = 0; j < NF; ++j) f[q][j] = 0.0;
    }

    for (int m = 0; m <= N; ++m) {
        const int64_t tmm = tri(m, m, N);
        const double* __restrict__ w = t.wnm + tmm;
        const double* __restrict__ ri = t.rinv + tmm;
        const int nch = (N - m + PS_CH) / PS_CH;
        LegExact rec[PPT];
        double pmm[PPT], ac[PPT][NCOL];
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            rec[q].init(cost[q], t.sect[(int64_t)m * R + r[q]], m);
            pmm[q] = t.pmm[(int64_t)m * R + r[q]];
#pragma unroll
            for (int c = 0; c < NCOL; ++c) ac[q][c] = 0.0;
        }

        // chunk `ch` of this order: coefficient rows straight into the other shared-memory buffer with cp.async (8-byte pieces: Cm
        // rows are only 8-byte aligned for odd column offsets; no staging registers), table entries through two registers
        double wv = 0.0, rv = 0.0;
        auto fetch = [&](int ch, int buf) {
            const int k0 = ch * PS_CH;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int idx = tid + PS_NT * i, nn = idx / NCOL, c = idx - nn * NCOL;
                if (k0 + nn <= N - m) {       // rows past nmax are never read by the loop below
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(&sc[buf][nn][c]);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(Cm + (tmm + k0 + nn) * t.Cs + col0 + c) : "memory");
                }
            }
            if (tid <= PS_CH) {
                const int k = k0 + tid - 1;
                const bool ok = k >= 0 && m + k <= N;
                wv = ok ? w[k] : 0.0;                      // entries past nmax are zero: the recursion then yields exact zeros
                rv = ok ? ri[k] : 0.0;
            }
        };
        auto stash = [&](int buf) {
            if (tid <= PS_CH) { sw[buf][tid] = wv; sr[buf][tid] = rv; }
            asm volatile("cp.async.wait_all;" ::: "memory");
        };
        fetch(0, 0);
        stash(0);
        __syncthreads();
        for (int ch = 0; ch < nch; ++ch) {
            const int buf = ch & 1;
            if (ch + 1 < nch) fetch(ch + 1, buf ^ 1);
            const int nvalid = min(PS_CH, N - m - ch * PS_CH + 1);
            const double* __restrict__ cw = sw[buf] + 1;
            const double* __restrict__ cr = sr[buf] + 1;
            auto body = [&](int nn, const double (&P)[PPT]) {
#pragma unroll
                for (int c = 0; c < NCOL; c += 2) {
                    const double2 cc = *reinterpret_cast<const double2*>(&sc[buf][nn][c]);
#pragma unroll
                    for (int q = 0; q < PPT; ++q) {
                        ac[q][c] = fma(P[q], cc.x, ac[q][c]);
                        ac[q][c + 1] = fma(P[q], cc.y, ac[q][c + 1]);
                    }
                }
            };
            if (ch == 0) {
                for (int nn = 0; nn < nvalid; ++nn) {
                    double P[PPT];
#pragma unroll
                    for (int q = 0; q < PPT; ++q) P[q] = rec[q].next_v(cw[nn], cw[nn - 1], cr[nn - 1], pmm[q]);
                    body(nn, P);
                }
            } else {
#pragma unroll 4
                for (int nn = 0; nn < nvalid; ++nn) {
                    double P[PPT];
#pragma unroll
                    for (int q = 0; q < PPT; ++q) {
                        const double pn = leg_step(rec[q].c, rec[q].p1, rec[q].p2, cw[nn], cw[nn - 1], cr[nn - 1]);
                        rec[q].p2 = rec[q].p1; rec[q].p1 = pn;
                        P[q] = __dmul_rn(pn, rec[q].sect);
                    }
                    body(nn, P);
                }
            }
            if (ch + 1 < nch) stash(buf ^ 1);
            __syncthreads();
        }
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            double sn, cs;
            sincos((double)m * lonr[q], &sn, &cs);           // Ynm.hpp:114,127
#pragma unroll
            for (int j = 0; j < NF; ++j) {
                const int g = j / GS, jj = j - g * GS;
                f[q][j] += ac[q][g * 2 * GS + jj] * cs + ac[q][g * 2 * GS + GS + jj] * sn;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        if (!live[q]) continue;
#pragma unroll
        for (int j = 0; j < NF; ++j) {
            const int b = (col0 / (2 * GS)) * GS + j;        // col0 is a multiple of 2 GS
            if (b < batch) out[(int64_t)b * gs_b + (int64_t)ipt[q] * gs_pt] = f[q][j];
        }
    }
}

// Plans for point sets with enough points to fill the GPU one thread per point use the fused kernel (and own no G array);
// small point sets keep the generic path, which also spreads the orders over CTAs.  SHB200_POINTS_FUSED = 0 / 1 overrides.
bool points_fused_wanted(int npoints) {
    const char* e = getenv("SHB200_POINTS_FUSED");
    if (e) return atoi(e) != 0;
    return npoints >= 16384;      // r02: 1e4 points at nmax 720 x 4 fields: 21.5 ms fused (79 CTAs) against 4.8 ms generic; 2e4 points: 5.8 against 9.0 ms
}

void launch_syn_points(const Plan& p, int batch, double* out, int64_t gs_b, int64_t gs_pt, cudaStream_t st) {
    trace_kernel(p, "k_syn_points");
    const int nblk1 = (p.nrows + PS_NT - 1) / PS_NT, nblk2 = (p.nrows + 2 * PS_NT - 1) / (2 * PS_NT);
    if (p.t.GS == 4) {                                       // tiny batches: NC = 8 (one group of 4 fields)
        for (int c = 0; c < p.NC; c += 8) k_syn_points<8, 4, 4, 1><<<dim3(nblk1, 1), PS_NT, 0, st>>>(p.t, batch, c, p.Cm, out, gs_b, gs_pt);
        return;
    }
    // SHB200_POINTS_SHAPE=1: r02a shape (one point per thread, blocks of 32 columns + a 16-column remainder)
    static const int shape = [] { const char* e = getenv("SHB200_POINTS_SHAPE"); return e ? atoi(e) : 2; }();
    if (shape == 1) {
        const int n32 = p.NC / 32, rem = p.NC - 32 * n32;    // NC is a multiple of 16
        if (n32) k_syn_points<32, 8, 3, 1><<<dim3(nblk1, n32), PS_NT, 0, st>>>(p.t, batch, 0, p.Cm, out, gs_b, gs_pt);
        if (rem) k_syn_points<16, 8, 3, 1><<<dim3(nblk1, 1), PS_NT, 0, st>>>(p.t, batch, 32 * n32, p.Cm, out, gs_b, gs_pt);
        return;
    }
    k_syn_points<16, 8, 3, 2><<<dim3(nblk2, p.NC / 16), PS_NT, 0, st>>>(p.t, batch, 0, p.Cm, out, gs_b, gs_pt);
}

}  // namespace shb
