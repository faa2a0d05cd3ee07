// FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) Legendre-stage kernels for sm_100a.
//
// Replaces, for a whole batch at once, the per-point  Ynm.set + dgemv / dger  of the reference
// (synthesis.pyx:175-189, analysis.pyx:125-139): for every order m the sum over degrees is the dense
// contraction  [rows x (nmax-m+1)] . [(nmax-m+1) x 2*batch]  with P_nm generated on the fly.
//
// Synthesis kernel (k_leg_syn_dmma), one CTA = (order m, tile of 64 mirror-paired latitude rows, 128 columns):
//   * 2 producer warps: one thread per row runs the column recursion bit-faithfully (recursion.cuh; values kept in
//     the reference's 1e-280-scaled domain and multiplied by the sectorial factor, 7 FP64 ops per degree,
//     so the tile holds exactly the reference's P_nm) and writes P into a shared-memory
//     ring, 16 degrees per stage; one elected producer thread streams the matching coefficient rows
//     Cm[n][col0..col0+127] with TMA bulk copies (cp.async.bulk, mbarrier complete_tx).
//   * 8 consumer warps: per stage 64 DMMA.8x8x4 each; even and odd (n-m) accumulate separately (E/O), which is
//     what makes one recursion serve a latitude and its mirror.  Accumulators live in registers for the whole
//     degree loop; the epilogue stores G[m][row][E/O][col].
//   full/empty mbarriers per stage; no __syncthreads in the main loop.
// Shared-memory rows are padded (stride = 16 B mod 64 B) so every 8x4 / 4x8 fragment load is conflict-free.
#include "internal.h"
#include "recursion.cuh"
#include "dmma_common.cuh"

namespace shb {
namespace dm {

constexpr int TR = 64;            // rows per CTA
constexpr int CH = 16;            // degrees per pipeline stage
constexpr int NCOLS = 128;        // columns per CTA
constexpr int PROW = TR + 2;      // doubles per P-tile row   (528 B = 16 mod 64)
constexpr int CROW = NCOLS + 2;   // doubles per C-tile row   (1040 B = 16 mod 64)
constexpr int STAGES = 6;
constexpr int NCONS = 256, NPROD = 64, NTHREADS = NCONS + NPROD;

struct __align__(16) Stage {
    double P[CH][PROW];
    double C[CH][CROW];
};
constexpr size_t SMEM_SYN = sizeof(Stage) * STAGES + 2 * STAGES * sizeof(uint64_t);

__global__ void __launch_bounds__(NTHREADS, 1) k_leg_syn_dmma(DevTables t, const double* __restrict__ Cm, double* __restrict__ G) {
    extern __shared__ __align__(128) unsigned char smraw[];
    Stage* st = reinterpret_cast<Stage*>(smraw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smraw + sizeof(Stage) * STAGES);
    uint64_t* empty = full + STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m = blockIdx.y, row0 = blockIdx.x * TR, col0 = blockIdx.z * NCOLS;
    const int N = t.nmax, R = t.nrows, NC = t.NC;
    const int ncols = min(NCOLS, NC - col0);
    const int nchunks = (N - m + CH) / CH;
    const int64_t tmm = tri(m, m, N);

    for (int i = tid; i < (int)(sizeof(Stage) * STAGES / 8); i += NTHREADS) reinterpret_cast<double*>(smraw)[i] = 0.0;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], NPROD); mbar_init(&empty[s], NCONS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if (warp >= NCONS / 32) {
        // ------------------------------------------------------------------ producers
        const int pt = tid - NCONS;
        const int row = row0 + pt;
        const bool live = row < R;
        const double c = t.cost[live ? row : R - 1];
        const double sect = t.sect[(int64_t)m * R + (live ? row : R - 1)];
        const double pmm = t.pmm[(int64_t)m * R + (live ? row : R - 1)];
        const double* __restrict__ w = t.wnm + tmm;
        const double* __restrict__ ri = t.rinv + tmm;
        double p1 = 0.0, p2 = 1e-280;
        for (int ch = 0; ch < nchunks; ++ch) {
            const int s = ch % STAGES;
            const uint32_t ph = (ch / STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            const int k0 = ch * CH;
            const int valid = min(CH, N - m - k0 + 1);
            if (pt == 0) {
                mbar_expect_tx(&full[s], (uint32_t)(valid * ncols * 8));
                const double* src = Cm + (tmm + k0) * t.Cs + col0;
                for (int j = 0; j < valid; ++j) bulk_g2s(&st[s].C[j][0], src + (int64_t)j * t.Cs, (uint32_t)(ncols * 8), &full[s]);
            }
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int k = k0 + j;
                double v = 0.0;
                if (j < valid) {
                    if (k == 0) {
                        v = pmm;                                                   // Legendre_nm.hpp:103,129-130
                    } else if (k == 1) {
                        p1 = __dmul_rn(__dmul_rn(__ldg(w + 1), c), p2);            // Legendre_nm.hpp:111
                        v = __dmul_rn(p1, sect);                                   // :112
                    } else {
                        const double pn = leg_step(c, p1, p2, __ldg(w + k), __ldg(w + k - 1), __ldg(ri + k - 1));
                        p2 = p1; p1 = pn; v = __dmul_rn(pn, sect);                 // :120
                    }
                }
                st[s].P[j][pt] = live ? v : 0.0;
            }
            mbar_arrive(&full[s]);
        }
    } else {
        // ------------------------------------------------------------------ consumers
        const int wr = warp >> 2, wc = warp & 3;
        const int lr = lane >> 2, lk = lane & 3;
        double acc[2][4][4][2];
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) { acc[e][g][cb][0] = 0.0; acc[e][g][cb][1] = 0.0; }
        int prev = -1;   // deferred release: see legendre_tab.cu (an arrive may be scheduled ahead of a stage's last fragment loads)
        for (int ch = 0; ch < nchunks; ++ch) {
            const int s = ch % STAGES;
            const uint32_t ph = (ch / STAGES) & 1;
            mbar_wait(&full[s], ph);
            if (prev >= 0) { __syncwarp(); if (lane == 0) mbar_arrive(&empty[prev]); }
            prev = s;
            const Stage& S = st[s];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = 8 * ks + 2 * lk + e;
                    double A[4], B[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) A[g] = S.P[j][32 * wr + 8 * g + lr];
#pragma unroll
                    for (int cb = 0; cb < 4; ++cb) B[cb] = S.C[j][32 * wc + 8 * cb + lr];
#pragma unroll
                    for (int g = 0; g < 4; ++g)
#pragma unroll
                        for (int cb = 0; cb < 4; ++cb) dmma(acc[e][g][cb][0], acc[e][g][cb][1], A[g], B[cb]);
                }
        }
        // epilogue: store E/O
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int row = row0 + 32 * wr + 8 * g + lr;
            if (row >= R) continue;
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) {
                    const int col = col0 + 32 * wc + 8 * cb + 2 * lk;
                    if (col < NC) {
                        double2 v = make_double2(acc[e][g][cb][0], acc[e][g][cb][1]);
                        *reinterpret_cast<double2*>(G + (((int64_t)m * R + row) * 2 + e) * t.Gs + col) = v;
                    }
                }
        }
    }
}

}  // namespace dm

// ------------------------------------------------------------------------------------------------
// Analysis kernel (k_leg_ana_dmma), one CTA = (order m, 128 columns):
//   C[n][col] = sum_rows P_nm(row) * G_parity(n-m)[row][col]            (the dger of analysis.pyx:136, batched)
//   * degrees are processed in super-blocks of 128 (accumulators 128 x 128 doubles in the consumers' registers);
//     within a super-block the reduction dimension (latitude rows) is streamed in tiles of 32 rows:
//     the tile's G rows (E and O parts) arrive by TMA bulk copies into one of two buffers, and a producer warp
//     generates P for 32 degrees x 32 rows per ring stage, carrying its recursion state in shared memory from
//     one super-block to the next (so every P_nm is generated exactly once).
//   * M = 8 degrees of equal parity, N = 8 columns, K = 4 rows per DMMA; consumer warp (e, q) owns parity e and
//     column quarter q for all 8 degree blocks of the super-block.
namespace ana {
constexpr int RT = 32;            // rows per tile (K extent of one G buffer)
constexpr int NS = 32;            // degrees per ring stage
constexpr int SB = 128;           // degrees per super-block
constexpr int NCOLS = 128;
constexpr int PROW = RT + 2;      // 272 B = 16 mod 64
constexpr int CROW = NCOLS + 2;
constexpr int PSTAGES = 8;       // = 2 producer warps x SB/NS stages: row tile `it` uses stages (it&1)*4 + inner, so every stage barrier
                                 // has ONE producer warp.  With a shared 4-stage ring the warp that finished tile `it` could reach the
                                 // "stage empty" wait of tile it+2 while the barrier was still in phase `it`; a parity wait cannot tell
                                 // phase it from it+2 and let it through (intermittent corruption / launch failure, found by stress runs).
constexpr int NCONS = 256, NPROD = 64, NTHREADS = NCONS + NPROD;
constexpr int MAXROWS = 1280;     // recursion state kept in shared memory

struct __align__(16) GTile { double g[RT][2][CROW]; };
struct __align__(16) PTile { double p[NS][PROW]; };
constexpr size_t smem_bytes(int rows_padded) {
    return sizeof(GTile) * 2 + sizeof(PTile) * PSTAGES + sizeof(double) * 2 * (size_t)rows_padded + sizeof(uint64_t) * (2 * PSTAGES + 4);
}
}  // namespace ana

__global__ void __launch_bounds__(ana::NTHREADS, 1) k_leg_ana_dmma(DevTables t, const double* __restrict__ G, double* __restrict__ Cm, int rows_padded) {
    using namespace ana;
    extern __shared__ __align__(128) unsigned char smraw[];
    GTile* gt = reinterpret_cast<GTile*>(smraw);
    PTile* ptile = reinterpret_cast<PTile*>(smraw + sizeof(GTile) * 2);
    double* sp1 = reinterpret_cast<double*>(smraw + sizeof(GTile) * 2 + sizeof(PTile) * PSTAGES);
    double* sp2 = sp1 + rows_padded;
    uint64_t* pfull = reinterpret_cast<uint64_t*>(sp2 + rows_padded);
    uint64_t* pempty = pfull + PSTAGES;
    uint64_t* gfull = pempty + PSTAGES;
    uint64_t* gempty = gfull + 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m = t.m0 + t.mstride * blockIdx.x, col0 = blockIdx.y * NCOLS;
    const int N = t.nmax, R = t.nrows, NC = t.NC;
    const int ncols = min(NCOLS, NC - col0);
    const int nsb = (N - m + SB) / SB;                 // super-blocks
    const int nrt = (R + RT - 1) / RT;                 // row tiles
    const int64_t tmm = tri(m, m, N);

    for (int i = tid; i < (int)((sizeof(GTile) * 2 + sizeof(PTile) * PSTAGES) / 8); i += NTHREADS) reinterpret_cast<double*>(smraw)[i] = 0.0;
    if (tid == 0) {
        for (int s = 0; s < PSTAGES; ++s) { mbar_init(&pfull[s], 32); mbar_init(&pempty[s], NCONS / 32); }
        for (int g = 0; g < 2; ++g) { mbar_init(&gfull[g], 1); mbar_init(&gempty[g], NCONS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if (warp >= NCONS / 32) {
        // ------------------------------------------------------------------ producers (warp pw handles row tiles rt = pw mod 2)
        const int pw = warp - NCONS / 32;
        const double* __restrict__ w = t.wnm + tmm;
        const double* __restrict__ ri = t.rinv + tmm;
        for (int sb = 0; sb < nsb; ++sb) {
            for (int rt = pw; rt < nrt; rt += 2) {
                const int it = sb * nrt + rt;                 // global row-tile counter (consumers see the same order)
                const int gb = it & 1;
                const int row = rt * RT + lane;
                const bool live = row < R;
                const int rr = live ? row : R - 1;
                if (lane == 0) {
                    mbar_wait(&gempty[gb], ((it >> 1) & 1) ^ 1);
                    const int nr = min(RT, R - rt * RT);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&gfull[gb])), "r"((uint32_t)(nr * 2 * ncols * 8)) : "memory");
                    const double* src = G + (((int64_t)m * R + rt * RT) * 2) * t.Gs + col0;
                    for (int j = 0; j < nr * 2; ++j) bulk_g2s(&gt[gb].g[j >> 1][j & 1][0], src + (int64_t)j * t.Gs, (uint32_t)(ncols * 8), &gfull[gb]);
                }
                const double c = t.cost[rr];
                const double sect = t.sect[(int64_t)m * R + rr];
                const double pmm = t.pmm[(int64_t)m * R + rr];
                double p1 = 0.0, p2 = 1e-280;
                if (sb > 0) { p1 = sp1[row]; p2 = sp2[row]; }
                for (int inner = 0; inner < SB / NS; ++inner) {
                    const int sidx = it * (SB / NS) + inner;      // global stage counter
                    const int s = sidx % PSTAGES;
                    mbar_wait(&pempty[s], ((sidx / PSTAGES) & 1) ^ 1);
                    const int k0 = sb * SB + inner * NS;
                    const int valid = min(NS, N - m - k0 + 1);
#pragma unroll 8
                    for (int j = 0; j < NS; ++j) {
                        const int k = k0 + j;
                        double v = 0.0;
                        if (j < valid) {
                            if (k == 0) {
                                v = pmm;
                            } else if (k == 1) {
                                p1 = __dmul_rn(__dmul_rn(__ldg(w + 1), c), p2);
                                v = __dmul_rn(p1, sect);
                            } else {
                                const double pn = leg_step(c, p1, p2, __ldg(w + k), __ldg(w + k - 1), __ldg(ri + k - 1));
                                p2 = p1; p1 = pn; v = __dmul_rn(pn, sect);
                            }
                        }
                        ptile[s].p[j][lane] = live ? v : 0.0;
                    }
                    mbar_arrive(&pfull[s]);
                }
                sp1[row] = p1; sp2[row] = p2;
            }
        }
    } else {
        // ------------------------------------------------------------------ consumers
        const int e = warp & 1, cq = warp >> 1;
        const int lr = lane >> 2, lk = lane & 3;
        double acc[8][4][2];
        int prev_s = -1, prev_g = -1;   // deferred release (see legendre_tab.cu)
        for (int sb = 0; sb < nsb; ++sb) {
#pragma unroll
            for (int nb = 0; nb < 8; ++nb)
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) { acc[nb][cb][0] = 0.0; acc[nb][cb][1] = 0.0; }
            for (int rt = 0; rt < nrt; ++rt) {
                const int it = sb * nrt + rt;
                const int gb = it & 1;
                mbar_wait(&gfull[gb], (it >> 1) & 1);
                if (prev_g >= 0) { __syncwarp(); if (lane == 0) mbar_arrive(&gempty[prev_g]); }
                prev_g = gb;
                const GTile& GT = gt[gb];
#pragma unroll
                for (int inner = 0; inner < SB / NS; ++inner) {
                    const int sidx = it * (SB / NS) + inner;
                    const int s = sidx % PSTAGES;
                    mbar_wait(&pfull[s], (sidx / PSTAGES) & 1);
                    if (prev_s >= 0) { __syncwarp(); if (lane == 0) mbar_arrive(&pempty[prev_s]); }
                    prev_s = s;
                    const PTile& PT = ptile[s];
#pragma unroll
                    for (int ks = 0; ks < RT / 4; ++ks) {
                        double A[2], B[4];
#pragma unroll
                        for (int q = 0; q < 2; ++q) A[q] = PT.p[16 * q + 2 * lr + e][4 * ks + lk];
#pragma unroll
                        for (int cb = 0; cb < 4; ++cb) B[cb] = GT.g[4 * ks + lk][e][32 * cq + 8 * cb + lr];
#pragma unroll
                        for (int q = 0; q < 2; ++q)
#pragma unroll
                            for (int cb = 0; cb < 4; ++cb) dmma(acc[2 * inner + q][cb][0], acc[2 * inner + q][cb][1], A[q], B[cb]);
                    }
                }
            }
            // store the finished super-block
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const int k = sb * SB + 16 * nb + 2 * lr + e;          // degree offset n - m
                if (k > N - m) continue;
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) {
                    const int col = col0 + 32 * cq + 8 * cb + 2 * lk;
                    if (col < NC) *reinterpret_cast<double2*>(Cm + (tmm + k) * t.Cs + col) = make_double2(acc[nb][cb][0], acc[nb][cb][1]);
                }
            }
        }
    }
}

bool launch_legendre_syn_dmma(const Plan& p, cudaStream_t st) {
    if (p.NC < 32) return false;   // tiny batches: the one-row-per-thread DFMA kernel is the better shape
    static PerDeviceOnce attr;
    if (attr.first()) {
        if (cudaFuncSetAttribute(dm::k_leg_syn_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dm::SMEM_SYN) != cudaSuccess) return false;
    }
    dim3 grid((p.nrows + dm::TR - 1) / dm::TR, p.nmax + 1, (p.NC + dm::NCOLS - 1) / dm::NCOLS);
    trace_kernel(p, "k_leg_syn_dmma");
    dm::k_leg_syn_dmma<<<grid, dm::NTHREADS, dm::SMEM_SYN, st>>>(p.t, p.Cm, p.G);
    return true;
}

bool launch_legendre_ana_dmma(const Plan& p, cudaStream_t st) {
    if (p.NC < 32 || p.nrows > ana::MAXROWS) return false;
    const int rows_padded = (p.nrows + ana::RT - 1) / ana::RT * ana::RT;
    const size_t smem = ana::smem_bytes(rows_padded);
    static PerDeviceOnce attr;
    if (attr.first()) {
        if (cudaFuncSetAttribute(k_leg_ana_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ana::smem_bytes(ana::MAXROWS)) != cudaSuccess) return false;
    }
    dim3 grid(ana_order_count(p), (p.NC + ana::NCOLS - 1) / ana::NCOLS);
    trace_kernel(p, "k_leg_ana_dmma");
    k_leg_ana_dmma<<<grid, ana::NTHREADS, smem, st>>>(p.t, p.G, p.Cm, rows_padded);
    return true;
}

}  // namespace shb
