// Legendre stage with a device-resident P_nm table: pure FP64 tensor-core (DMMA) contractions.
//
// Why a table: on sm_100a the DMMA and DFMA instructions share one FP64 datapath per SM sub-partition
// (profiles/fp64_peaks_r01.json: mixed DMMA+DFMA adds up to the same ~37 TFLOP/s).  A producer warp that runs
// the bit-faithful recursion inside the contraction kernel (legendre_dmma.cu) pays a full DMMA slot of pipe
// latency for almost every one of its 7 FP64 instructions per degree and cannot feed 8 consumer warps
// (ncu r01: consumers wait ~40 % of their time on the "P tile full" barrier).  The normalised P_nm depend
// only on (nmax, latitudes) -- not on the data -- so the planner generates them ONCE per plan with the same
// bit-faithful device recursion (k_ptab_gen, one thread per (row, order)) and keeps them in HBM:
//     Ptab[order m][16-degree block][32-row tile][16][34]      (zero padded; 34 = 32 rows + 2 pad doubles)
// i.e. already in the bank-conflict-free padded layout of the shared-memory tiles, so that ONE cp.async.bulk moves a
// whole tile (the first version issued one copy per row and the TMA-issuing warp became the bottleneck: ncu r01,
// consumers 16 % of samples on the "tile full" barrier).  0.87 GB for nmax=720 on 724 latitudes (362 mirror pairs).
// Cm rows carry 2 pad columns for the same reason (internal.h: Cs); G rows stay 128-byte aligned for the longitude
// stage and are fetched row by row (64 copies per 32-row tile, against 8 for its P tiles).  Each transform then streams the table once
// (0.12 ms at HBM speed, overlapped with the contraction) and the kernels below are TMA-fed DMMA GEMMs:
//   synthesis  G[m][row][E/O][col] = sum_n Ptab[n,m][row] * Cm[n,m][col]      CTA = (m, 64 rows, 128 cols)
//   analysis   Cm[n,m][col] = sum_row Ptab[n,m][row] * G[m][row][par(n-m)][col] CTA = (m, 128 cols), 128-degree super-blocks
// One TMA warp (elected lane issues cp.async.bulk row copies into padded shared-memory tiles, mbarrier
// complete_tx) + 8 consumer warps (64 DMMA.8x8x4 per stage each), full/empty mbarrier ring.
// Plans whose table would not fit the budget fall back to the fused on-the-fly kernels.
#include <algorithm>
#include <cstdlib>

#include "internal.h"
#include "recursion.cuh"
#include "dmma_common.cuh"

namespace shb {

constexpr int PT_SUB = 16 * 34;      // doubles per [16 degrees][32 rows + 2] sub-tile

// ------------------------------------------------------------------------------------------------
// table generation: thread <-> (row, order); a warp writes 32 consecutive rows of one sub-tile row (256 B)
// first_k (optional): per (order, 32-row tile) the smallest degree offset n - m at which some row of the tile has |P_nm| >= POLAR_EPS
__global__ void __launch_bounds__(128) k_ptab_gen(DevTables t, double* __restrict__ Ptab, int* __restrict__ first_k) {
    const int m = blockIdx.y;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = t.nmax, R = t.nrows;
    if (row >= R) return;                               // pad rows stay zero (memset)
    const int64_t tmm = tri(m, m, N);
    LegExact rec;
    rec.init(t.cost[row], t.sect[(int64_t)m * R + row], m);
    const double pmm = t.pmm[(int64_t)m * R + row];
    const double* __restrict__ w = t.wnm + tmm;
    const double* __restrict__ ri = t.rinv + tmm;
    double* base = Ptab + (t.pt_base[m] * t.RT32 + (row >> 5)) * PT_SUB + (row & 31);
    int kfirst = 0x7f7f7f7f;
    for (int k = 0; k <= N - m; ++k) {
        const double v = rec.next(w, ri, pmm);
        base[(int64_t)(k >> 4) * t.RT32 * PT_SUB + (k & 15) * 34] = v;
        if (kfirst == 0x7f7f7f7f && fabs(v) >= POLAR_EPS) kfirst = k;
    }
    if (first_k && kfirst != 0x7f7f7f7f) atomicMin(&first_k[m * t.RT32 + (row >> 5)], kfirst);
}

void launch_ptab_gen(const Plan& p, int* first_k, cudaStream_t st) {
    dim3 grid((p.nrows + 127) / 128, p.nmax + 1);
    k_ptab_gen<<<grid, 128, 0, st>>>(p.t, p.Ptab, first_k);
}

// ------------------------------------------------------------------------------------------------
// Tile shapes: the 8 consumer warps form NWR x NWC blocks of 32 rows x 32 columns.  NWC = 4 (128 columns, 64 rows) for
// plans with more than 64 columns; NWC = 2 / 1 (64 / 32 columns, 128 / 256 rows) for narrower plans, so that 16 or 32
// fields do not pay for 128 columns (r01: 16 fields at nmax 2160 cost the same 18.5 ms as 64).
template <int NWC>
struct TS {
    static constexpr int NWR = 8 / NWC, TR = 32 * NWR, CH = 16, NCOLS = 32 * NWC;
    static constexpr int CROW = NCOLS + 2;
    static constexpr int PSZ = NWR * PT_SUB;          // NWR consecutive 32-row sub-tiles [16][34] of the table: one bulk copy
    static constexpr int CSZ = CH * CROW;             // [16][crow], crow = Cs (single column tile, one bulk copy) or NCOLS + 2 (row copies)
    static constexpr int STAGE = PSZ + CSZ;           // doubles
    static constexpr int STAGES = NWC == 1 ? 5 : 8;
    static constexpr int NCONS = 256, NTHREADS = NCONS + 32;
    static constexpr size_t SMEM = sizeof(double) * STAGE * STAGES + 2 * STAGES * sizeof(uint64_t);
};

// Persistent: gridDim.x CTAs (one per SM) walk the (order, row-tile) list in snake order -- tile k*P + c on even
// rounds, k*P + P-1-c on odd rounds; orders ascend, so work descends and every CTA gets the same total -- and the TMA
// warp keeps filling the ring across tile boundaries, so the tensor pipe does not drain 29 times per SM per launch.
template <int NWC>
__global__ void __launch_bounds__(TS<NWC>::NTHREADS, 1) k_leg_syn_tab(DevTables t, const double* __restrict__ Ptab, const double* __restrict__ Cm,
                                                                       double* __restrict__ G, int nrt, int ntiles, const int2* __restrict__ order) {
    using C_ = TS<NWC>;
    constexpr int NWR = C_::NWR, TR = C_::TR, CH = C_::CH, NCOLS = C_::NCOLS, CROW = C_::CROW, STAGES = C_::STAGES, NCONS = C_::NCONS, NTHREADS = C_::NTHREADS;
    extern __shared__ __align__(128) unsigned char smraw[];
    double* stg = reinterpret_cast<double*>(smraw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smraw + sizeof(double) * C_::STAGE * STAGES);
    uint64_t* empty = full + STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int col0 = blockIdx.y * NCOLS;
    const int N = t.nmax, R = t.nrows, NC = t.NC;
    const int ncols = min(NCOLS, NC - col0);
    const bool onecopy = t.Cs <= CROW;                 // whole (padded) rows of Cm fit the tile: one contiguous block, stride Cs
    const int crow = onecopy ? t.Cs : CROW;
    const int P = gridDim.x, c = blockIdx.x;

    // no zero fill: degrees beyond nmax have P = 0 in the table and the Cm rows fetched with them are finite
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCONS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (!onecopy) {      // pad columns / rows of a partial column tile are never written by the row copies
        for (int i = tid; i < C_::STAGE * STAGES; i += NTHREADS) stg[i] = 0.0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == NCONS / 32) {
        // ------------------------------------------------------------------ TMA warp
        int gc = 0;                                     // chunks issued so far (ring position)
        for (int rnd = 0;; ++rnd) {
            const int ti = rnd * P + ((rnd & 1) ? P - 1 - c : c);
            if (ti >= ntiles) { if (rnd * P >= ntiles) break; else continue; }
            const int2 td = order ? __ldg(&order[ti]) : make_int2(ti, 0);
            const int tile = td.x, ch0 = td.y;
            const int m = tile / nrt, rt = tile - m * nrt;
            const int nchunks = (N - m + CH) / CH;
            const int64_t tmm = tri(m, m, N);
            // sub-tiles past the last row tile of the order belong to the next degree block (or the slack behind the
            // table): fetched with the rest, never used
            const double* psrc = Ptab + (t.pt_base[m] * t.RT32 + NWR * rt) * PT_SUB;
            for (int ch = ch0; ch < nchunks; ++ch, ++gc) {
                const int s = gc % STAGES;
                mbar_wait(&empty[s], ((gc / STAGES) & 1) ^ 1);
                const int k0 = ch * CH;
                double* sp = stg + s * C_::STAGE;
                if (onecopy) {
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&full[s], (uint32_t)((C_::PSZ + CH * crow) * 8));
                        bulk_g2s(sp, psrc + (int64_t)ch * t.RT32 * PT_SUB, C_::PSZ * 8, &full[s]);
                        bulk_g2s(sp + C_::PSZ, Cm + (tmm + k0) * t.Cs, (uint32_t)(CH * crow * 8), &full[s]);
                    }
                } else {
                    const int valid = min(CH, N - m - k0 + 1);
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&full[s], (uint32_t)(C_::PSZ * 8 + valid * ncols * 8));
                        bulk_g2s(sp, psrc + (int64_t)ch * t.RT32 * PT_SUB, C_::PSZ * 8, &full[s]);
                        for (int j = 0; j < valid; ++j)
                            bulk_g2s(sp + C_::PSZ + j * crow, Cm + (tmm + k0 + j) * t.Cs + col0, (uint32_t)(ncols * 8), &full[s]);
                    }
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ consumers
        const int wr = warp / NWC, wc = warp % NWC;
        const int lr = lane >> 2, lk = lane & 3;
        int gc = 0;
        int prev = -1;      // stage whose fragments were consumed last; released only after the NEXT stage has been acquired
        // tile descriptors are fetched one round ahead: the load of round rnd + 1 is in flight while round rnd is computed
        auto tile_of = [&](int rnd) { return rnd * P + ((rnd & 1) ? P - 1 - c : c); };
        auto fetch = [&](int rnd) {
            const int ti = tile_of(rnd);
            return (order && ti < ntiles) ? __ldg(&order[ti]) : make_int2(ti, 0);
        };
        int2 td_next = fetch(0);
        for (int rnd = 0;; ++rnd) {
            const int ti = tile_of(rnd);
            const int2 td = td_next;
            td_next = fetch(rnd + 1);
            if (ti >= ntiles) { if (rnd * P >= ntiles) break; else continue; }
            const int tile = td.x, ch0 = td.y;
            const int m = tile / nrt, rt = tile - m * nrt;
            const int nchunks = (N - m + CH) / CH;
            const int row0 = rt * TR;
            // 8-row groups of this warp that hold latitude rows: 4 except in the last row tile (cfg3: 362 row pairs = 5 tiles of
            // 64 + 42 -> the second warp row of the last tile owns 10 rows).  0: nothing to do; <= 2: the half-width body below
            // (a predicated-off DMMA still takes its issue slot, r01; a separate straight-line body does not)
            const int ng = min(4, max(0, (R - row0 - 32 * wr + 7) >> 3));
            double acc[2][4][4][2];
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int g = 0; g < 4; ++g)
#pragma unroll
                    for (int cb = 0; cb < 4; ++cb) { acc[e][g][cb][0] = 0.0; acc[e][g][cb][1] = 0.0; }
            for (int ch = ch0; ch < nchunks; ++ch, ++gc) {
                const int s = gc % STAGES;
                mbar_wait(&full[s], (gc / STAGES) & 1);
                // Release of the previous stage is deferred to here: ptxas is free to schedule an mbarrier arrive ahead of
                // the last DMMAs of a stage, i.e. while their LDS operands are still in flight (observed: the TMA refill
                // then raced with the slowest warps' last fragment loads).  After this wait, every load of stage `prev`
                // has been consumed by DMMAs issued in the previous loop iteration.
                if (prev >= 0) { __syncwarp(); if (lane == 0) mbar_arrive(&empty[prev]); }
                prev = s;
                if (ng == 0) continue;
                const double* __restrict__ SP = stg + s * C_::STAGE + wr * PT_SUB + lr;
                const double* __restrict__ SC = stg + s * C_::STAGE + C_::PSZ + 32 * wc + lr;
                if (ng > 2) {
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int j = 8 * ks + 2 * lk + e;
                            double A[4], B[4];
#pragma unroll
                            for (int g = 0; g < 4; ++g) A[g] = SP[j * 34 + 8 * g];
#pragma unroll
                            for (int cb = 0; cb < 4; ++cb) B[cb] = SC[j * crow + 8 * cb];
#pragma unroll
                            for (int g = 0; g < 4; ++g)
#pragma unroll
                                for (int cb = 0; cb < 4; ++cb) dmma(acc[e][g][cb][0], acc[e][g][cb][1], A[g], B[cb]);
                        }
                } else {
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int j = 8 * ks + 2 * lk + e;
                            double A[2], B[4];
#pragma unroll
                            for (int g = 0; g < 2; ++g) A[g] = SP[j * 34 + 8 * g];
#pragma unroll
                            for (int cb = 0; cb < 4; ++cb) B[cb] = SC[j * crow + 8 * cb];
#pragma unroll
                            for (int g = 0; g < 2; ++g)
#pragma unroll
                                for (int cb = 0; cb < 4; ++cb) dmma(acc[e][g][cb][0], acc[e][g][cb][1], A[g], B[cb]);
                        }
                }
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int row = row0 + 32 * wr + 8 * g + lr;
                if (row >= R) continue;
#pragma unroll
                for (int e = 0; e < 2; ++e)
#pragma unroll
                    for (int cb = 0; cb < 4; ++cb) {
                        const int col = col0 + 32 * wc + 8 * cb + 2 * lk;
                        if (col < NC) *reinterpret_cast<double2*>(G + (((int64_t)m * R + row) * 2 + e) * t.Gs + col) = make_double2(acc[e][g][cb][0], acc[e][g][cb][1]);
                    }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Analysis.  The 8 consumer warps are (parity e) x (NCQ column quarters of 32) x (NDG = 4 / NCQ degree groups); each holds
// 64 degrees of its parity x 32 columns of sums in registers.  NCQ = 4: CTA = (m, 128 columns), 128-degree super-blocks;
// NCQ = 2 / 1 for plans of <= 64 / 32 columns: the spare warps take further degrees instead of idle columns, so a
// super-block (= one pass over all latitude rows of G) spans 256 / 512 degrees and a stage 64 / 128.
template <int NCQ>
struct TA {
    static constexpr int NDG = 4 / NCQ;
    // stages per super-block: 2 * SPB accumulator blocks of 8 degrees x 32 columns per warp (192 registers at SPB = 6, which the
    // setmaxnreg reallocation pays for).  Every super-block is one pass over all rows of G: 128 -> 192 degrees per pass took the
    // kernel from 0.849 to 0.831 ms at cfg3 (160: 0.833)
    static constexpr int SPB = 6;
    static constexpr int RT = 32, NS = 32 * NDG, SB = SPB * NS, NCOLS = 32 * NCQ;
    static constexpr int CROW = NCOLS + 2;
    static constexpr int PSTAGES = NCQ == 1 ? 5 : 8;
    // warps 0-7 (two warpgroups) consume, warp 8 feeds them; warps 9-11 only complete the third warpgroup so that it can
    // hand its registers to the consumers (setmaxnreg works on whole warpgroups)
    static constexpr int NCONS = 256, NTHREADS = 384;
    static constexpr int REGS_CONSUMER = 232, REGS_PRODUCER = 40;
    // the register pool is what the CTA was launched with: 384 threads x 168 registers = 8 warps x 232 + 4 warps x 40 exactly.  A
    // producer share of 48 leaves the consumers' setmaxnreg.inc waiting for registers that never come (r02: the kernel hung)
    static_assert(8 * REGS_CONSUMER + 4 * REGS_PRODUCER <= 12 * 168, "setmaxnreg: more registers requested than the CTA owns");
    static constexpr int GSZ = RT * 2 * CROW;         // [32 rows][E/O][grow] doubles
    static constexpr int PSZ = 2 * NDG * PT_SUB;      // 2 NDG 16-degree sub-tiles [16][34]
    static constexpr size_t SMEM = sizeof(double) * (2 * GSZ + PSTAGES * PSZ) + sizeof(uint64_t) * (2 * PSTAGES + 4);
};

template <int NCQ>
__global__ void __launch_bounds__(TA<NCQ>::NTHREADS, 1) k_leg_ana_tab(DevTables t, const double* __restrict__ Ptab, const double* __restrict__ G,
                                                                       double* __restrict__ Cm, int dbg_no_restream) {
    using C_ = TA<NCQ>;
    constexpr int NDG = C_::NDG, RT = C_::RT, NS = C_::NS, SB = C_::SB, NCOLS = C_::NCOLS, PSTAGES = C_::PSTAGES, NCONS = C_::NCONS;
    extern __shared__ __align__(128) unsigned char smraw[];
    double* gt = reinterpret_cast<double*>(smraw);                  // [2][GSZ]
    double* ptile = gt + 2 * C_::GSZ;                               // [PSTAGES][PSZ]
    uint64_t* pfull = reinterpret_cast<uint64_t*>(ptile + PSTAGES * C_::PSZ);
    uint64_t* pempty = pfull + PSTAGES;
    uint64_t* gfull = pempty + PSTAGES;
    uint64_t* gempty = gfull + 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m = t.m0 + t.mstride * blockIdx.x, col0 = blockIdx.y * NCOLS;
    const int N = t.nmax, R = t.nrows, NC = t.NC;
    const int ncols = min(NCOLS, NC - col0);
    const int nsb = (N - m + SB) / SB;
    const int nrt = (R + RT - 1) / RT;
    const int64_t tmm = tri(m, m, N);
    constexpr int grow = C_::CROW;

    if (tid == 0) {
        for (int s = 0; s < PSTAGES; ++s) { mbar_init(&pfull[s], 1); mbar_init(&pempty[s], NCONS / 32); }
        for (int g = 0; g < 2; ++g) { mbar_init(&gfull[g], 1); mbar_init(&gempty[g], NCONS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // Polar skipping: s0[rt] = first stage (NS degrees) of this order that row tile rt has to visit; the stages before it hold
    // only |P_nm| < POLAR_EPS for its 32 rows.  Producer and consumers derive the same schedule from it: a (super-block, row
    // tile) pair whose stages all lie before s0[rt] is not visited at all (no G tile is fetched for it).
    __shared__ int s0_s[256];                          // plans have at most 8192 rows (plan.cu)
    for (int rt = tid; rt < nrt; rt += C_::NTHREADS) s0_s[rt] = t.first_blk ? t.first_blk[m * t.RT32 + rt] / (2 * NDG) : 0;
    __syncthreads();

    // Register reallocation: compiled for 168 registers per thread (12 warps), the accumulators alone take 128 of them and
    // ptxas cannot keep the fragment loads of the next k-step in flight (ncu r01: 20 % of the consumers' samples wait for
    // LDS results, while the same inner loop with small accumulators saturates the pipe, tools/dmma_feed.cu).  The third
    // warpgroup shrinks to 40 registers and the consumers grow to 232.
    if (warp >= NCONS / 32) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C_::REGS_PRODUCER));
        if (warp > NCONS / 32) return;
        // ------------------------------------------------------------------ TMA warp
        const double* pbase = Ptab + t.pt_base[m] * t.RT32 * PT_SUB;
        int sidx = 0, it = 0;
        for (int sb = 0; sb < nsb; ++sb) {
            const int ninner = min(SB / NS, (N - m - sb * SB + NS) / NS);     // stages of this super-block that hold degrees <= nmax
            for (int rt = 0; rt < nrt; ++rt) {
                const int in0 = max(0, s0_s[rt] - sb * (SB / NS));
                if (in0 >= ninner) continue;
                const int gb = it & 1;
                mbar_wait(&gempty[gb], ((it >> 1) & 1) ^ 1);
                ++it;
                if (dbg_no_restream && sb > 0) {
                    // measurement only (SHB200_DEBUG_ANA_NO_RESTREAM=1, wrong results): the G tiles of every super-block after the
                    // first are NOT fetched again -- the time this saves bounds what sharing one G pass between the CTAs of a cluster
                    // (multicast / DSMEM) could gain
                    if (lane == 0) mbar_arrive(&gfull[gb]);
                } else {
                    const int nr = min(RT, R - rt * RT);
                    if (lane == 0) mbar_arrive_expect_tx(&gfull[gb], (uint32_t)(nr * 2 * ncols * 8));
                    __syncwarp();
                    if (lane < nr) {
                        const double* src = G + (((int64_t)m * R + rt * RT + lane) * 2) * t.Gs + col0;
                        bulk_g2s(&gt[gb * C_::GSZ + (lane * 2 + 0) * grow], src, (uint32_t)(ncols * 8), &gfull[gb]);
                        bulk_g2s(&gt[gb * C_::GSZ + (lane * 2 + 1) * grow], src + t.Gs, (uint32_t)(ncols * 8), &gfull[gb]);
                    }
                }
                for (int inner = in0; inner < ninner; ++inner, ++sidx) {
                    const int s = sidx % PSTAGES;
                    mbar_wait(&pempty[s], ((sidx / PSTAGES) & 1) ^ 1);
                    const int k0 = sb * SB + inner * NS;
                    const int nblk = min(2 * NDG, (N - m - k0 + 16) / 16);          // 16-degree blocks of this stage that exist
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&pfull[s], (uint32_t)(nblk * PT_SUB * 8));
                        for (int q = 0; q < nblk; ++q)
                            bulk_g2s(&ptile[s * C_::PSZ + q * PT_SUB], pbase + ((int64_t)(k0 / 16 + q) * t.RT32 + rt) * PT_SUB, PT_SUB * 8, &pfull[s]);
                    }
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ consumers
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C_::REGS_CONSUMER));
        const int e = warp & 1, cq = (warp >> 1) % NCQ, dg = (warp >> 1) / NCQ;
        const int lr = lane >> 2, lk = lane & 3;
        double acc[2 * C_::SPB][4][2];
        int sidx = 0, it = 0;
        int prev_s = -1, prev_g = -1;   // P stage / G buffer consumed last: released after the next one has been acquired (see k_leg_syn_tab)
        for (int sb = 0; sb < nsb; ++sb) {
            const int ninner = min(SB / NS, (N - m - sb * SB + NS) / NS);
#pragma unroll
            for (int nb = 0; nb < 2 * C_::SPB; ++nb)
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) { acc[nb][cb][0] = 0.0; acc[nb][cb][1] = 0.0; }
            for (int rt = 0; rt < nrt; ++rt) {
                const int in0 = max(0, s0_s[rt] - sb * (SB / NS));
                if (in0 >= ninner) continue;
                const int gb = it & 1;
                const int nr = min(RT, R - rt * RT);
                mbar_wait(&gfull[gb], (it >> 1) & 1);
                ++it;
                if (prev_g >= 0) { __syncwarp(); if (lane == 0) mbar_arrive(&gempty[prev_g]); }
                prev_g = gb;
                const double* __restrict__ GT = gt + gb * C_::GSZ + e * grow + 32 * cq + lr;
#pragma unroll
                for (int inner = 0; inner < SB / NS; ++inner) {
                    if (inner >= ninner) break;
                    if (inner < in0) continue;
                    const int s = sidx % PSTAGES;
                    mbar_wait(&pfull[s], (sidx / PSTAGES) & 1);
                    ++sidx;
                    if (prev_s >= 0) { __syncwarp(); if (lane == 0) mbar_arrive(&pempty[prev_s]); }
                    prev_s = s;
                    // degrees of this warp in the stage: two 16-degree blocks starting at kw
                    const int kw = sb * SB + inner * NS + 32 * dg;
                    if (NDG > 1 && kw > N - m) continue;                            // both blocks past nmax: nothing was copied
                    const double* __restrict__ PT = ptile + s * C_::PSZ + 2 * dg * PT_SUB + (2 * lr + e) * 34 + lk;
                    const bool two = (N - m - kw + 1) > 16;                         // second 16-degree block present
                    auto kstep = [&](int ks, bool rowok) {
                        double A[2], B[4];
                        A[0] = PT[4 * ks];
                        A[1] = two ? PT[PT_SUB + 4 * ks] : 0.0;
#pragma unroll
                        for (int cb = 0; cb < 4; ++cb) B[cb] = rowok ? GT[(4 * ks + lk) * 2 * grow + 8 * cb] : 0.0;
#pragma unroll
                        for (int q = 0; q < 2; ++q)
#pragma unroll
                            for (int cb = 0; cb < 4; ++cb) dmma(acc[2 * inner + q][cb][0], acc[2 * inner + q][cb][1], A[q], B[cb]);
                    };
                    if (nr == RT) {
                        // full tile: straight-line schedule of all 8 k-steps
#pragma unroll
                        for (int ks = 0; ks < RT / 4; ++ks) kstep(ks, true);
                    } else {
                        // last tile of the latitude range: only the k-steps (4 rows each) that hold rows; rows past the last
                        // latitude were never copied.  A separate rolled loop: an early exit inside the unrolled loop above cost
                        // the full tiles 9 % (measured)
                        const int kmax = (nr + 3) >> 2;
#pragma unroll 1
                        for (int ks = 0; ks < kmax; ++ks) kstep(ks, 4 * ks + lk < nr);
                    }
                }
            }
#pragma unroll
            for (int nb = 0; nb < 2 * C_::SPB; ++nb) {
                const int k = sb * SB + (nb >> 1) * NS + 32 * dg + 16 * (nb & 1) + 2 * lr + e;
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

template <int NWC>
static bool launch_syn_tab_nwc(const Plan& p, cudaStream_t st) {
    using C_ = TS<NWC>;
    static PerDeviceOnce attr;
    if (attr.first()) {
        if (cudaFuncSetAttribute(k_leg_syn_tab<NWC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_::SMEM) != cudaSuccess) return false;
    }
    const int sms = device_sm_count();
    const int nrt = (p.nrows + C_::TR - 1) / C_::TR, ntiles = nrt * (p.nmax + 1);
    const int ncolt = (p.NC + C_::NCOLS - 1) / C_::NCOLS;
    dim3 grid(std::max(1, std::min(ntiles, sms / std::min(ncolt, sms))), ncolt);
    trace_kernel(p, NWC == 4 ? "k_leg_syn_tab<4>" : (NWC == 2 ? "k_leg_syn_tab<2>" : "k_leg_syn_tab<1>"));
    const int2* order = p.t.first_blk ? p.syn_order[NWC == 4 ? 0 : (NWC == 2 ? 1 : 2)] : nullptr;
    k_leg_syn_tab<NWC><<<grid, C_::NTHREADS, C_::SMEM, st>>>(p.t, p.Ptab, p.Cm, p.G, nrt, ntiles, order);
    return true;
}

bool launch_legendre_syn_tab(const Plan& p, cudaStream_t st) {
    if (!p.Ptab || p.NC < 32) return false;
    if (p.NC <= 32) return launch_syn_tab_nwc<1>(p, st);
    if (p.NC <= 64) return launch_syn_tab_nwc<2>(p, st);
    return launch_syn_tab_nwc<4>(p, st);
}

template <int NCQ>
static bool launch_ana_tab_ncq(const Plan& p, cudaStream_t st) {
    using C_ = TA<NCQ>;
    static PerDeviceOnce attr;
    if (attr.first()) {
        if (cudaFuncSetAttribute(k_leg_ana_tab<NCQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_::SMEM) != cudaSuccess) return false;
    }
    dim3 grid(ana_order_count(p), (p.NC + C_::NCOLS - 1) / C_::NCOLS);
    trace_kernel(p, NCQ == 4 ? "k_leg_ana_tab<4>" : (NCQ == 2 ? "k_leg_ana_tab<2>" : "k_leg_ana_tab<1>"));
    static const int dbg = [] { const char* e = getenv("SHB200_DEBUG_ANA_NO_RESTREAM"); return e ? atoi(e) : 0; }();
    k_leg_ana_tab<NCQ><<<grid, C_::NTHREADS, C_::SMEM, st>>>(p.t, p.Ptab, p.G, p.Cm, dbg);
    return true;
}

bool launch_legendre_ana_tab(const Plan& p, cudaStream_t st) {
    if (!p.Ptab || p.NC < 32 || p.nrows > 8192) return false;      // s0_s[256]: skip schedule of up to 256 row tiles
    if (p.NC <= 32) return launch_ana_tab_ncq<1>(p, st);
    if (p.NC <= 64) return launch_ana_tab_ncq<2>(p, st);
    return launch_ana_tab_ncq<4>(p, st);
}

}  // namespace shb
