// Internal declarations shared by the translation units of libshb200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/shb200.h"

namespace shb {

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define SHB_CUDA(call)                                                      \
    do {                                                                    \
        cudaError_t _e = (call);                                            \
        if (_e != cudaSuccess) return shb::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

// packed m-major position of (n,m): same indexing as the reference's P_nm table (Legendre_nm.hpp:25-29)
__host__ __device__ inline int64_t tri(int n, int m, int nmax) {
    return (int64_t)m * (nmax + 1) - ((int64_t)m * (m + 1)) / 2 + n;
}
inline int64_t tri_size(int nmax) { return (int64_t)(nmax + 1) * (nmax + 2) / 2; }

struct DevTables {
    // Legendre recursion
    const double* wnm;    // [tri_size]   sqrt((2n+1)/(n+m)*(2n-1)/(n-m))       (Legendre_nm.hpp:73-78)
    const double* rinv;   // [tri_size]   correctly rounded 1/wnm                (division by Markstein correction)
    const double* cost;   // [nrows]      sin(lat*pi/180) of the primary row     (Ynm.hpp:109)
    const double* sect;   // [(nmax+1)*nrows] sectorial multiplier of order m    (Legendre_nm.hpp:101,127)
    const double* pmm;    // [(nmax+1)*nrows] P_mm                               (Legendre_nm.hpp:103,129-130)
    const int* row_i1;    // [nrows] latitude index of the primary row
    const int* row_i2;    // [nrows] latitude index of the mirror row or -1
    const double* w1;     // [nrows] analysis weight of the primary row
    const double* w2;     // [nrows] analysis weight of the mirror row (0 if none)
    // longitude stage
    const double2* tw;    // [K] exp(-2 pi i k / K)
    const double2* phase; // [nmax+1] exp(+i m lon0)
    int phase_unit;       // 1: lon0 == 0, every phase factor is exactly 1
    const double* lonr;   // [nlon] lon*pi/180 (dense path; Ynm.hpp:114)
    // nm packing
    const int64_t* slot_src;  // [tri_size*2] index into the user nm list for (slot, cos/sin) or -1
    int nmax, nrows, nlat, nlon, Bp, NC;
    int GS;               // field-group size of the column order (8, or 4 for tiny batches): always a power of two
    int GSH;              // log2(GS): the pack / unpack kernels turn columns into (cos|sin, field) with shifts -- with runtime
                          // divisions by GS they were integer-instruction bound, not memory bound (r02)
    int Cs, Gs;           // row strides (doubles) of Cm (NC + 2) and G (NC).  The 2 pad columns of Cm make every row start
                          // 16 B mod 64 B, so a CONTIGUOUS block of rows lands in shared memory as a bank-conflict-free padded
                          // tile with a single cp.async.bulk (legendre_tab.cu).  G keeps 128-byte aligned rows: the longitude
                          // stage touches it in 128-byte groups and lost 20 % when its rows were padded the same way (r01).
    int m0, mstride;         // analysis order shard: this plan integrates orders m = m0 + i*mstride only (multi-GPU m-sharding)
    const int64_t* pt_base;  // [nmax+2] first 16-degree block of each order in the tiled P table
    int RT32;                // row tiles (of 32) in the P table = Rpad/32
    const int* first_blk;    // [(nmax+1)*RT32] polar skipping: first 16-degree block of order m in which some |P_nm| of the 32-row
                             // tile reaches POLAR_EPS (blocks before it are not visited); null = visit everything
};
constexpr double POLAR_EPS = 1e-24;

// Column of (cos/sin cs, field b) inside a Cm / G row: fields come in groups of GS with the cosine and sine part of a
// group adjacent, [b/GS][cs][b%GS], so that the longitude stage (GS fields per CTA) touches full 128-byte lines.
__host__ __device__ inline int colidx(int cs, int b, int GS) { return (b / GS) * (2 * GS) + cs * GS + (b % GS); }
// the same with the shift, and its inverse: column c -> (cs, b)
__host__ __device__ inline int colidx_sh(int cs, int b, int gsh) { return ((b >> gsh) << (gsh + 1)) + (cs << gsh) + (b & ((1 << gsh) - 1)); }
__host__ __device__ inline void col_to_field(int c, int gsh, int& cs, int& b) { cs = (c >> gsh) & 1; b = ((c >> (gsh + 1)) << gsh) + (c & ((1 << gsh) - 1)); }

struct Plan;

// ---- kernels (launch wrappers; all asynchronous on `st`) --------------------------------------
// pack: user coefficients -> Cm[tri][NC] ; unpack: Cm -> user (full triangle, position n^2+n+m)
// cudaFuncSetAttribute applies to the current device only: launchers remember per device whether they have raised the
// dynamic shared-memory limit of their kernel (one process may drive several GPUs through several plans).
struct PerDeviceOnce {
    bool done[64] = {};
    bool first() {
        int d = 0;
        cudaGetDevice(&d);
        if (d < 0 || d >= 64) return true;
        const bool f = !done[d];
        done[d] = true;
        return f;
    }
};
int device_sm_count();   // multiprocessors of the current device (cached per device)

void launch_pack(const Plan& p, int batch, const double* coef, int64_t cs_b, int64_t cs_nm, cudaStream_t st);
void launch_unpack(const Plan& p, int batch, double* coef, int64_t cs_b, int64_t cs_nm, cudaStream_t st);
void launch_pack_shard(const Plan& p, int batch, double* packed, int64_t ps_b, cudaStream_t st);   // shard.cu
// Legendre stage
void launch_legendre_syn(const Plan& p, cudaStream_t st);   // Cm -> G
bool launch_legendre_ana(const Plan& p, cudaStream_t st);   // G  -> Cm
// longitude stage
int launch_lon_syn(const Plan& p, int batch, double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st);
int launch_lon_ana(const Plan& p, int batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, cudaStream_t st);
bool lon_fft_supported(int K, int Bp);
void lon_fft_perm(int K, int Bp, std::vector<int>& perm);
// three-pass composite-radix kernels (lon_fft3.cu)
bool lon_fft3_plan(int K, int Bp, int* group_size);
int build_syn_seeds(Plan& p, double* visited);      // legendre_dfma.cu
int build_ana_seeds(Plan& p, double* visited);
// fused point-mode synthesis (points.cu)
bool points_fused_wanted(int npoints);
void launch_syn_points(const Plan& p, int batch, double* out, int64_t gs_b, int64_t gs_pt, cudaStream_t st);
void lon_fft3_perm(int K, std::vector<int>& perm);

// CUDA-graph replay of the launch sequence of small (launch-latency-bound) transforms: plan.cu, "graphs".
struct GraphKey {
    int kind = 0, batch = 0, shard_rank = 0, shard_world = 1;
    uint64_t serial[3] = {0, 0, 0};       // plans involved (unique serial numbers, not addresses: a destroyed plan's address is reused)
    const void* ptr[4] = {nullptr, nullptr, nullptr, nullptr};
    int64_t stride[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    bool operator==(const GraphKey& o) const {
        if (kind != o.kind || batch != o.batch || shard_rank != o.shard_rank || shard_world != o.shard_world) return false;
        for (int i = 0; i < 3; ++i) if (serial[i] != o.serial[i]) return false;
        for (int i = 0; i < 4; ++i) if (ptr[i] != o.ptr[i]) return false;
        for (int i = 0; i < 10; ++i) if (stride[i] != o.stride[i]) return false;
        return true;
    }
};
struct GraphEntry {
    GraphKey key;
    cudaGraphExec_t exec = nullptr;
    std::string trace;
    bool bad = false;      // capture failed once: always launched directly
    uint64_t used = 0;     // LRU stamp
};

struct Plan {
    uint64_t serial = 0;
    int cur_batch = 0;                   // fields of the call whose launch sequence is being issued (CallScope)
    std::vector<GraphEntry> graphs;      // owned by the plan the call is "on" (the analysis plan for multiply / apply_mask)
    cudaStream_t cap_stream = nullptr;
    uint64_t graph_clock = 0;
    int device = 0, nmax = 0, maxB = 0, Bp = 0, NC = 0;
    int mode = 0, nlat = 0, nlon = 0, quadrature = 0, flags = 0;
    int nrows = 0;
    int64_t nsh = 0;
    bool full_triangle = true;
    bool lon_fft = false;
    bool dmma = false;
    bool points_fused = false;   // point mode served by k_syn_points (points.cu): no G array
    int K = 0;
    const int* fft_perm = nullptr;   // device [K] digit-reversal table of the in-place FFT
    const int* fft3_perm = nullptr;  // device [K] digit-reversal table of the three-pass kernels (null: length not covered)
    int64_t workspace_bytes = 0;
    DevTables t{};
    // owned device memory
    std::vector<void*> owned;
    double* Cm = nullptr;   // [tri_size (+16 slack rows)][Cs]
    double* G = nullptr;    // [(nmax+1)][nrows][2][Gs] (+64 slack rows)
    double* Ptab = nullptr; // tiled device-resident P_nm table (legendre_tab.cu) or null
    const int* syn_seed_k0 = nullptr; const double2* syn_seed_state = nullptr; int syn_seed_rows = 0;   // polar-skipping seeds of k_leg_syn_dfma (legendre_dfma.cu)
    const int* ana_seed_k0 = nullptr; const double2* ana_seed_state = nullptr; int ana_seed_rows = 0;   // ... of k_leg_ana_one
    double seed_visited_syn = 1.0, seed_visited_ana = 1.0;
    mutable double* ana_scratch = nullptr; mutable size_t ana_scratch_bytes = 0;   // per-tile partial sums of k_leg_ana_small (split mode), grown on first use
    int Rpad = 0;
    // polar skipping (table kernels): tile lists of the persistent synthesis kernel, heaviest tile first, for its three tile
    // shapes (32 * {2, 4, 8} rows); fraction of the (n, m, row) triples visited per shape (synthesis) / stage width (analysis)
    const int2* syn_order[3] = {nullptr, nullptr, nullptr};   // (tile, first chunk) per list position
    double visited_syn[3] = {1.0, 1.0, 1.0}, visited_ana[3] = {1.0, 1.0, 1.0};
    // duplicates in the nm list (rare; Ynm.hpp accepts them, dgemv sums them)
    int64_t ndup = 0;
    const int64_t* dup_slot = nullptr;  // device [ndup] slot*2+cs
    const int64_t* dup_src = nullptr;   // device [ndup]
    const int64_t* fwd_slot = nullptr;  // device [nsh] list index -> slot*2+cs (-1 for repeated entries)
    const int32_t* deg_of_k = nullptr;  // device [nsh] degree of each list entry
    double* dscale[2] = {nullptr, nullptr};   // device [nmax+1] per-degree scales (synthesis input / analysis output)
    bool dscale_on[2] = {false, false};
    // device grids of shb200_multiply (multiply.cu)
    double* mul_grid = nullptr; int64_t mul_grid_elems = 0;
    int launches_syn = 0, launches_ana = 0;
    cudaEvent_t ev[8] = {};     // SHB200_FLAG_TIMING: syn 0..3, ana 4..7
    bool timed_syn = false, timed_ana = false;
    // ---- concurrency (plan.cu: CallScope) ----------------------------------------------------------------------------
    // A plan is shared by every caller of the same transform (the engine caches plans; shxarray engines are process-wide
    // singletons that dask may call from several threads).  What a call changes -- effective width, order shard, degree
    // scale -- is applied under `mu` for the duration of the host-side launch sequence only (kernel arguments are copied at
    // launch) and restored afterwards; the workspace (Cm, G) is handed from one stream to the next through `ev_done`.
    std::mutex mu;
    cudaEvent_t ev_done = nullptr;
    cudaStream_t last_stream = nullptr;
    bool has_last = false;
    // names of the kernels launched by the most recent call (shb200_plan_last_kernels; tests assert the kernel family)
    mutable std::string trace;
    struct HostIO* hio = nullptr;   // host entry points (host_io.cu): streams, staging, events
};

inline void trace_kernel(const Plan& p, const char* name) {
    if (!p.trace.empty()) p.trace += ',';
    p.trace += name;
}

// What one call may change on a plan (C ABI: shb200_call_opts).
struct CallOpts {
    int shard_rank = 0, shard_world = 1;
    const double* degree_scale = nullptr;   // host, nmax+1
    int min_fields = 0;                     // internal: effective width of the call is at least this many fields (host pipeline: 8-field edge chunks)
    bool no_events = false;                 // internal: launch sequence under stream capture (the workspace hand-over events stay outside the graph)
};
// graphs (plan.cu)
bool graph_eligible(const Plan& p, const CallOpts& o);
void drop_graphs(Plan& p);
// Runs `body(stream, opts_no_events)` directly, or -- from the second call with the same key on -- as a captured graph replayed on
// `st`.  `touched`: the plans whose workspaces the sequence uses (hand-over events are handled here when the graph is replayed).
int run_graphed(Plan& owner, const GraphKey& key, cudaStream_t st, Plan* const* touched, int ntouched,
                const std::function<int(cudaStream_t, bool)>& body);
int parse_opts(const Plan& p, const shb200_call_opts* o, CallOpts& out);
// Device-pointer transforms without locking (the caller holds p.mu): used by the public entry points, the host entry
// points and shb200_multiply.
int do_synthesis(Plan& p, int batch, const double* coef, int64_t cs_b, int64_t cs_nm, double* grid, int64_t gs_b,
                 int64_t gs_lat, int64_t gs_lon, cudaStream_t st, const CallOpts& o);
int do_analysis(Plan& p, int batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, double* coef,
                int64_t cs_b, int64_t cs_nm, cudaStream_t st, const CallOpts& o);
void free_host_io(Plan& p);

// number of orders m = m0, m0+mstride, ... <= nmax integrated by this plan's analysis
inline int ana_order_count(const Plan& p) { return p.t.m0 > p.nmax ? 0 : (p.nmax - p.t.m0) / p.t.mstride + 1; }

}  // namespace shb

struct shb200_plan { shb::Plan p; };
