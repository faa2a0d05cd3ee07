// Multi-GPU exchange formats: what one rank hands to the collective and how the gathered shards become
// the caller's array.  The reference has no distributed path (one process, prange over latitude rows: synthesis.pyx:179,
// analysis.pyx:128); these kernels sit either side of the ONE collective of a sharded transform.
//
//  m-order-sharded analysis (config 4 of BASELINE.json, what north_star names): rank r of W integrates the orders
//  m = r, r+W, r+2W, ... and emits ONLY those, packed m-major:
//        packed[b][ off_j + 2 (n - m_j) + cs ],   m_j = r + j W,   off_j = 2 [ j (N + 1 - r) - W j (j - 1) / 2 ],   cs = 0 cos / 1 sin
//  (4.6 MB per rank at nmax 2160 instead of the zero-padded 37 MB vector); ncclAllGather of the W shards; k_unpack_shards
//  writes the n-major nm vector (position n^2+n+m: analysis.pyx:29, sh_indexing.py:66) from the gathered buffer.
//
//  latitude-sharded synthesis: every rank synthesises its mirror-paired rows into its slot of a [W][B][Rmax][K] buffer;
//  ncclAllGather; k_scatter_rows moves the rows to their place in the [B][L][K] grid.
#include <cstring>

#include "internal.h"

namespace shb {

__host__ __device__ inline int64_t shard_offset(int N, int r, int W, int j) { return 2 * ((int64_t)j * (N + 1 - r) - (int64_t)W * j * (j - 1) / 2); }
__host__ __device__ inline int shard_orders(int N, int r, int W) { return r > N ? 0 : (N - r) / W + 1; }

// Cm (m-major, all fields of a row adjacent) -> packed shard.  CTA = (one own order, a block of 64 degrees); thread <-> (degree, cs)
// pair, loop over fields: consecutive threads write consecutive packed elements.
__global__ void __launch_bounds__(128) k_pack_shard(DevTables t, int batch, const double* __restrict__ Cm, double* __restrict__ packed, int64_t ps_b,
                                                    const double* __restrict__ dscale) {
    const int N = t.nmax, W = t.mstride, r = t.m0;
    const int j = blockIdx.x, m = r + j * W;
    const int e = blockIdx.y * 128 + threadIdx.x;           // element inside the order's block: 2 (n - m) + cs
    if (m > N || e >= 2 * (N - m + 1)) return;
    const int k = e >> 1, cs = e & 1;
    const double* src = Cm + (tri(m, m, N) + k) * t.Cs;
    const double sc = dscale ? dscale[m + k] : 1.0;
    double* dst = packed + shard_offset(N, r, W, j) + e;
    for (int b = 0; b < batch; ++b) {
        const double v = src[colidx(cs, b, t.GS)];
        dst[(int64_t)b * ps_b] = dscale ? v * sc : v;
    }
}

// gathered shards -> user nm vector.  thread <-> n-major index k; reads are gathers (L2), writes coalesced.
__global__ void __launch_bounds__(256) k_unpack_shards(int N, int W, int batch, const double* __restrict__ all, int64_t rank_stride, int64_t ps_b,
                                                       double* __restrict__ coef, int64_t cs_b, int64_t cs_nm) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (int64_t)(N + 1) * (N + 1)) return;
    int n = (int)sqrt((double)k);
    while ((int64_t)n * n > k) --n;
    while ((int64_t)(n + 1) * (n + 1) <= k) ++n;
    const int m = (int)(k - (int64_t)n * n - n), am = m < 0 ? -m : m;
    const int r = am % W, j = am / W;
    const double* src = all + (int64_t)r * rank_stride + shard_offset(N, r, W, j) + 2 * (n - am) + (m < 0 ? 1 : 0);
    for (int b = 0; b < batch; ++b) coef[(int64_t)b * cs_b + k * cs_nm] = src[(int64_t)b * ps_b];
}

// slot s of the gathered row buffer -> grid row row_of_slot[s] (-1: padding).  CTA = (slot, field), threads along longitude.
__global__ void __launch_bounds__(256) k_scatter_rows(const double* __restrict__ src, int64_t rank_stride, int64_t src_b, int rmax, int nlon,
                                                      const int32_t* __restrict__ row_of_slot, double* __restrict__ dst, int64_t dst_b, int64_t dst_lat) {
    const int s = blockIdx.x, b = blockIdx.y;
    const int row = row_of_slot[s];
    if (row < 0) return;
    const int r = s / rmax, i = s - r * rmax;
    const double2* a = reinterpret_cast<const double2*>(src + (int64_t)r * rank_stride + (int64_t)b * src_b + (int64_t)i * nlon);
    double* d = dst + (int64_t)b * dst_b + (int64_t)row * dst_lat;
    if ((nlon & 1) == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d) & 15) == 0) {
        double2* d2 = reinterpret_cast<double2*>(d);
        for (int x = threadIdx.x; x < nlon / 2; x += blockDim.x) d2[x] = a[x];
    } else {
        const double* a1 = reinterpret_cast<const double*>(a);
        for (int x = threadIdx.x; x < nlon; x += blockDim.x) d[x] = a1[x];
    }
}

// own row i of field b -> its final place in the grid of every rank (peer pointers: stores travel over NVLink).
// CTA = (own row, field, destination rank); 16-byte stores along longitude.
struct PushDst { double* p[16]; };
__global__ void __launch_bounds__(256) k_push_rows(PushDst dst, const double* __restrict__ src, int64_t src_b, int nlon, const int32_t* __restrict__ rows,
                                                   int64_t dst_b, int64_t dst_lat) {
    const int i = blockIdx.x, b = blockIdx.y, w = blockIdx.z;
    const double* a = src + (int64_t)b * src_b + (int64_t)i * nlon;
    double* d = dst.p[w] + (int64_t)b * dst_b + (int64_t)rows[i] * dst_lat;
    if ((nlon & 1) == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d) & 15) == 0) {
        const double2* a2 = reinterpret_cast<const double2*>(a);
        double2* d2 = reinterpret_cast<double2*>(d);
        for (int x = threadIdx.x; x < nlon / 2; x += blockDim.x) d2[x] = a2[x];
    } else {
        for (int x = threadIdx.x; x < nlon; x += blockDim.x) d[x] = a[x];
    }
}

int do_analysis_packed(Plan& p, int batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, double* packed, int64_t ps_b,
                       cudaStream_t st, const CallOpts& o);

}  // namespace shb

using namespace shb;

extern "C" {

int64_t shb200_shard_size(int32_t nmax, int32_t rank, int32_t world) {
    if (nmax < 0 || world < 1 || rank < 0 || rank >= world) return -1;
    return shard_offset(nmax, rank, world, shard_orders(nmax, rank, world));
}

int shb200_analysis_packed(shb200_plan* plan, int32_t batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon,
                           double* packed, int64_t ps_b, void* stream, const shb200_call_opts* opts) {
    if (!plan || !grid || !packed) { set_error("analysis_packed: null argument"); return SHB200_EINVAL; }
    Plan& p = plan->p;
    CallOpts o;
    int rc = parse_opts(p, opts, o);
    if (rc) return rc;
    if (ps_b < shb200_shard_size(p.nmax, o.shard_rank, o.shard_world)) { set_error("analysis_packed: ps_b is smaller than the shard (shb200_shard_size)"); return SHB200_EINVAL; }
    std::lock_guard<std::mutex> lock(p.mu);
    return do_analysis_packed(p, batch, grid, gs_b, gs_lat, gs_lon, packed, ps_b, (cudaStream_t)stream, o);
}

int shb200_unpack_shards(int32_t nmax, int32_t world, int32_t batch, const double* packed_all, int64_t rank_stride, int64_t ps_b,
                         double* coef, int64_t cs_b, int64_t cs_nm, void* stream) {
    if (nmax < 0 || world < 1 || batch < 1 || !packed_all || !coef) { set_error("unpack_shards: invalid argument"); return SHB200_EINVAL; }
    const int64_t nfull = (int64_t)(nmax + 1) * (nmax + 1);
    k_unpack_shards<<<(unsigned)((nfull + 255) / 256), 256, 0, (cudaStream_t)stream>>>(nmax, world, batch, packed_all, rank_stride, ps_b, coef, cs_b, cs_nm);
    SHB_CUDA(cudaGetLastError());
    return 0;
}

int shb200_scatter_rows(int32_t batch, int32_t world, int32_t rmax, int32_t nlon, const double* src, int64_t rank_stride, int64_t src_b,
                        const int32_t* row_of_slot, double* dst, int64_t dst_b, int64_t dst_lat, void* stream) {
    if (batch < 1 || world < 1 || rmax < 1 || nlon < 1 || !src || !row_of_slot || !dst) { set_error("scatter_rows: invalid argument"); return SHB200_EINVAL; }
    dim3 g((unsigned)(world * rmax), (unsigned)batch);
    k_scatter_rows<<<g, 256, 0, (cudaStream_t)stream>>>(src, rank_stride, src_b, rmax, nlon, row_of_slot, dst, dst_b, dst_lat);
    SHB_CUDA(cudaGetLastError());
    return 0;
}

int shb200_peer_alloc(int64_t bytes, void** ptr, void* handle64) {
    if (bytes <= 0 || !ptr || !handle64) { set_error("peer_alloc: invalid argument"); return SHB200_EINVAL; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* d = nullptr;
    if (cudaMalloc(&d, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); set_error("peer_alloc: out of device memory"); return SHB200_ENOMEM; }
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, d) != cudaSuccess) { cudaGetLastError(); cudaFree(d); set_error("peer_alloc: cudaIpcGetMemHandle failed"); return SHB200_EUNSUPPORTED; }
    memcpy(handle64, &h, 64);
    *ptr = d;
    return 0;
}

int shb200_peer_open(const void* handle64, void** ptr) {
    if (!handle64 || !ptr) { set_error("peer_open: null argument"); return SHB200_EINVAL; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* d = nullptr;
    if (cudaIpcOpenMemHandle(&d, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        set_error("peer_open: cudaIpcOpenMemHandle failed (no peer access between the two devices?)");
        return SHB200_EUNSUPPORTED;
    }
    *ptr = d;
    return 0;
}

int shb200_peer_close(void* ptr) {
    if (ptr && cudaIpcCloseMemHandle(ptr) != cudaSuccess) { cudaGetLastError(); set_error("peer_close: cudaIpcCloseMemHandle failed"); return SHB200_ECUDA; }
    return 0;
}

int shb200_peer_free(void* ptr) {
    if (ptr && cudaFree(ptr) != cudaSuccess) { cudaGetLastError(); set_error("peer_free: cudaFree failed"); return SHB200_ECUDA; }
    return 0;
}

int shb200_push_rows(int32_t batch, int32_t world, int32_t nown, int32_t nlon, const double* src, int64_t src_b, const int32_t* rows_dev,
                     void* const* dst, int64_t dst_b, int64_t dst_lat, void* stream) {
    if (batch < 1 || world < 1 || world > 16 || nown < 0 || nlon < 1 || !src || !rows_dev || !dst) { set_error("push_rows: invalid argument (world <= 16)"); return SHB200_EINVAL; }
    if (nown == 0) return 0;
    PushDst d{};
    for (int w = 0; w < world; ++w) {
        if (!dst[w]) { set_error("push_rows: null destination"); return SHB200_EINVAL; }
        d.p[w] = (double*)dst[w];
    }
    dim3 g((unsigned)nown, (unsigned)batch, (unsigned)world);
    k_push_rows<<<g, 256, 0, (cudaStream_t)stream>>>(d, src, src_b, nlon, rows_dev, dst_b, dst_lat);
    SHB_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"

namespace shb {
void launch_pack_shard(const Plan& p, int batch, double* packed, int64_t ps_b, cudaStream_t st) {
    const int nj = ana_order_count(p);
    if (nj <= 0) return;
    const double* ds = p.dscale_on[1] ? p.dscale[1] : nullptr;
    dim3 g((unsigned)nj, (unsigned)((2 * (p.nmax - p.t.m0 + 1) + 127) / 128));
    trace_kernel(p, "k_pack_shard");
    k_pack_shard<<<g, 128, 0, st>>>(p.t, batch, p.Cm, packed, ps_b, ds);
}
}  // namespace shb
