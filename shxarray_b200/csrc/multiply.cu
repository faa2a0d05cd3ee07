// engine.multiply on the device: synthesis of both operands, pointwise product, analysis -- the
// pattern of shtns.py:394-420 / shlib.pyx:140-152 (method="spatial") with the grids never leaving HBM.
#include <algorithm>

#include "internal.h"

namespace shb {

__global__ void k_pointwise_mul(double* __restrict__ a, const double* __restrict__ b, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) a[i] *= b[i];
}

// grid[b][i] *= mask[i]  (mask: one [nlat, nlon] field with its own strides, shared by the batch)
__global__ void __launch_bounds__(256) k_mask_mul(double* __restrict__ g, const double* __restrict__ mask, int nlat, int nlon, int64_t ms_lat, int64_t ms_lon,
                                                  int64_t per, int batch) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per) return;
    const int il = (int)(i / nlon), ik = (int)(i - (int64_t)il * nlon);
    const double w = mask[il * ms_lat + ik * ms_lon];
    for (int b = 0; b < batch; ++b) g[(int64_t)b * per + i] *= w;
}

static int grow_mul_grid(Plan& pc, int64_t need) {
    if (pc.mul_grid_elems < need) {
        if (pc.mul_grid) { SHB_CUDA(cudaDeviceSynchronize()); cudaFree(pc.mul_grid); drop_graphs(pc); }
        pc.mul_grid = nullptr; pc.mul_grid_elems = 0;
        SHB_CUDA(cudaMalloc((void**)&pc.mul_grid, sizeof(double) * need));
        pc.mul_grid_elems = need;
    }
    return 0;
}

}  // namespace shb

using namespace shb;

// Grid-space operator  out = analysis( mask x synthesis(coef) ):  the ocean function of the sea-level equation applied to a
// load (spectralsealevel.py:97-104 does it with an O(nsh^2) product-to-sum matrix limited to nmax <= 150), basin masks, any
// pointwise weight.  The grid never leaves HBM; per-degree scales (Love-number kernels) ride on either transform.
extern "C" int shb200_apply_mask(shb200_plan* syn, shb200_plan* ana, int32_t batch, const double* coef, int64_t cs_b, int64_t cs_nm,
                                 const double* mask, int64_t ms_lat, int64_t ms_lon, double* coef_out, int64_t os_b, int64_t os_nm,
                                 void* stream, const shb200_call_opts* syn_opts, const shb200_call_opts* ana_opts) {
    if (!syn || !ana || !coef || !mask || !coef_out) { set_error("apply_mask: null argument"); return SHB200_EINVAL; }
    Plan& pa = syn->p; Plan& pc = ana->p;
    if (pa.mode != SHB200_MODE_GRID || pa.nlat != pc.nlat || pa.nlon != pc.nlon || pa.device != pc.device) {
        set_error("apply_mask: the two plans must share one lon/lat grid and device");
        return SHB200_EINVAL;
    }
    if (batch < 1 || batch > pa.maxB || batch > pc.maxB) { set_error("apply_mask: batch exceeds a plan's max_batch"); return SHB200_EINVAL; }
    CallOpts oa, oc;
    int rc;
    if ((rc = parse_opts(pa, syn_opts, oa)) || (rc = parse_opts(pc, ana_opts, oc))) return rc;
    Plan* uniq[2] = {&pa, &pc};
    std::sort(uniq, uniq + 2);
    Plan** end = std::unique(uniq, uniq + 2);
    for (Plan** q = uniq; q != end; ++q) (*q)->mu.lock();
    struct Unlock { Plan** b; Plan** e; ~Unlock() { for (Plan** q = b; q != e; ++q) (*q)->mu.unlock(); } } unlock{uniq, end};
    SHB_CUDA(cudaSetDevice(pc.device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t per = (int64_t)pc.nlat * pc.nlon, n = per * batch;
    if ((rc = grow_mul_grid(pc, n))) return rc;
    double* g = pc.mul_grid;
    auto body = [&](cudaStream_t s, bool capturing) {
        CallOpts a = oa, c = oc;
        a.no_events = c.no_events = capturing;
        int r;
        if ((r = do_synthesis(pa, batch, coef, cs_b, cs_nm, g, per, pc.nlon, 1, s, a))) return r;
        k_mask_mul<<<(unsigned)((per + 255) / 256), 256, 0, s>>>(g, mask, pc.nlat, pc.nlon, ms_lat, ms_lon, per, batch);
        return do_analysis(pc, batch, g, per, pc.nlon, 1, coef_out, os_b, os_nm, s, c);
    };
    if (graph_eligible(pa, oa) && graph_eligible(pc, oc)) {      // the sea-level iteration calls this with the same buffers every time
        GraphKey k;
        k.kind = 3; k.batch = batch; k.serial[0] = pa.serial; k.serial[1] = pc.serial; k.shard_rank = oc.shard_rank; k.shard_world = oc.shard_world;
        k.ptr[0] = coef; k.ptr[1] = mask; k.ptr[2] = coef_out; k.ptr[3] = g;
        k.stride[0] = cs_b; k.stride[1] = cs_nm; k.stride[2] = ms_lat; k.stride[3] = ms_lon; k.stride[4] = os_b; k.stride[5] = os_nm;
        return run_graphed(pc, k, st, uniq, (int)(end - uniq), body);
    }
    return body(st, false);
}

extern "C" int shb200_multiply(shb200_plan* syn_a, shb200_plan* syn_b, shb200_plan* ana, int32_t batch,
                               const double* coef_a, int64_t as_b, int64_t as_nm, const double* coef_b, int64_t bs_b, int64_t bs_nm,
                               double* coef_out, int64_t os_b, int64_t os_nm, void* stream) {
    if (!syn_a || !syn_b || !ana || !coef_a || !coef_b || !coef_out) { set_error("multiply: null argument"); return SHB200_EINVAL; }
    Plan& pa = syn_a->p; Plan& pb = syn_b->p; Plan& pc = ana->p;
    if (pa.mode != SHB200_MODE_GRID || pa.nlat != pc.nlat || pa.nlon != pc.nlon || pb.nlat != pc.nlat || pb.nlon != pc.nlon ||
        pa.device != pc.device || pb.device != pc.device) {
        set_error("multiply: the three plans must share one lon/lat grid and device");
        return SHB200_EINVAL;
    }
    if (batch < 1 || batch > pa.maxB || batch > pb.maxB || batch > pc.maxB) { set_error("multiply: batch exceeds a plan's max_batch"); return SHB200_EINVAL; }
    // the (up to three distinct) plans are locked in address order for the whole launch sequence
    Plan* uniq[3] = {&pa, &pb, &pc};
    std::sort(uniq, uniq + 3);
    Plan** end = std::unique(uniq, uniq + 3);
    for (Plan** q = uniq; q != end; ++q) (*q)->mu.lock();
    struct Unlock { Plan** b; Plan** e; ~Unlock() { for (Plan** q = b; q != e; ++q) (*q)->mu.unlock(); } } unlock{uniq, end};
    SHB_CUDA(cudaSetDevice(pc.device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t per = (int64_t)pc.nlat * pc.nlon, n = per * batch;
    // the two grids live in ONE buffer owned by the analysis plan (grown on demand; growing synchronises the device, so
    // no kernel of an earlier call can still be using the old allocation)
    { int rcg = grow_mul_grid(pc, 2 * n); if (rcg) return rcg; }
    double* ga = pc.mul_grid;
    double* gb = pc.mul_grid + n;
    auto body = [&](cudaStream_t s, bool capturing) {
        CallOpts none;
        none.no_events = capturing;
        int r;
        if ((r = do_synthesis(pa, batch, coef_a, as_b, as_nm, ga, per, pc.nlon, 1, s, none))) return r;
        if ((r = do_synthesis(pb, batch, coef_b, bs_b, bs_nm, gb, per, pc.nlon, 1, s, none))) return r;
        k_pointwise_mul<<<148 * 8, 256, 0, s>>>(ga, gb, n);
        return do_analysis(pc, batch, ga, per, pc.nlon, 1, coef_out, os_b, os_nm, s, none);
    };
    const CallOpts plain;
    if (graph_eligible(pa, plain) && graph_eligible(pb, plain) && graph_eligible(pc, plain)) {      // ten launches -> one graph launch
        GraphKey k;
        k.kind = 4; k.batch = batch; k.serial[0] = pa.serial; k.serial[1] = pb.serial; k.serial[2] = pc.serial;
        k.ptr[0] = coef_a; k.ptr[1] = coef_b; k.ptr[2] = coef_out; k.ptr[3] = ga;
        k.stride[0] = as_b; k.stride[1] = as_nm; k.stride[2] = bs_b; k.stride[3] = bs_nm; k.stride[4] = os_b; k.stride[5] = os_nm;
        return run_graphed(pc, k, st, uniq, (int)(end - uniq), body);
    }
    return body(st, false);
}
