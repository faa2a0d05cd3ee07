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

}  // namespace shb

using namespace shb;

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
    if (pc.mul_grid_elems < 2 * n) {
        if (pc.mul_grid) { SHB_CUDA(cudaDeviceSynchronize()); cudaFree(pc.mul_grid); }
        pc.mul_grid = nullptr; pc.mul_grid_elems = 0;
        SHB_CUDA(cudaMalloc((void**)&pc.mul_grid, sizeof(double) * 2 * n));
        pc.mul_grid_elems = 2 * n;
    }
    double* ga = pc.mul_grid;
    double* gb = pc.mul_grid + n;
    const CallOpts none;
    int rc;
    if ((rc = do_synthesis(pa, batch, coef_a, as_b, as_nm, ga, per, pc.nlon, 1, st, none))) return rc;
    if ((rc = do_synthesis(pb, batch, coef_b, bs_b, bs_nm, gb, per, pc.nlon, 1, st, none))) return rc;
    k_pointwise_mul<<<148 * 8, 256, 0, st>>>(ga, gb, n);
    return do_analysis(pc, batch, ga, per, pc.nlon, 1, coef_out, os_b, os_nm, st, none);
}
