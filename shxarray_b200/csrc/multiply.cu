// engine.multiply on the device: synthesis of both operands, pointwise product, analysis -- the
// pattern of shtns.py:394-420 / shlib.pyx:140-152 (method="spatial") with the grids never leaving HBM.
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
    SHB_CUDA(cudaSetDevice(pc.device));
    cudaStream_t st = (cudaStream_t)stream;
    int64_t per = (int64_t)pc.nlat * pc.nlon, n = per * batch;
    // grids live in the staging buffers of the two synthesis plans
    auto need = [&](Plan& p) -> int {
        if (p.stage_grid_elems >= n) return 0;
        if (p.stage_grid) cudaFree(p.stage_grid);
        p.stage_grid = nullptr; p.stage_grid_elems = 0;
        SHB_CUDA(cudaMalloc((void**)&p.stage_grid, sizeof(double) * n));
        p.stage_grid_elems = n;
        return 0;
    };
    int rc;
    if ((rc = need(pa))) return rc;
    double* gb = nullptr;
    if (syn_a == syn_b) { SHB_CUDA(cudaMallocAsync((void**)&gb, sizeof(double) * n, st)); }
    else { if ((rc = need(pb))) return rc; gb = pb.stage_grid; }
    if ((rc = shb200_synthesis(syn_a, batch, coef_a, as_b, as_nm, pa.stage_grid, per, pc.nlon, 1, stream))) return rc;
    if ((rc = shb200_synthesis(syn_b, batch, coef_b, bs_b, bs_nm, gb, per, pc.nlon, 1, stream))) return rc;
    k_pointwise_mul<<<148 * 8, 256, 0, st>>>(pa.stage_grid, gb, n);
    rc = shb200_analysis(ana, batch, pa.stage_grid, per, pc.nlon, 1, coef_out, os_b, os_nm, stream);
    if (syn_a == syn_b) SHB_CUDA(cudaFreeAsync(gb, st));
    return rc;
}
