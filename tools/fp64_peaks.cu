// FP64 peak microbenchmarks for B200 (sm_100a): DFMA issue rate, DMMA m8n8k4 rate, mixed, DDIV, and
// a shared-memory-fed DFMA variant.  Output: one JSON object on stdout (-> profiles/fp64_peaks_r01.json).
// MEASURED_PEAKS.json only carries HBM GB/s and bf16 TF/s; this supplies the FP64 roofline denominator
// (SURVEY.md §8d asks for it).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a, double b) {
    double d0[ILP], d1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { d0[i] = threadIdx.x * 1e-3 + i; d1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma884(d0[i], d1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d0[i] + d1[i];
    if (s == 12345.678) out[0] = s;
}

// mixed: per iteration ILP DMMAs and NF DFMAs (independent chains)
template <int ILP, int NF>
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double a, double b) {
    double d0[ILP], d1[ILP], f[NF];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { d0[i] = threadIdx.x * 1e-3 + i; d1[i] = i; }
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma884(d0[i], d1[i], a, b);
#pragma unroll
        for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d0[i] + d1[i];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
    if (s == 12345.678) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) k_ddiv(double* out, int iters, double a) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + 1.0 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = __ddiv_rn(acc[i], a);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

// DFMA with the multiplicand broadcast from shared memory (LDS.128 -> 2 FMAs x ROWS rows)
template <int ROWS, int COLS>
__global__ void __launch_bounds__(256) k_dfma_lds(double* out, int iters, double a) {
    __shared__ double2 sm[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sm[i] = make_double2(1.0 + i * 1e-9, 1.0 - i * 1e-9);
    __syncthreads();
    double acc[ROWS][COLS];
    double p[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) { p[r] = a + r * 1e-9 + threadIdx.x * 1e-12;
#pragma unroll
        for (int c = 0; c < COLS; ++c) acc[r][c] = r + c; }
    for (int it = 0; it < iters; ++it) {
        const double2* row = sm + ((it * (COLS / 2)) & 255);
#pragma unroll
        for (int c = 0; c < COLS; c += 2) {
            double2 v = row[c / 2];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) { acc[r][c] = fma(p[r], v.x, acc[r][c]); acc[r][c + 1] = fma(p[r], v.y, acc[r][c + 1]); }
        }
    }
    double s = 0;
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int c = 0; c < COLS; ++c) s += acc[r][c];
    if (s == 12345.678) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 1024));
    const int iters = 4096;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", prop.name, sms, prop.clockRate);
    for (int occ = 1; occ <= 8; occ *= 2) {
        int grid = sms * occ;
        double ms = time_ms([&] { k_dfma<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double fl = 2.0 * 8 * iters * 256.0 * grid;
        printf(", \"dfma_tflops_occ%d\": %.3f", occ, fl / ms * 1e-9);
    }
    for (int occ = 1; occ <= 8; occ *= 2) {
        int grid = sms * occ;
        double ms = time_ms([&] { k_dmma<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double fl = 2.0 * 256 * 8 * iters * 8.0 * grid;   // 8 warps per CTA, 8x8x4 FMAs per mma
        printf(", \"dmma884_tflops_occ%d\": %.3f", occ, fl / ms * 1e-9);
    }
    {
        int grid = sms * 4;
        double ms = time_ms([&] { k_dmma<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double fl = 2.0 * 256 * 16 * iters * 8.0 * grid;
        printf(", \"dmma884_ilp16_tflops\": %.3f", fl / ms * 1e-9);
    }
    {
        int grid = sms * 4;
        double ms = time_ms([&] { k_mixed<8, 8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double fl_mma = 2.0 * 256 * 8 * iters * 8.0 * grid, fl_fma = 2.0 * 8 * iters * 256.0 * grid;
        printf(", \"mixed_8mma_8fma_ms\": %.4f, \"mixed_8mma_8fma_tflops_mma\": %.3f, \"mixed_8mma_8fma_tflops_fma\": %.3f", ms, fl_mma / ms * 1e-9, fl_fma / ms * 1e-9);
        ms = time_ms([&] { k_mixed<8, 32><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        fl_fma = 2.0 * 32 * iters * 256.0 * grid;
        printf(", \"mixed_8mma_32fma_ms\": %.4f, \"mixed_8mma_32fma_tflops_mma\": %.3f, \"mixed_8mma_32fma_tflops_fma\": %.3f", ms, fl_mma / ms * 1e-9, fl_fma / ms * 1e-9);
    }
    {
        int grid = sms * 4;
        double ms = time_ms([&] { k_ddiv<8><<<grid, 256>>>(out, 1024, 1.0000001); }, 5);
        double ops = 8.0 * 1024 * 256.0 * grid;
        printf(", \"ddiv_gops\": %.3f", ops / ms * 1e-6);
    }
    {
        int grid = sms * 2;
        double ms = time_ms([&] { k_dfma_lds<1, 32><<<grid, 256>>>(out, iters, 1.0000001); }, 5);
        printf(", \"dfma_lds_r1c32_tflops\": %.3f", 2.0 * 32 * iters * 256.0 * grid / ms * 1e-9);
        ms = time_ms([&] { k_dfma_lds<2, 16><<<grid, 256>>>(out, iters, 1.0000001); }, 5);
        printf(", \"dfma_lds_r2c16_tflops\": %.3f", 2.0 * 32 * iters * 256.0 * grid / ms * 1e-9);
        ms = time_ms([&] { k_dfma_lds<4, 8><<<grid, 256>>>(out, iters, 1.0000001); }, 5);
        printf(", \"dfma_lds_r4c8_tflops\": %.3f", 2.0 * 32 * iters * 256.0 * grid / ms * 1e-9);
    }
    // sustained DFMA for ~2 s to see the clock the FP64 pipe settles at
    {
        int grid = sms * 4;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        for (int r = 0; r < 400; ++r) k_dfma<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf(", \"dfma_sustained_tflops\": %.3f, \"dfma_sustained_s\": %.2f", 400.0 * 2.0 * 8 * iters * 256.0 * grid / ms * 1e-9, ms * 1e-3);
        CK(cudaEventRecord(e0));
        for (int r = 0; r < 400; ++r) k_dmma<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf(", \"dmma_sustained_tflops\": %.3f, \"dmma_sustained_s\": %.2f", 400.0 * 2.0 * 256 * 8 * iters * 8.0 * grid / ms * 1e-9, ms * 1e-3);
    }
    printf("}\n");
    return 0;
}
