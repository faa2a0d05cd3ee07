// How fast can 8 warps per SM issue DMMA.8x8x4 when every operand fragment comes from shared memory?
// One CTA of 256 threads per SM (and, for the analysis shape, of 128 and 64 threads: one and half a warp per SM sub-partition); per k-step NA A-fragment loads + NB B-fragment loads (LDS.64, conflict-free padded rows)
// feed NA*NB DMMAs on NA*NB independent accumulators -- the inner-loop shapes of k_leg_ana_tab (NA=2, NB=4) and
// k_leg_syn_tab (NA=4, NB=4).  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/dmma_feed.cu -o /tmp/dmma_feed
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NA, int NB, int KS, int NT>
__global__ void __launch_bounds__(NT, 1) k_feed(double* out, int iters) {
    extern __shared__ double sm[];
    // A tile: [16 rows][34], B tile: [KS*4 rows][130]  (padded like the real kernels)
    double* sA = sm;
    double* sB = sm + 4 * 16 * 34;
    for (int i = threadIdx.x; i < 4 * 16 * 34 + KS * 4 * 2 * 130; i += NT) sm[i] = 1e-3 * (i % 7);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lr = lane >> 2, lk = lane & 3;
    double acc[NA][NB][2];
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
    const double* PA = sA + (2 * lr + (warp & 1)) * 34 + lk;
    const double* PB = sB + (warp & 1) * 130 + 32 * (warp >> 1) + lr;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            double A[NA], B[NB];
#pragma unroll
            for (int a = 0; a < NA; ++a) A[a] = PA[a * 16 * 34 + 4 * ks];
#pragma unroll
            for (int b = 0; b < NB; ++b) B[b] = PB[(4 * ks + lk) * 2 * 130 + 8 * b];
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int b = 0; b < NB; ++b) dmma(acc[a][b][0], acc[a][b][1], A[a], B[b]);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) s += acc[a][b][0] + acc[a][b][1];
    if (s == 12345.678) out[0] = s;
}

template <int NA, int NB, int KS, int NT = 256>
int run(const char* tag, int sms, double* out) {
    const int iters = 2048;
    const size_t smem = sizeof(double) * (4 * 16 * 34 + KS * 4 * 2 * 130);
    CK(cudaFuncSetAttribute(k_feed<NA, NB, KS, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_feed<NA, NB, KS, NT><<<sms, NT, smem>>>(out, iters);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k_feed<NA, NB, KS, NT><<<sms, NT, smem>>>(out, iters);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double fl = 2.0 * 256 * NA * NB * KS * (double)iters * (NT / 32) * sms;
    printf("\"%s\": %.3f, ", tag, fl / ms * 1e-9);
    return 0;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 1024));
    printf("{");
    if (run<2, 4, 8>("ana_shape_2A_4B_tflops", sms, out)) return 1;
    if (run<4, 4, 8>("syn_shape_4A_4B_tflops", sms, out)) return 1;
    if (run<1, 4, 8>("shape_1A_4B_tflops", sms, out)) return 1;
    if (run<2, 2, 8>("shape_2A_2B_tflops", sms, out)) return 1;
    if (run<4, 2, 8>("shape_4A_2B_tflops", sms, out)) return 1;
    if (run<2, 4, 8, 128>("ana_shape_4warps_tflops", sms, out)) return 1;
    if (run<2, 4, 8, 64>("ana_shape_2warps_tflops", sms, out)) return 1;
    printf("\"note\": \"8 warps per SM, all fragments from padded shared-memory tiles, 8 k-steps per iteration\"}\n");
    return 0;
}
