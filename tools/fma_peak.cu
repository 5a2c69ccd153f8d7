// fp32 FMA peak of this GPU: register-resident FFMA chains (16 independent accumulators per thread, 2048 threads per SM),
// burst (best single launch) and sustained (back to back for a few seconds, i.e. at the power-capped clock).
// The denominator of the "FMA-bound" claim for the degradation stencils (SURVEY.md §8d caveat).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fma_peak tools/fma_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) fma_kernel(float* out, int iters, float a, float b) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = threadIdx.x * 1e-3f + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = fmaf(acc[j], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += acc[j];
    if (s == 12345.678f) out[0] = s;  // never true: keeps the chain alive
}

// the same chains as packed fp32x2 FMAs (Blackwell FFMA2): 8 float2 accumulators per thread
__global__ void __launch_bounds__(256) fma2_kernel(float* out, int iters, float a, float b) {
    float2 acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = make_float2(threadIdx.x * 1e-3f + j, threadIdx.x * 2e-3f + j);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = __ffma2_rn(acc[j], a2, b2);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[j].x + acc[j].y;
    if (s == 12345.678f) out[0] = s;
}

int main(int argc, char** argv) {
    const double seconds = argc > 1 ? atof(argv[1]) : 3.0;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float* out;
    cudaMalloc(&out, 4);
    const int iters = 4096, blocks = sms * 8;
    const double fma_per_launch = static_cast<double>(blocks) * 256 * iters * 8 * 16;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) fma_kernel<<<blocks, 256>>>(out, iters, 0.999f, 1e-3f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int i = 0; i < 10; ++i) {
        cudaEventRecord(e0);
        fma_kernel<<<blocks, 256>>>(out, iters, 0.999f, 1e-3f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double total = 0; float last = 0;
    while (total < seconds * 1e3) {
        cudaEventRecord(e0);
        for (int k = 0; k < 16; ++k) fma_kernel<<<blocks, 256>>>(out, iters, 0.999f, 1e-3f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&last, e0, e1);
        total += last;
    }
    const double burst = fma_per_launch / (best * 1e-3) / 1e12, sustained = 16 * fma_per_launch / (last * 1e-3) / 1e12;
    {   // packed variant: same FMA count per launch (8 float2 accumulators x 8 x iters)
        float best2 = 1e30f;
        for (int i = 0; i < 3; ++i) fma2_kernel<<<blocks, 256>>>(out, iters, 0.999f, 1e-3f);
        cudaDeviceSynchronize();
        for (int i = 0; i < 10; ++i) {
            cudaEventRecord(e0);
            fma2_kernel<<<blocks, 256>>>(out, iters, 0.999f, 1e-3f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best2) best2 = ms;
        }
        fprintf(stderr, "FFMA2 stream: %.2f TFMA/s (burst; same FMA count in half the instructions)\n", fma_per_launch / (best2 * 1e-3) / 1e12);
    }
    printf("{\"fp32_tfma_burst\": %.2f, \"fp32_tfma_sustained\": %.2f, \"fp32_tflops_burst\": %.2f, \"fp32_tflops_sustained\": %.2f, "
           "\"sms\": %d, \"how\": \"FFMA chains, 16 accumulators/thread, 2048 threads/SM; burst = best of 10 single launches, "
           "sustained = back to back for %.0f s\"}\n", burst, sustained, 2 * burst, 2 * sustained, sms, seconds);
    return 0;
}
