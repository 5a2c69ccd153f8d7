// Microbenchmark: sustained pure tcgen05.mma stream on every SM (random fp16 operands resident in shared memory) for a
// given N, long enough for the board power limit to act. Prints achieved TFLOP/s; run under tools/power_run.py to get the
// NVML power / clock beside it. Development aid: what would an N = 96 / 192 / 256 MMA stream deliver under the 1000 W cap?
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../real_esrgan-pytorch_b200/csrc/ptx.cuh"
using namespace resr;
static constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t desc_of(uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; }

__global__ void __launch_bounds__(128, 1) bench(int N, int iters, unsigned seed) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    unsigned st = seed + blockIdx.x * 7919u + threadIdx.x * 104729u;
    for (int i = threadIdx.x; i < (17408 + 3 * 256 * 128) / 4; i += blockDim.x) {
        st = st * 1664525u + 1013904223u;
        // two fp16 values in [-2, 2) with random mantissas
        const uint32_t lo = 0x3800u | ((st >> 3) & 0x87FFu), hi = 0x3800u | ((st >> 17) & 0x87FFu);
        reinterpret_cast<uint32_t*>(smem)[i] = lo | (hi << 16);
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_ptr;
    if (warp == 1) {
        const uint32_t a_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
        const uint32_t b_lo = (smem_u32(smem + 17408) & 0x3FFFFu) >> 4;
        const uint32_t idesc = make_idesc_f16(0, 128, N);
        const uint32_t WT = (N * 128) >> 4;
        const int nslots = 512 / N;
        for (int it = 0; it < iters; ++it) {
            const uint32_t d = tbase + (it % nslots) * N;
            if (elect_one()) {
#pragma unroll
                for (int i = 0; i < 12; ++i) {
                    const int dx = i >> 2, ks = i & 3;
                    umma_f16(d, desc_of(a_lo + dx * 8 + ks * 2), desc_of(b_lo + dx * WT + ks * 2), idesc, (it >= nslots) ? 1 : 0);
                }
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 96;
    const double seconds = argc > 2 ? atof(argv[2]) : 3.0;
    const int smem = 17408 + 3 * 256 * 128 + 2048;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 20000;  // 12 MMAs each
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<<<148, 128, smem>>>(N, 100, 1);
    cudaDeviceSynchronize();
    double total_ms = 0; long long launches = 0; float last = 0;
    while (total_ms < seconds * 1e3) {
        cudaEventRecord(e0);
        for (int k = 0; k < 4; ++k) bench<<<148, 128, smem>>>(N, iters, 1 + launches);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        cudaEventElapsedTime(&last, e0, e1);
        total_ms += last; launches += 4;
    }
    const double flop = 4.0 * 148 * iters * 12.0 * 2.0 * 128 * N * 16;
    printf("N=%d: last group %.2f ms -> %.0f TFLOP/s (fp16 operands, random data, pure MMA stream, all 148 SMs)\n", N, last, flop / (last * 1e-3) / 1e12);
    return 0;
}
