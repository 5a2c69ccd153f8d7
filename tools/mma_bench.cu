// Microbenchmark: sustained tcgen05.mma rate (cycles per instruction) for M=128, K=16, kind::f16, operands resident in
// shared memory (K-major, SWIZZLE_128B), as a function of N and of the A start-address shift. Development aid.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../real_esrgan-pytorch_b200/csrc/ptx.cuh"
using namespace resr;

static constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t desc_of(uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; }

__global__ void __launch_bounds__(128, 1) bench(int N, int shift_rows, int iters, int same_d, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    for (int i = threadIdx.x; i < (17408 + 256 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
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
        const uint32_t idesc = make_idesc_f16(1, 128, N);
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            // warm-up
            for (int i = 0; i < 16; ++i) umma_f16(tbase, desc_of(a_lo), desc_of(b_lo), idesc, 1);
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        tc_fence_after();
        if (elect_one()) {
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t dcol = same_d ? 0 : ((i & 1) * 256);
                        umma_f16(tbase + dcol, desc_of(a_lo + dx * shift_rows * 8 + ks * 2), desc_of(b_lo + dx * ((N * 128) >> 4) + ks * 2), idesc, 1);
                    }
            }
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 1);
        if (elect_one()) {
            t1 = clock64();
            if (blockIdx.x == 0) out[0] = (t1 - t0);
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    for (int grid : {1, 148}) {
        for (int N : {32, 64, 96, 128, 192, 256}) {
            if (3 * N * 128 + 17408 + 2048 > 200 * 1024) continue;
            for (int shift : {0, 1}) {
                bench<<<grid, 128, 200 * 1024>>>(N, shift, iters, 1, d);
                long long c = 0;
                cudaError_t e = cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                const double per = double(c) / (iters * 12.0);
                printf("grid=%3d N=%3d shift=%d : %.1f cycles/MMA  (floor %d)  -> %.0f%% of peak\n", grid, N, shift, per, N / 2, 100.0 * (N / 2) / per);
            }
        }
    }
    return 0;
}
