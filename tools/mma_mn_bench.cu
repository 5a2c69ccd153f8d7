// Microbenchmark: sustained tcgen05.mma rate (cycles per instruction) for M = 128, K = 16, kind::f16 (bf16), operands resident
// in shared memory with SWIZZLE_128B, as a function of N and of the operand MAJOR-ness (K-major vs MN-major, both A and B):
// the weight-gradient kernel (csrc/wgrad_mn.cu) reads both operands MN-major. Development aid.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../real_esrgan-pytorch_b200/csrc/ptx.cuh"
using namespace resr;

static constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t desc_k(uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | (lo & 0x3FFFu); }
__device__ __forceinline__ uint64_t desc_mn(uint32_t lo, uint32_t lbo16) { return (static_cast<uint64_t>(kDescHi) << 32) | (lbo16 << 16) | (lo & 0x3FFFu); }

// amaj / bmaj: 0 K-major, 1 MN-major. nacc: accumulators cycled through (1 or 3). shift: dx shifts A by one row (128 B).
__global__ void __launch_bounds__(128, 1) bench(int N, int amaj, int bmaj, int nacc, int shift, int iters, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_ptr;
    if (warp == 1) {
        const uint32_t a_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
        const uint32_t b_lo = (smem_u32(smem + 64 * 1024) & 0x3FFFFu) >> 4;
        const uint32_t idesc = make_idesc_f16(1, 128, N) | (amaj ? (1u << 15) : 0u) | (bmaj ? (1u << 16) : 0u);
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            for (int i = 0; i < 16; ++i)
                umma_f16(tbase, amaj ? desc_mn(a_lo, 9216 >> 4) : desc_k(a_lo), bmaj ? desc_mn(b_lo, 8192 >> 4) : desc_k(b_lo), idesc, 1);
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
                        const uint32_t dcol = (nacc == 3 ? dx : 0) * N;
                        const uint64_t ad = amaj ? desc_mn(a_lo + ks * (2048 >> 4) + dx * shift * 8, 9216 >> 4) : desc_k(a_lo + dx * shift * 8 + ks * 2);
                        const uint64_t bd = bmaj ? desc_mn(b_lo + ks * (2048 >> 4), 8192 >> 4) : desc_k(b_lo + ks * 2);
                        umma_f16(tbase + dcol, ad, bd, idesc, 1);
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
    for (int grid : {1, 148})
        for (int N : {64, 128})
            for (int maj = 0; maj < 4; ++maj)
                for (int nacc : {1, 3})
                    for (int shift : {0, 1}) {
                        if (grid == 1 && (nacc == 1 || shift == 0)) continue;
                        bench<<<grid, 128, 200 * 1024>>>(N, maj & 1, maj >> 1, nacc, shift, iters, d);
                        long long c = 0;
                        cudaError_t e = cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                        const double per = double(c) / (iters * 12.0);
                        printf("grid=%3d N=%3d A=%s B=%s acc=%d shift=%d : %.1f cycles/MMA (floor %d) -> %.0f%% of peak\n", grid, N,
                               (maj & 1) ? "MN" : "K ", (maj >> 1) ? "MN" : "K ", nacc, shift, per, N / 4, 100.0 * (N / 4) / per);
                    }
    return 0;
}
