// Microbenchmark 2: what slows the N=96 MMA stream inside the conv kernel? Variants add, one at a time, the other
// activities of the real kernel. Development aid.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../real_esrgan-pytorch_b200/csrc/ptx.cuh"
using namespace resr;

static constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t desc_of(uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; }

// flags: 1 = commit every 12 MMAs; 2 = rotate D slot per 12 MMAs; 4 = concurrent TMEM ld/st warps; 8 = concurrent bulk loads
__global__ void __launch_bounds__(256, 1) bench(int N, int iters, int flags, const uint8_t* gsrc, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar, cbar[4], lbar[4];
    __shared__ uint32_t tmem_ptr;
    __shared__ volatile int stop;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    uint8_t* wsm = smem + 17408;
    uint8_t* junk = wsm + 3 * 96 * 128 * 2;  // 4 x 17408 landing buffers for the bulk loads
    for (int i = threadIdx.x; i < (17408 + 3 * 96 * 128 * 2) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        for (int i = 0; i < 4; ++i) { mbar_init(&cbar[i], 1); mbar_init(&lbar[i], 1); }
        fence_mbar_init();
        stop = 0;
    }
    if (warp == 0) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_ptr;
    if (warp == 1) {
        const uint32_t a_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
        const uint32_t b_lo = (smem_u32(wsm) & 0x3FFFFu) >> 4;
        const uint32_t idesc = make_idesc_f16(1, 128, N);
        long long t0 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const uint32_t dcol = (flags & 2) ? ((i % 5) * 96) : 0;
#pragma unroll
                for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_f16(tbase + dcol, desc_of(a_lo + dx * 8 + ks * 2), desc_of(b_lo + dx * ((N * 128) >> 4) + ks * 2), idesc, 1);
                if (flags & 1) umma_commit(&cbar[i & 3]);
            }
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        if (elect_one()) {
            if (blockIdx.x == 0) out[0] = clock64() - t0;
            stop = 1;
        }
        __syncwarp();
    } else if (warp >= 4 && (flags & 4)) {
        const uint32_t lane_base = tbase + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 480;
        float v[32];
        while (!stop) {
            tmem_ld32(lane_base, v);
            tmem_ld_wait();
            tmem_st_zero32(lane_base);
            tmem_st_wait();
            __nanosleep(200);
        }
    } else if (warp == 2 && (flags & 8)) {
        uint32_t ph = 0;
        int k = 0;
        while (!stop) {
            if (elect_one()) {
                mbar_expect_tx(&lbar[k], 16640);
                bulk_load_1d(junk + k * 17408, gsrc + (size_t)((blockIdx.x * 64 + (k + ph * 4)) % 4096) * 17408, 16640, &lbar[k]);
            }
            __syncwarp();
            if (k == 3) {  // wait for the oldest batch before reusing
                for (int j = 0; j < 4; ++j) mbar_wait(&lbar[j], ph & 1);
                ph++;
            }
            k = (k + 1) & 3;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
    long long* d;
    uint8_t* g;
    cudaMalloc(&d, 8);
    cudaMalloc(&g, (size_t)4096 * 17408);
    cudaMemset(g, 0, (size_t)4096 * 17408);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 3000, N = 96;
    const char* names[] = {"baseline", "+commit/12", "+rotateD", "+commit+rotateD", "+tmem ld/st warps", "+bulk loads", "all"};
    const int flagv[] = {0, 1, 2, 3, 4, 8, 15};
    for (int grid : {1, 148})
        for (int v = 0; v < 7; ++v) {
            bench<<<grid, 256, 200 * 1024>>>(N, iters, flagv[v], g, d);
            long long c = 0;
            cudaError_t e = cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            printf("grid=%3d %-20s: %.1f cycles/MMA\n", grid, names[v], double(c) / (iters * 12.0));
        }
    return 0;
}
