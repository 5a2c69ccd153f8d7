// Microbenchmark 3: cost of the per-stage control sequence of the MMA-issuing warp around 12 N=96 MMAs (672 cycles of
// tensor work). Development aid.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../real_esrgan-pytorch_b200/csrc/ptx.cuh"
using namespace resr;
static constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t desc_of(uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; }

template <int I0, int I1>
__device__ __forceinline__ void issue(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
#pragma unroll
    for (int i = I0; i < I1; ++i) {
        const int dx = i >> 2, ks = i & 3;
        umma_f16(d, desc_of(a_lo + dx * 8 + ks * 2), desc_of(b_lo + dx * 768 + ks * 2), idesc, 1);
    }
}

__global__ void __launch_bounds__(128, 1) bench(int variant, int iters, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar, done_bar[8], ready[8];
    __shared__ uint32_t tmem_ptr;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    for (int i = threadIdx.x; i < (17408 + 3 * 96 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        for (int i = 0; i < 8; ++i) { mbar_init(&done_bar[i], 1); mbar_init(&ready[i], 1); }
        fence_mbar_init();
        for (int i = 0; i < 8; ++i) mbar_arrive(&ready[i]);  // phase 0 complete: parity 0 probes succeed forever
    }
    if (warp == 0) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_ptr;
    if (warp == 1) {
        const uint32_t a_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
        const uint32_t b_lo = (smem_u32(smem + 17408) & 0x3FFFFu) >> 4;
        const uint32_t idesc = make_idesc_f16(1, 128, 96);
        const long long t0 = clock64();
        bool rdy = true;
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tbase + (i % 5) * 96;
            uint64_t* rb = &ready[i & 7];
            if (variant == 0) {  // pure stream
                if (elect_one()) issue<0, 12>(d, a_lo, b_lo, idesc);
                __syncwarp();
            } else if (variant == 1) {  // current kernel: split issue, 2 test_wait probes, 2 commits
                if (!rdy) mbar_wait(rb, 0);
                tc_fence_after();
                if (elect_one()) issue<0, 6>(d, a_lo, b_lo, idesc);
                __syncwarp();
                rdy = mbar_test_wait(&ready[(i + 1) & 7], 0);
                bool s = mbar_test_wait(&ready[(i + 2) & 7], 0);
                if (elect_one()) { issue<6, 12>(d, a_lo, b_lo, idesc); umma_commit(&done_bar[i & 7]); umma_commit(&done_bar[(i + 4) & 7]); }
                __syncwarp();
                rdy = rdy && s;
            } else if (variant == 2) {  // classic: try_wait, fence, 12 MMAs, 1 commit
                mbar_wait(rb, 0);
                tc_fence_after();
                if (elect_one()) { issue<0, 12>(d, a_lo, b_lo, idesc); umma_commit(&done_bar[i & 7]); }
                __syncwarp();
            } else if (variant == 3) {  // classic + 2 commits
                mbar_wait(rb, 0);
                tc_fence_after();
                if (elect_one()) { issue<0, 12>(d, a_lo, b_lo, idesc); umma_commit(&done_bar[i & 7]); umma_commit(&done_bar[(i + 4) & 7]); }
                __syncwarp();
            } else if (variant == 4) {  // two waits + fence + 12 MMAs + 2 commits (v2 before probes)
                mbar_wait(rb, 0);
                mbar_wait(&ready[(i + 3) & 7], 0);
                tc_fence_after();
                if (elect_one()) { issue<0, 12>(d, a_lo, b_lo, idesc); umma_commit(&done_bar[i & 7]); umma_commit(&done_bar[(i + 4) & 7]); }
                __syncwarp();
            } else if (variant == 5) {  // no waits at all, 1 commit
                if (elect_one()) { issue<0, 12>(d, a_lo, b_lo, idesc); umma_commit(&done_bar[i & 7]); }
                __syncwarp();
            } else if (variant == 7) {  // 6 MMA, commit, 6 MMA, commit (spaced commits)
                if (elect_one()) { issue<0, 6>(d, a_lo, b_lo, idesc); umma_commit(&done_bar[i & 7]); issue<6, 12>(d, a_lo, b_lo, idesc); umma_commit(&done_bar[(i + 4) & 7]); }
                __syncwarp();
            } else if (variant == 8) {  // 12 MMA, then 2 commits issued by ANOTHER elected pass after a syncwarp
                if (elect_one()) { issue<0, 12>(d, a_lo, b_lo, idesc); }
                __syncwarp();
                if (elect_one()) { umma_commit(&done_bar[i & 7]); }
                __syncwarp();
            } else if (variant == 9) {  // 24 MMAs then 1 commit
                if (elect_one()) { issue<0, 12>(d, a_lo, b_lo, idesc); issue<0, 12>(d, a_lo, b_lo, idesc); umma_commit(&done_bar[i & 7]); }
                __syncwarp();
            } else if (variant == 6) {  // one wait per 2 stages (24 MMAs), 1 commit per 12
                if ((i & 1) == 0) { mbar_wait(rb, 0); tc_fence_after(); }
                if (elect_one()) { issue<0, 12>(d, a_lo, b_lo, idesc); umma_commit(&done_bar[i & 7]); }
                __syncwarp();
            }
        }
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        if (elect_one() && blockIdx.x == 0) out[0] = clock64() - t0;
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int iters = 4000;
    const char* names[] = {"pure MMA stream", "split issue + 2 probes + 2 commits (current)", "try_wait + fence + 12 MMA + commit",
                           "try_wait + fence + 12 MMA + 2 commits", "2 try_wait + fence + 12 MMA + 2 commits", "12 MMA + commit, no waits",
                           "1 wait per 24 MMA, commit per 12", "6 MMA, commit, 6 MMA, commit", "12 MMA | commit (separate elect)", "24 MMA + commit (per 2 stages)"};
    for (int v = 0; v < 10; ++v) {
        bench<<<148, 128, 100 * 1024>>>(v, iters, d);
        long long c = 0;
        cudaError_t e = cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        printf("%-48s: %.0f cycles/stage (tensor work 672)\n", names[v], double(c) / iters);
    }
    return 0;
}
