// Microbenchmark: per-SM TMA load rate for the conv kernel's activation boxes ([64 ch x R px x 1 x 1], 128B swizzle).
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../real_esrgan-pytorch_b200/csrc/ptx.cuh"
using namespace resr;

__global__ void __launch_bounds__(128, 1) bench(const __grid_constant__ CUtensorMap tm, int nstages, int iters, int box_bytes, int hot,
                                                int rows_per_cta, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t full[16], empty[16];
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        fence_mbar_init();
    }
    __syncthreads();
    const long long t0 = clock64();
    if (warp == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int i = 0; i < iters; ++i) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&full[stage], box_bytes);
                const int row = hot ? 0 : (blockIdx.x * rows_per_cta + i) % (64 * 128);
                tma_load_4d(smem + stage * 17408, &tm, &full[stage], 0, -1, row % 128, row / 128);
            }
            __syncwarp();
            if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        int stage = 0; uint32_t phase = 0;
        for (int i = 0; i < iters; ++i) {
            mbar_wait(&full[stage], phase);
            if (elect_one()) mbar_arrive(&empty[stage]);
            __syncwarp();
            if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        if (elect_one() && blockIdx.x == 0) out[0] = clock64() - t0;
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    PFN_encodeTiled enc = (PFN_encodeTiled)p;
    const int N = 64, H = 128, W = 128;
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    for (int C : {192, 64}) {
        uint16_t* act; cudaMalloc(&act, (size_t)N * H * W * C * 2); cudaMemset(act, 0, (size_t)N * H * W * C * 2);
        for (int boxw : {130, 128}) {
            CUtensorMap tm;
            cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
            cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
            cuuint32_t box[4] = {64, (cuuint32_t)boxw, 1, 1}; cuuint32_t es[4] = {1, 1, 1, 1};
            enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, act, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            for (int hot : {1, 0}) for (int ns : {2, 4, 8, 12}) {
                const int iters = 55;
                bench<<<148, 128, 220 * 1024>>>(tm, ns, iters, boxw * 128, hot, 55, d);
                bench<<<148, 128, 220 * 1024>>>(tm, ns, iters, boxw * 128, hot, 55, d);
                long long c = 0;
                if (cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error\n"); return 1; }
                printf("C=%3d box=%d px %s stages=%2d : %.0f cycles/load  (%.1f B/cycle/SM)\n", C, boxw, hot ? "L2-hot " : "stream", ns,
                       double(c) / iters, boxw * 128.0 * iters / c);
            }
        }
        cudaFree(act);
    }
    return 0;
}
