// Microbenchmark + layout probe for tcgen05.mma.cta_group::2 (CTA pair, M = 256) as the conv kernel wants to use it:
//   check   one K = 64 group of MMAs on small-integer operands; both CTAs dump their TMEM accumulators and the host
//           verifies D_cta[m, n] = sum_k A_cta[m, k] * B[n, k] with B rows [0, N/2) taken from CTA 0's shared memory and
//           [N/2, N) from CTA 1's (i.e. which CTA supplies which accumulator columns), at a non-zero column offset.
//   power   sustained pure MMA stream on all 74 pairs (random fp16 operands), prints TFLOP/s: pair N = 96 / 192 against
//           tools/mma_power (cta_group::1).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_pair tools/mma_pair.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../real_esrgan-pytorch_b200/csrc/ptx.cuh"
using namespace resr;
static constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t desc_of(uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; }

static constexpr int kABytes = 17408;          // 136 rows x 128 B
static constexpr int kBBytes = 3 * 128 * 128;  // 3 dx tiles of up to 128 rows (N/2 <= 128)

__host__ __device__ inline int a_val(int cta, int m, int k) { return ((m * 7 + k * 3 + cta * 5) % 5) - 2; }
__host__ __device__ inline int b_val(int n, int k) { return ((n * 11 + k * 5) % 7) - 3; }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pair_kernel(int N, int iters, unsigned seed, int check,
                                                                             float* dump) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const uint32_t rank = cluster_ctarank();
    uint8_t* A = smem;
    uint8_t* Bm = smem + kABytes;
    const int NH = N / 2;
    if (check) {
        for (int i = threadIdx.x; i < 136 * 64; i += blockDim.x) {
            const int m = i / 64, k = i % 64;
            const __half h = __float2half(m < 128 ? static_cast<float>(a_val(rank, m, k)) : 0.f);
            *reinterpret_cast<__half*>(A + m * 128 + ((((k >> 3) ^ (m & 7)) << 4) | ((k & 7) << 1))) = h;
        }
        for (int i = threadIdx.x; i < NH * 64; i += blockDim.x) {
            const int nl = i / 64, k = i % 64;
            const __half h = __float2half(static_cast<float>(b_val(rank * NH + nl, k)));
            *reinterpret_cast<__half*>(Bm + nl * 128 + ((((k >> 3) ^ (nl & 7)) << 4) | ((k & 7) << 1))) = h;
        }
    } else {
        unsigned st = seed + blockIdx.x * 7919u + threadIdx.x * 104729u;
        for (int i = threadIdx.x; i < (kABytes + kBBytes) / 4; i += blockDim.x) {
            st = st * 1664525u + 1013904223u;
            const uint32_t lo = 0x3800u | ((st >> 3) & 0x87FFu), hi = 0x3800u | ((st >> 17) & 0x87FFu);
            reinterpret_cast<uint32_t*>(smem)[i] = lo | (hi << 16);
        }
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc2(&tmem_ptr, 512); tmem_relinquish2(); }
    fence_proxy_async_smem();
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tbase = tmem_ptr;
    const int col0 = check ? 32 : 0;  // non-zero accumulator column offset in the layout probe
    if (warp == 1 && rank == 0) {
        const uint32_t a_lo = (smem_u32(A) & 0x3FFFFu) >> 4;
        const uint32_t b_lo = (smem_u32(Bm) & 0x3FFFFu) >> 4;
        const uint32_t idesc = make_idesc_f16(0, 256, N);
        const uint32_t WT = (NH * 128) >> 4;
        if (check) {
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma2_f16(tbase + col0, desc_of(a_lo + ks * 2), desc_of(b_lo + ks * 2), idesc, ks > 0);
            }
            __syncwarp();
        } else {
            const int nslots = 512 / N;
            for (int it = 0; it < iters; ++it) {
                const uint32_t d = tbase + (it % nslots) * N;
                if (elect_one()) {
#pragma unroll
                    for (int i = 0; i < 12; ++i) {
                        const int dx = i >> 2, ks = i & 3;
                        umma2_f16(d, desc_of(a_lo + dx * 8 + ks * 2), desc_of(b_lo + dx * WT + ks * 2), idesc, (it >= nslots) ? 1 : 0);
                    }
                }
                __syncwarp();
            }
        }
        if (elect_one()) umma2_commit_mc(smem_u32(&bar), 3);
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    if (check) {
        float v[32];
        for (int c = 0; c < N; c += 32) {
            tmem_ld32(tbase + (static_cast<uint32_t>(warp * 32) << 16) + col0 + c, v);
            tmem_ld_wait();
            const int m = warp * 32 + (threadIdx.x & 31);
            for (int j = 0; j < 32 && c + j < N; ++j) dump[(static_cast<size_t>(rank) * 128 + m) * N + c + j] = v[j];
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc2(tbase, 512);
}

int main(int argc, char** argv) {
    const char* what = argc > 1 ? argv[1] : "check";
    const int N = argc > 2 ? atoi(argv[2]) : 96;
    const double seconds = argc > 3 ? atof(argv[3]) : 3.0;
    const int smem = kABytes + kBBytes + 2048;
    cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (!strcmp(what, "check")) {
        float* dump;
        cudaMalloc(&dump, 2 * 128 * N * sizeof(float));
        cudaMemset(dump, 0, 2 * 128 * N * sizeof(float));
        pair_kernel<<<2, 128, smem>>>(N, 1, 1, 1, dump);
        const cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("check N=%d: CUDA error %s\n", N, cudaGetErrorString(e)); return 1; }
        std::vector<float> h(2 * 128 * N);
        cudaMemcpy(h.data(), dump, h.size() * 4, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int cta = 0; cta < 2; ++cta)
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    int ref = 0;
                    for (int k = 0; k < 64; ++k) ref += a_val(cta, m, k) * b_val(n, k);
                    if (h[(cta * 128 + m) * N + n] != static_cast<float>(ref)) {
                        if (bad < 8) printf("  mismatch cta %d m %d n %d: got %g want %d\n", cta, m, n, h[(cta * 128 + m) * N + n], ref);
                        ++bad;
                    }
                }
        printf("check N=%d (M=256 pair, D column offset 32): %ld mismatches of %d -> %s\n", N, bad, 2 * 128 * N,
               bad ? "LAYOUT ASSUMPTION WRONG" : "B rows [0,N/2) come from CTA0, [N/2,N) from CTA1; each CTA holds its own 128 rows x N");
        return bad ? 2 : 0;
    }
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    pair_kernel<<<148, 128, smem>>>(N, 100, 1, 0, nullptr);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    double total_ms = 0; long long launches = 0; float last = 0;
    while (total_ms < seconds * 1e3) {
        cudaEventRecord(e0);
        for (int k = 0; k < 4; ++k) pair_kernel<<<148, 128, smem>>>(N, iters, 1 + launches, 0, nullptr);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        cudaEventElapsedTime(&last, e0, e1);
        total_ms += last; launches += 4;
    }
    const double flop = 4.0 * 74 * iters * 12.0 * 2.0 * 256 * N * 16;
    printf("pair N=%d: last group %.2f ms -> %.0f TFLOP/s (cta_group::2, M=256, fp16 random data, pure MMA stream, 74 pairs)\n", N,
           last, flop / (last * 1e-3) / 1e12);
    return 0;
}
