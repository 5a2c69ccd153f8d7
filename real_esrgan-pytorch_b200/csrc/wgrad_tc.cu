// Weight gradient of a 3x3 convolution on the tensor cores (tcgen05 + TMEM):
//     dW[co][ci][dy][dx] = sum_{n,y,x} dY[n,y,x,co] * X[n, y+dy-1, x+dx-1, ci]        (autograd of model.py:75-79)
// as a K-major GEMM over pixels. Both operands are read from CHANNELS-FIRST bf16 copies (X^T [ci][n][y][x],
// dY^T [co][n][y][x]) so that a TMA box of 64 pixels x 128 (or 32) channels lands in shared memory as a standard
// K-major, 128B-swizzled operand tile (rows = channels, K = pixels). One pipeline stage = one 64-pixel segment of one
// image row:
//     A  = X^T tile            [128 ci x 64 px]                        (M = 128)
//     B  = 9 shifted dY^T tiles [(dx, co) x 64 px] for dy = 0..2         (N = 96 per dy, zero-filled outside the image;
//          the dx shift comes from three pre-shifted channels-first copies of dY: TMA cannot start a box at an odd
//          2-byte offset of its innermost dimension, the dy shift is a plain coordinate offset)
//     D[dy] (TMEM, 128 x 96 fp32) += A * B[dy]^T
// A CTA owns one (ci block, co slice) pair and a contiguous range of pixel segments (split-K); partial sums go to a
// workspace and a second kernel reduces the splits into the OIHW gradient.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "conv3x3.cuh"
#include "device_state.h"
#include "ptx.cuh"
#include "wgrad.cuh"

namespace resr {

static constexpr int kWgStageBytes = 16384 + 9 * 4096;  // 53,248 = 52 KB
static constexpr int kWgMaxStages = 4;
static constexpr uint32_t kWgDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t wg_desc(uint32_t lo) { return (static_cast<uint64_t>(kWgDescHi) << 32) | lo; }

__device__ __forceinline__ void wg_wait(uint64_t* bar, uint32_t parity, const WgradArgs&, unsigned) { mbar_wait(bar, parity); }

__global__ void __launch_bounds__(256, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmapX, const __grid_constant__ CUtensorMap tmapDY, const WgradArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* misc = smem + a.nstages * kWgStageBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(misc);
    uint64_t* empty = full + kWgMaxStages;
    uint64_t* done = empty + kWgMaxStages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int unit = blockIdx.x;               // (mb, cs)
    const int mb = a.nunits ? a.unit_mb[unit] : unit / a.n_cs, cs = a.nunits ? a.unit_cs[unit] : unit % a.n_cs;
    const int split = blockIdx.y;
    const long long k0 = a.kstages_total * split / gridDim.y;
    const long long k1 = a.kstages_total * (split + 1) / gridDim.y;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmapX);
        prefetch_tmap(&tmapDY);
        for (int i = 0; i < a.nstages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        mbar_init(done, 1);
        fence_mbar_init();
    }
    if (warp == 2) { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tmem_ptr;
    const int segs = a.segs_per_row;

    if (warp == 0) {
        int stage = 0; uint32_t phase = 0;
        // (xs, y, n) once, then incremented (64-bit divisions per stage were this loop's critical path)
        int xs = static_cast<int>(k0 % segs);
        int y = static_cast<int>((k0 / segs) % a.H);
        int n = static_cast<int>(k0 / (static_cast<long long>(segs) * a.H));
        for (long long k = k0; k < k1; ++k) {
            wg_wait(empty + stage, phase ^ 1, a, 1);
            if (elect_one()) {
                uint8_t* st = smem + stage * kWgStageBytes;
                mbar_expect_tx(full + stage, 16384 + 9 * 4096);
                tma_load_4d(st, &tmapX, full + stage, xs * 64, y, n, mb * 128);
                for (int dy = 0; dy < 3; ++dy)
                    for (int dx = 0; dx < 3; ++dx)
                        tma_load_4d(st + 16384 + (dy * 3 + dx) * 4096, &tmapDY, full + stage, xs * 64, y - dy + 1, n, dx * a.dy_rows + cs * 32);
            }
            __syncwarp();
            if (++stage == a.nstages) { stage = 0; phase ^= 1; }
            if (++xs == segs) { xs = 0; if (++y == a.H) { y = 0; ++n; } }
        }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc_f16(1, 128, 96);
        const uint32_t s_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
        int stage = 0; uint32_t phase = 0;
        uint32_t acc = 0;
        for (long long k = k0; k < k1; ++k) {
            wg_wait(full + stage, phase, a, 2);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_lo = s_lo + stage * (kWgStageBytes >> 4);
                const uint32_t b_lo = a_lo + (16384 >> 4);
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_f16(tbase + dy * 96, wg_desc(a_lo + ks * 2), wg_desc(b_lo + dy * (12288 >> 4) + ks * 2), idesc,
                                 (acc | ks) ? 1u : 0u);
                }
                umma_commit(empty + stage);
            }
            __syncwarp();
            acc = 1;
            if (++stage == a.nstages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(done);
        __syncwarp();
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int m = q * 32 + lane;  // ci row within the block
        wg_wait(done, 0, a, 3);
        tc_fence_after();
        float* dst = a.partial + (((static_cast<size_t>(split) * gridDim.x + unit) * 3) * 128 + m) * 96;
        const bool any = k1 > k0;
#pragma unroll 1
        for (int dy = 0; dy < 3; ++dy) {
            float* row = dst + static_cast<size_t>(dy) * 128 * 96;
#pragma unroll 1
            for (int c = 0; c < 3; ++c) {
                float v[32];
                tmem_ld32(tbase + (static_cast<uint32_t>(q * 32) << 16) + dy * 96 + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    reinterpret_cast<float4*>(row + c * 32)[i] =
                        any ? make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tbase, 512);
}

// dW[co][ci][dy][dx] = sum over splits of partial[split][mb][cs][dy][ci % 128][dx * 32 + co % 32]. One thread per partial
// element: the split-strided reads are coalesced, the OIHW write happens once.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int cin, int cout,
                                                          int n_mb, int n_cs, int nsplit) {
    const size_t stride = static_cast<size_t>(n_mb) * n_cs * 3 * 128 * 96;
    for (size_t e = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; e < stride; e += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int col = e % 96;
        size_t t = e / 96;
        const int cil = t % 128; t /= 128;
        const int dy = t % 3; t /= 3;
        const int cs = t % n_cs;
        const int mb = static_cast<int>(t / n_cs);
        const int ci = mb * 128 + cil, co = cs * 32 + (col & 31), dx = col >> 5;
        if (ci >= cin || co >= cout) continue;
        float s = 0.f;
        for (int sp = 0; sp < nsplit; ++sp) s += partial[sp * stride + e];
        dw[((static_cast<size_t>(co) * cin + ci) * 3 + dy) * 3 + dx] = s;
    }
}

// The same reduction for the five layers of a dense block at once (wgrad_launch_rdb): unit u = (mb, cs) from the explicit list,
// slice cs -> (layer gradient, cin, first output channel) from the table; the last 192 threads copy the bias gradients.
struct WgradUnits { int n; unsigned char mb[16], cs[16]; };
__global__ void __launch_bounds__(256) wgrad_reduce_rdb_kernel(const float* __restrict__ partial, const WgradRdbTable tb,
                                                              const WgradUnits un, int nsplit) {
    const size_t stride = static_cast<size_t>(un.n) * 3 * 128 * 96;
    for (size_t e = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; e < stride; e += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int col = e % 96;
        size_t t = e / 96;
        const int cil = t % 128; t /= 128;
        const int dy = t % 3;
        const int u = static_cast<int>(t / 3);
        const int mb = un.mb[u], cs = un.cs[u];
        const int ci = mb * 128 + cil, cin = tb.cin[cs], co = tb.co_base[cs] + (col & 31), dx = col >> 5;
        if (ci >= cin) continue;
        float s = 0.f;
        for (int sp = 0; sp < nsplit; ++sp) s += partial[sp * stride + e];
        tb.dw[cs][((static_cast<size_t>(co) * cin + ci) * 3 + dy) * 3 + dx] = s;
    }
    const size_t gt = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (gt < 192 && tb.dbcat) tb.db[gt >> 5][gt & 31] = tb.dbcat[gt];
}

// db[co] += sum_p dY^T[co][p]  (channels-first bf16); grid (cout, chunks), db must be zero-initialised.
__global__ void __launch_bounds__(256) bias_grad_kernel(const uint16_t* __restrict__ dyt, size_t P, int cout, float* __restrict__ db) {
    __shared__ float red[8];
    const int co = blockIdx.x;
    const uint16_t* row = dyt + static_cast<size_t>(co) * P;
    float s = 0.f;
    for (size_t p = blockIdx.y * static_cast<size_t>(blockDim.x) + threadIdx.x; p < P; p += static_cast<size_t>(gridDim.y) * blockDim.x)
        s += __uint_as_float(static_cast<uint32_t>(row[p]) << 16);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float w = 0.f;
        for (int i = 0; i < 8; ++i) w += red[i];
        atomicAdd(db + co, w);
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled wg_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// channels-first bf16 tensor [C][N][H][W]; box = 64 px x rows channels
// (function attributes are per device)
static int wgrad_set_smem_attr(int smem);

static int make_cf_map(CUtensorMap* out, const void* base, int C, int N, int H, int W, int rows) {
    PFN_encodeTiled enc = wg_encode_fn();
    if (!enc) return -1;
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(C)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(W) * 2, static_cast<cuuint64_t>(H) * W * 2, static_cast<cuuint64_t>(N) * H * W * 2};
    const cuuint32_t box[4] = {64, 1, 1, static_cast<cuuint32_t>(rows)};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(base), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

size_t wgrad_partial_bytes(int num_sms) {
    // worst case: 4 units (cin 192, cout 64) x (num_sms / 4) splits, or 1 unit x num_sms splits
    return static_cast<size_t>(num_sms + 8) * 3 * 128 * 96 * sizeof(float);
}

static int wgrad_set_smem_attr(int smem) {
    static PerDevice<bool> attr;
    if (!attr.cur()) {
        if (cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -3;
        attr.cur() = true;
    }
    return 0;
}

int wgrad_launch_rdb(const uint16_t* xt, const uint16_t* dyt, int N, int H, int W, float* partial, const WgradRdbTable& tb,
                     int num_sms, cudaStream_t s) {
    if (W % 8 != 0) return -2;
    WgradArgs a;
    memset(&a, 0, sizeof(a));
    a.N = N; a.H = H; a.W = W; a.cin = 192; a.cout = 192;
    a.n_mb = 2; a.n_cs = 6;
    a.segs_per_row = (W + 63) / 64;
    a.kstages_total = static_cast<long long>(N) * H * a.segs_per_row;
    a.partial = partial;
    a.nstages = kWgMaxStages;
    a.dy_rows = 192;
    WgradUnits un;
    memset(&un, 0, sizeof(un));
    // ci block 0 (channels 0..127) for every slice; ci block 1 (128..191) only where the layer has more than 128 inputs
    for (int cs = 0; cs < 6; ++cs) { un.mb[un.n] = 0; un.cs[un.n++] = static_cast<unsigned char>(cs); }
    for (int cs = 0; cs < 6; ++cs)
        if (tb.cin[cs] > 128) { un.mb[un.n] = 1; un.cs[un.n++] = static_cast<unsigned char>(cs); }
    a.nunits = un.n;
    memcpy(a.unit_mb, un.mb, 16);
    memcpy(a.unit_cs, un.cs, 16);
    long long nsplit = num_sms / un.n;
    if (nsplit > a.kstages_total / 16) nsplit = a.kstages_total / 16;
    if (nsplit < 1) nsplit = 1;
    CUtensorMap mx, my;
    int rc = make_cf_map(&mx, xt, 192, N, H, W, 128);
    rc |= make_cf_map(&my, dyt, 3 * 192, N, H, W, 32);
    if (rc != 0) return rc;
    const int smem = 1024 + kWgMaxStages * kWgStageBytes + 256;
    if (wgrad_set_smem_attr(smem) != 0) return -3;
    wgrad_tc_kernel<<<dim3(un.n, static_cast<unsigned>(nsplit)), 256, smem, s>>>(mx, my, a);
    const size_t stride = static_cast<size_t>(un.n) * 3 * 128 * 96;
    wgrad_reduce_rdb_kernel<<<static_cast<unsigned>((stride + 255) / 256), 256, 0, s>>>(partial, tb, un, static_cast<int>(nsplit));
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

int wgrad_launch(const uint16_t* xt, int x_channels, const uint16_t* dyt, int dy_channels, int N, int H, int W, int cin, int cout,
                 float* partial, float* dw, float* db, int num_sms, cudaStream_t s) {
    if (W % 8 != 0) return -2;  // TMA global strides must be multiples of 16 bytes
    WgradArgs a;
    memset(&a, 0, sizeof(a));
    a.N = N; a.H = H; a.W = W; a.cin = cin; a.cout = cout;
    a.n_mb = (cin + 127) / 128;
    a.n_cs = (cout + 31) / 32;
    a.segs_per_row = (W + 63) / 64;
    a.kstages_total = static_cast<long long>(N) * H * a.segs_per_row;
    a.partial = partial;
    a.nstages = kWgMaxStages;
    const int units = a.n_mb * a.n_cs;
    // split-K over pixels: at least ~16 pipeline steps per CTA (every split writes a 147 KB partial tile)
    long long nsplit = num_sms / units;
    if (nsplit > a.kstages_total / 16) nsplit = a.kstages_total / 16;
    if (nsplit < 1) nsplit = 1;
    CUtensorMap mx, my;
    int rc = make_cf_map(&mx, xt, x_channels, N, H, W, 128);
    rc |= make_cf_map(&my, dyt, 3 * dy_channels, N, H, W, 32);
    a.dy_rows = dy_channels;
    if (rc != 0) return rc;
    const int smem = 1024 + kWgMaxStages * kWgStageBytes + 256;
    if (wgrad_set_smem_attr(smem) != 0) return -3;
    wgrad_tc_kernel<<<dim3(units, static_cast<unsigned>(nsplit)), 256, smem, s>>>(mx, my, a);
    const size_t pstride = static_cast<size_t>(a.n_mb) * a.n_cs * 3 * 128 * 96;
    wgrad_reduce_kernel<<<static_cast<unsigned>((pstride + 255) / 256), 256, 0, s>>>(partial, dw, cin, cout, a.n_mb, a.n_cs,
                                                                                    static_cast<int>(nsplit));
    if (db) {  // from the unshifted copy (dx = 1); accumulates with atomics into a zeroed db
        const size_t P = static_cast<size_t>(N) * H * W;
        int chunks = static_cast<int>((P + 8191) / 8192);
        if (chunks > 64) chunks = 64;
        cudaMemsetAsync(db, 0, cout * sizeof(float), s);
        bias_grad_kernel<<<dim3(cout, chunks), 256, 0, s>>>(dyt + static_cast<size_t>(dy_channels) * P, P, cout, db);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace resr
