// Weight gradient of 3x3 convolutions on the tensor cores, operands read STRAIGHT from the NHWC buffers the forward and
// the data-gradient convolutions leave behind (no channels-first copies, no pre-shifted copies):
//     dW[co][ci][dy][dx] = sum_{n,y,x} dY[n,y,x,co] * X[n, y+dy-1, x+dx-1, ci]          (autograd of model.py:75-79)
// The contraction runs over pixels, and in an NHWC tensor the pixel index is the SLOW dimension of both operands: a TMA box
// of 64 channels x R pixels lands in shared memory as R rows of 128 bytes (128B swizzle), which is exactly the canonical
// MN-MAJOR operand layout of tcgen05.mma (cute/arch/mma_sm100_desc.hpp: ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units;
// LBO = distance between 64-channel chunks, SBO = 1024 B between groups of eight pixel rows). Both operands are MN-major
// (instruction-descriptor bits 15 / 16). A one-pixel shift along x is a 128-byte shift of the descriptor start address
// (the swizzle is a function of the absolute shared-memory address, the same property the forward kernel uses for its
// dx shifts), a shift along y is a TMA coordinate; borders are zero-filled by TMA.
//
// One pipeline stage = 64 pixels of one image row of dY (x0 .. x0+63) and the 66 pixels x0-1 .. x0+64 of row y+dy-1 of X:
//     A  = X tile   [128 ci x 66 px]   (two 64-channel chunks; channels past the tensor's end are zero-filled)   M = 128
//     B  = dY tile  [n co x 64 px]     (n = 64 or 128)
//     D[dx] (TMEM, 128 x n fp32) += A(rows dx .. dx+63) * B       for dx = 0, 1, 2  -> 3 accumulators, 3n <= 384 columns
// A CTA owns one unit (ci block, co chunk), ONE dy, and a contiguous range of pixel stages (split-K); its three
// accumulators go to a partial buffer and wgrad_mn_reduce_kernel sums the splits into the OIHW gradients of the layers
// that own the channel slices (a dense block's five layers share X and sit side by side in dYcat', train.cu).
// Requires X and dY in ONE 16-bit format (bf16 here: the bf16 recipe of the training path).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "conv3x3.cuh"
#include "device_state.h"
#include "ptx.cuh"
#include "wgrad.cuh"

namespace resr {

static constexpr int kMnXChunk = 72 * 128;                       // 66 rows loaded, rounded up to whole 8-row groups
static constexpr int kMnYChunk = 64 * 128;
static constexpr int kMnStageBytes = 2 * kMnXChunk + 2 * kMnYChunk;  // 34,816
static constexpr int kMnStages = 6;
static constexpr int kMnTileFloats = 3 * 128 * 128;              // one CTA's partial tile [dx][ci][co]

// MN-major, 128B-swizzled operand: start address | LBO (bits 16..29) in the low word; SBO = 1024 B, version 1, SWIZZLE_128B.
static constexpr uint32_t kMnDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t mn_desc(uint32_t addr16, uint32_t lbo16) {
    return (static_cast<uint64_t>(kMnDescHi) << 32) | (lbo16 << 16) | (addr16 & 0x3FFFu);
}

__global__ void __launch_bounds__(256, 1)
wgrad_mn_kernel(const __grid_constant__ CUtensorMap tmapX, const __grid_constant__ CUtensorMap tmapDY, const WgradMnArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* misc = smem + kMnStages * kMnStageBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(misc);
    uint64_t* empty = full + kMnStages;
    uint64_t* done = empty + kMnStages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    int kind = 0;
    while (kind + 1 < a.nkinds && static_cast<int>(blockIdx.x) >= a.kind[kind + 1].cta0) ++kind;
    const WgradMnKind kd = a.kind[kind];
    const int split = blockIdx.x - kd.cta0;
    const long long k0 = a.kslabs * split / kd.nsplit;
    const long long k1 = a.kslabs * (split + 1) / kd.nsplit;
    const int nyc = kd.n >> 6;  // 64-channel chunks of dY

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmapX);
        prefetch_tmap(&tmapDY);
        for (int i = 0; i < kMnStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
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
        for (long long k = k0; k < k1; ++k) {
            const int xs = static_cast<int>(k % segs);
            const int y = static_cast<int>((k / segs) % a.H);
            const int n = static_cast<int>(k / (static_cast<long long>(segs) * a.H));
            mbar_wait(empty + stage, phase ^ 1);
            if (elect_one()) {
                uint8_t* st = smem + stage * kMnStageBytes;
                mbar_expect_tx(full + stage, 2 * 66 * 128 + nyc * kMnYChunk);
                tma_load_4d(st, &tmapX, full + stage, kd.ci0, xs * 64 - 1, y + kd.dy - 1, n);
                tma_load_4d(st + kMnXChunk, &tmapX, full + stage, kd.ci0 + 64, xs * 64 - 1, y + kd.dy - 1, n);
                for (int j = 0; j < nyc; ++j)
                    tma_load_4d(st + 2 * kMnXChunk + j * kMnYChunk, &tmapDY, full + stage, kd.co0 + 64 * j, xs * 64, y, n);
            }
            __syncwarp();
            if (++stage == kMnStages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc_f16(1, 128, kd.n) | (1u << 15) | (1u << 16);
        const uint32_t s16 = (smem_u32(smem) & 0x3FFFFu) >> 4;
        int stage = 0; uint32_t phase = 0;
        uint32_t acc = 0;
        for (long long k = k0; k < k1; ++k) {
            mbar_wait(full + stage, phase);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t xa = s16 + stage * (kMnStageBytes >> 4);
                const uint32_t ya = xa + ((2 * kMnXChunk) >> 4);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t bd = mn_desc(ya + ks * (2048 >> 4), kMnYChunk >> 4);
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx)
                        umma_f16(tbase + dx * kd.n, mn_desc(xa + ks * (2048 >> 4) + dx * (128 >> 4), kMnXChunk >> 4), bd, idesc,
                                 (acc | ks) ? 1u : 0u);
                }
                umma_commit(empty + stage);
            }
            __syncwarp();
            acc = 1;
            if (++stage == kMnStages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(done);
        __syncwarp();
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int m = q * 32 + lane;  // ci row within the unit
        mbar_wait(done, 0);
        tc_fence_after();
        float* dst = a.partial + static_cast<size_t>(blockIdx.x) * kMnTileFloats + static_cast<size_t>(m) * 128;
        const bool any = k1 > k0;
#pragma unroll 1
        for (int dx = 0; dx < 3; ++dx) {
            float* row = dst + static_cast<size_t>(dx) * 128 * 128;
#pragma unroll 1
            for (int c = 0; c < kd.n; c += 32) {
                float v[32];
                tmem_ld32(tbase + (static_cast<uint32_t>(q * 32) << 16) + dx * kd.n + c, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    reinterpret_cast<float4*>(row + c)[i] =
                        any ? make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tbase, 512);
}

// dW of the layer owning each 32-channel slice of dY: sum over the splits of a kind. One thread per partial element
// (kind, dx, ci, co): the split-strided reads are coalesced along co, the OIHW write happens once. The first `nbias`
// threads also finish the bias gradients from the per-block column sums.
__global__ void __launch_bounds__(256) wgrad_mn_reduce_kernel(const WgradMnArgs a, const WgradMnTable tb, const float* __restrict__ colsum,
                                                             int ncolblocks, int nbias) {
    const size_t total = static_cast<size_t>(a.nkinds) * kMnTileFloats;
    for (size_t e = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; e < total; e += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int col = e & 127;
        const int m = (e >> 7) & 127;
        const int dx = static_cast<int>((e >> 14) % 3);
        const int kind = static_cast<int>(e / kMnTileFloats);
        const WgradMnKind kd = a.kind[kind];
        if (col >= kd.n) continue;
        const int ci = kd.ci0 + m, cc = kd.co0 + col, cs = cc >> 5;
        const int co = tb.co_base[cs] + (cc & 31);
        if (ci >= tb.cin[cs] || co >= tb.cout[cs]) continue;
        const float* p = a.partial + static_cast<size_t>(kd.cta0) * kMnTileFloats + (e - static_cast<size_t>(kind) * kMnTileFloats);
        float s = 0.f;
        for (int sp = 0; sp < kd.nsplit; ++sp) s += p[static_cast<size_t>(sp) * kMnTileFloats];
        tb.dw[cs][((static_cast<size_t>(co) * tb.cin[cs] + ci) * 3 + kd.dy) * 3 + dx] = s;
    }
    const size_t gt = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (gt < static_cast<size_t>(nbias) && colsum) {
        const int cs = static_cast<int>(gt >> 5);
        const int co = tb.co_base[cs] + static_cast<int>(gt & 31);
        if (co < tb.cout[cs]) {
            float s = 0.f;
            for (int b = 0; b < ncolblocks; ++b) s += colsum[static_cast<size_t>(b) * 192 + gt];
            tb.db[cs][co] = s;
        }
    }
}

// Per-block column sums of a 16-bit NHWC gradient buffer: out[block][c] = sum over the block's pixels of dy[p][choff + c]
// (c < C <= 192, C % 8 == 0). 16-byte loads, fixed summation order (the bias gradient is bit-reproducible).
__global__ void __launch_bounds__(384) colsum_bf16_kernel(const uint16_t* __restrict__ dy, int cstride, int choff, int C, size_t P,
                                                         float* __restrict__ out) {
    __shared__ float red[16][192];
    const int tpr = C >> 3;                 // threads per pixel row
    const int rows = 384 / tpr;             // pixel rows in flight
    const int r = threadIdx.x / tpr, t = threadIdx.x % tpr;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const size_t per = (P + gridDim.x - 1) / gridDim.x;
    const size_t p0 = blockIdx.x * per, p1 = p0 + per < P ? p0 + per : P;
    if (r < rows && r < 16) {
        const int rr = rows < 16 ? rows : 16;
        for (size_t p = p0 + r; p < p1; p += rr) {
            const uint4 v = *reinterpret_cast<const uint4*>(dy + p * cstride + choff + t * 8);
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[2 * j] += __uint_as_float(w4[j] << 16);
                acc[2 * j + 1] += __uint_as_float(w4[j] & 0xFFFF0000u);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) red[r][t * 8 + j] = acc[j];
    }
    __syncthreads();
    if (static_cast<int>(threadIdx.x) < C) {
        const int rr = rows < 16 ? rows : 16;
        float s = 0.f;
        for (int i = 0; i < rr; ++i) s += red[i][threadIdx.x];
        out[static_cast<size_t>(blockIdx.x) * 192 + threadIdx.x] = s;
    }
}

typedef CUresult (*PFN_encodeTiledMn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiledMn mn_encode_fn() {
    static PFN_encodeTiledMn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiledMn>(p);
    }
    return fn;
}

// NHWC 16-bit tensor [N][H][W][cstride], channels [0, C) visible (the rest of a 64-channel box is zero-filled);
// box = 64 channels x `rows` pixels of one image row.
static int make_nhwc_map(CUtensorMap* out, const void* base, int C, int cstride, int N, int H, int W, int rows) {
    PFN_encodeTiledMn enc = mn_encode_fn();
    if (!enc) return -1;
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(cstride) * 2, static_cast<cuuint64_t>(W) * cstride * 2,
                                   static_cast<cuuint64_t>(H) * W * cstride * 2};
    const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(rows), 1, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(base), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

static constexpr int kMnColBlocks = 128;

size_t wgrad_mn_workspace_bytes(int num_sms) {
    return static_cast<size_t>(num_sms + 16) * kMnTileFloats * sizeof(float) + static_cast<size_t>(kMnColBlocks) * 192 * sizeof(float);
}

static int wgrad_mn_set_smem_attr(int smem) {
    static PerDevice<bool> attr;
    if (!attr.cur()) {
        if (cudaFuncSetAttribute(wgrad_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -3;
        attr.cur() = true;
    }
    return 0;
}

// units: (ci0, co0, n) triples; every unit is expanded into its three dy kinds. CTAs are dealt to the kinds in proportion
// to their per-stage cost (MMA columns + the fixed X tile), every kind gets at least one.
int wgrad_mn_launch(const uint16_t* x, int x_cstride, int x_channels, const uint16_t* dy, int dy_cstride, int dy_channels,
                    int N, int H, int W, const int (*units)[3], int nunits, const WgradMnTable& tb, bool with_bias,
                    float* workspace, int num_sms, cudaStream_t s) {
    if (nunits < 1 || nunits * 3 > kMnMaxKinds) return -2;
    if ((x_cstride & 7) || (dy_cstride & 7) || (dy_channels & 7) || dy_channels > 192) return -2;
    WgradMnArgs a;
    memset(&a, 0, sizeof(a));
    a.N = N; a.H = H; a.W = W;
    a.segs_per_row = (W + 63) / 64;
    a.kslabs = static_cast<long long>(N) * H * a.segs_per_row;
    a.partial = workspace;
    a.nkinds = nunits * 3;
    int weight[kMnMaxKinds], share[kMnMaxKinds];
    long long wsum = 0;
    for (int u = 0; u < nunits; ++u)
        for (int d = 0; d < 3; ++d) {
            WgradMnKind& k = a.kind[u * 3 + d];
            k.ci0 = units[u][0]; k.co0 = units[u][1]; k.n = units[u][2]; k.dy = d;
            if (k.n != 64 && k.n != 128) return -2;
            weight[u * 3 + d] = k.n + 32;
            wsum += weight[u * 3 + d];
        }
    long long cap = a.kslabs / 4;
    if (cap < 1) cap = 1;
    int used = 0;
    for (int i = 0; i < a.nkinds; ++i) {
        long long v = static_cast<long long>(num_sms) * weight[i] / wsum;
        if (v < 1) v = 1;
        if (v > cap) v = cap;
        share[i] = static_cast<int>(v);
        used += share[i];
    }
    // hand the SMs left over by the rounding to the heaviest kinds first
    for (int pass = 0; pass < 2 && used < num_sms; ++pass)
        for (int i = 0; i < a.nkinds && used < num_sms; ++i)
            if (weight[i] >= (pass == 0 ? 160 : 0) && share[i] < cap) { ++share[i]; ++used; }
    int cta = 0;
    for (int i = 0; i < a.nkinds; ++i) { a.kind[i].cta0 = cta; a.kind[i].nsplit = share[i]; cta += share[i]; }
    if (cta > num_sms + 16) return -2;
    CUtensorMap mx, my;
    int rc = make_nhwc_map(&mx, x, x_channels, x_cstride, N, H, W, 66);
    rc |= make_nhwc_map(&my, dy, dy_channels, dy_cstride, N, H, W, 64);
    if (rc != 0) return rc;
    const int smem = 1024 + kMnStages * kMnStageBytes + 256;
    if (wgrad_mn_set_smem_attr(smem) != 0) return -3;
    float* colsum = workspace + static_cast<size_t>(num_sms + 16) * kMnTileFloats;
    const size_t P = static_cast<size_t>(N) * H * W;
    int colblocks = static_cast<int>((P + 255) / 256);
    if (colblocks > kMnColBlocks) colblocks = kMnColBlocks;
    if (with_bias) colsum_bf16_kernel<<<colblocks, 384, 0, s>>>(dy, dy_cstride, 0, dy_channels, P, colsum);
    wgrad_mn_kernel<<<cta, 256, smem, s>>>(mx, my, a);
    const size_t total = static_cast<size_t>(a.nkinds) * kMnTileFloats;
    wgrad_mn_reduce_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(a, tb, with_bias ? colsum : nullptr, colblocks,
                                                                                     with_bias ? dy_channels : 0);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace resr
