// 3x3 / stride 1 / zero-pad 1 convolution on the 5th-gen tensor cores (tcgen05 + TMEM), NHWC 16-bit activations.
//
// Replaces nn.Conv2d(…, (3,3), (1,1), (1,1)) + torch.cat + LeakyReLU + residual arithmetic of the reference
// generator (/root/reference/model.py:75-98, 123-132, 255-272). Design ("row-rolling implicit GEMM"):
//
//   * One M tile = 128 output columns of ONE image row (or BW columns x BN images when W < 128): TMEM lane = pixel.
//   * The three vertical taps are folded into the MMA N dimension: for input row r the tensor core computes
//         D'_r[x, (dy, co)] = sum_{dx, ci} in[r, x + dx - 1, ci] * w[co, ci, dy, dx]            (N = 3 * Cout_slice)
//     so every activation row is read from shared memory once per dx instead of once per tap, and
//         out[y] = D'_{y-1}[dy=0] + D'_y[dy=1] + D'_{y+1}[dy=2]
//     is a same-lane sum that the epilogue warps keep in registers while rows roll through a TMEM ring.
//   * Horizontal taps: mode 0 loads BW+2 pixels once (TMA zero-fills x = -1 and x = W) and shifts the UMMA
//     shared-memory descriptor by dx * 128 B; mode 1 issues one TMA load per dx (used when the lanes span images).
//   * Weights of the CTA's Cout slice stay resident in shared memory for the whole launch (one bulk copy).
//   * CTAs are persistent over a contiguous range of (column group, row) work; grid = #SMs / #slices.
//
// Warp roles (256 threads): warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer, warp 2 = TMEM allocator,
// warps 4-7 = epilogue (TMEM -> registers -> bias / LeakyReLU / residual -> global).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "conv3x3.cuh"
#include "ptx.cuh"

namespace resr {

static constexpr int kStageBytes = 17408;  // 136 rows x 128 B (mode 0 uses 130 rows, mode 1 uses 128)
static constexpr int kMaxStages = 12;
static constexpr int kTmemSlots = 4;
static constexpr int kTmemSlotCols = 128;
static constexpr int kMiscBytes = 1024;
static constexpr int kSmemMax = 232448;  // 227 KB opt-in limit per CTA

struct RowRange {
    long long g0, g1;
};

__device__ __forceinline__ RowRange cta_rows(const ConvArgs& a) {
    RowRange r;
    r.g0 = a.rows_total * static_cast<long long>(blockIdx.x) / gridDim.x;
    r.g1 = a.rows_total * static_cast<long long>(blockIdx.x + 1) / gridDim.x;
    return r;
}

template <int NOUT>
struct Emit {
    const ConvArgs& a;
    const float* bias_s;
    int slice;
    __device__ __forceinline__ void operator()(int n, int y, int x, float* v) const {
        const size_t pix = (static_cast<size_t>(n) * a.H + y) * a.W + x;
#pragma unroll
        for (int i = 0; i < NOUT; ++i) v[i] = __fadd_rn(v[i], bias_s[i]);
        if (a.ep_mode != EP_PLAIN) {
            const float4* r1 = reinterpret_cast<const float4*>(a.res1 + pix * a.res_cstride + a.res_choff + slice * NOUT);
#pragma unroll
            for (int i = 0; i < NOUT / 4; ++i) {
                const float4 q = __ldg(r1 + i);
                if (a.ep_mode == EP_SKIP) {
                    v[4 * i + 0] = __fadd_rn(q.x, v[4 * i + 0]);
                    v[4 * i + 1] = __fadd_rn(q.y, v[4 * i + 1]);
                    v[4 * i + 2] = __fadd_rn(q.z, v[4 * i + 2]);
                    v[4 * i + 3] = __fadd_rn(q.w, v[4 * i + 3]);
                } else {
                    v[4 * i + 0] = __fadd_rn(__fmul_rn(v[4 * i + 0], 0.2f), q.x);
                    v[4 * i + 1] = __fadd_rn(__fmul_rn(v[4 * i + 1], 0.2f), q.y);
                    v[4 * i + 2] = __fadd_rn(__fmul_rn(v[4 * i + 2], 0.2f), q.z);
                    v[4 * i + 3] = __fadd_rn(__fmul_rn(v[4 * i + 3], 0.2f), q.w);
                }
            }
            if (a.ep_mode == EP_RRDB) {
                const float4* r2 =
                    reinterpret_cast<const float4*>(a.res2 + pix * a.res_cstride + a.res_choff + slice * NOUT);
#pragma unroll
                for (int i = 0; i < NOUT / 4; ++i) {
                    const float4 q = __ldg(r2 + i);
                    v[4 * i + 0] = __fadd_rn(__fmul_rn(v[4 * i + 0], 0.2f), q.x);
                    v[4 * i + 1] = __fadd_rn(__fmul_rn(v[4 * i + 1], 0.2f), q.y);
                    v[4 * i + 2] = __fadd_rn(__fmul_rn(v[4 * i + 2], 0.2f), q.z);
                    v[4 * i + 3] = __fadd_rn(__fmul_rn(v[4 * i + 3], 0.2f), q.w);
                }
            }
        }
        if (a.lrelu) {
#pragma unroll
            for (int i = 0; i < NOUT; ++i) v[i] = v[i] > 0.f ? v[i] : __fmul_rn(v[i], 0.2f);
        }
        if (a.clamp01) {
#pragma unroll
            for (int i = 0; i < NOUT; ++i) v[i] = fminf(fmaxf(v[i], 0.f), 1.f);
        }
        if (a.outf) {
            float4* o = reinterpret_cast<float4*>(a.outf + pix * a.outf_cstride + a.outf_choff + slice * NOUT);
#pragma unroll
            for (int i = 0; i < NOUT / 4; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        if (a.out16) {
            uint32_t pk[NOUT / 2];
#pragma unroll
            for (int i = 0; i < NOUT / 2; ++i) {
                if (a.out16_fmt == 1) {
                    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                    pk[i] = *reinterpret_cast<uint32_t*>(&h);
                } else {
                    __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                    pk[i] = *reinterpret_cast<uint32_t*>(&h);
                }
            }
            uint16_t* base = reinterpret_cast<uint16_t*>(a.out16);
            if (!a.out16_up2) {
                uint4* o = reinterpret_cast<uint4*>(base + pix * a.out16_cstride + a.out16_choff + slice * NOUT);
#pragma unroll
                for (int i = 0; i < NOUT / 8; ++i) o[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
            } else {
                const int Ho = 2 * a.H, Wo = 2 * a.W;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const size_t po = (static_cast<size_t>(n) * Ho + 2 * y + (s >> 1)) * Wo + 2 * x + (s & 1);
                    uint4* o = reinterpret_cast<uint4*>(base + po * a.out16_cstride + a.out16_choff + slice * NOUT);
#pragma unroll
                    for (int i = 0; i < NOUT / 8; ++i)
                        o[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
                }
            }
        }
        if (a.out_nchw) {
            const size_t plane = static_cast<size_t>(a.H) * a.W;
            float* o = a.out_nchw + static_cast<size_t>(n) * a.out_nchw_c * plane + static_cast<size_t>(y) * a.W + x;
#pragma unroll
            for (int c = 0; c < NOUT; ++c) {
                const int cc = slice * NOUT + c;
                if (cc < a.out_nchw_c) o[static_cast<size_t>(cc) * plane] = v[c];
            }
        }
    }
};

template <int NOUT>
__global__ void __launch_bounds__(256, 1) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmapA, const ConvArgs a) {
    constexpr int NT = 3 * NOUT;        // MMA N: (dy, co)
    constexpr int WTILE = NT * 128;     // bytes of one (chunk, dx) weight tile: NT rows x 64 ch x 2 B
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int slice = blockIdx.y;
    const uint32_t wbytes = static_cast<uint32_t>(a.nchunks) * 3u * WTILE;

    uint8_t* wsm = smem;
    uint8_t* stg = smem + wbytes;
    uint8_t* misc = stg + a.nstages * kStageBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(misc);
    uint64_t* empty = full + kMaxStages;
    uint64_t* tfull = empty + kMaxStages;
    uint64_t* tempty = tfull + kTmemSlots;
    uint64_t* wbar = tempty + kTmemSlots;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(wbar + 1);
    float* bias_s = reinterpret_cast<float*>(misc + 512);

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmapA);
        for (int i = 0; i < a.nstages; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, 1);
        }
        for (int i = 0; i < kTmemSlots; ++i) {
            mbar_init(tfull + i, 1);
            mbar_init(tempty + i, 128);
        }
        mbar_init(wbar, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    if (threadIdx.x >= 128 && threadIdx.x < 128 + NOUT) bias_s[threadIdx.x - 128] = a.bias[slice * NOUT + threadIdx.x - 128];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tmem_ptr;

    const RowRange rr = cta_rows(a);
    const int H = a.H;

    if (threadIdx.x == 0) {
        // ------------------------------------------------------------------ TMA producer
        mbar_expect_tx(wbar, wbytes);
        const uint8_t* wsrc = a.wpack + static_cast<size_t>(slice) * wbytes;
        for (int c = 0; c < a.nchunks; ++c) bulk_load_1d(wsm + c * 3 * WTILE, wsrc + static_cast<size_t>(c) * 3 * WTILE, 3 * WTILE, wbar);
        int stage = 0;
        uint32_t phase = 0;
        for (long long g = rr.g0; g < rr.g1;) {
            const int cg = static_cast<int>(g / H);
            const int ya = static_cast<int>(g % H);
            const int yb = static_cast<int>(min(static_cast<long long>(H), ya + (rr.g1 - g)));
            const int n0 = (cg / a.nxs) * a.BN;
            const int x0 = (cg % a.nxs) * a.BW;
            const int ra = max(ya - 1, 0), rb = min(yb, H - 1);
            for (int r = ra; r <= rb; ++r) {
                for (int c = 0; c < a.nchunks; ++c) {
                    if (a.mode == 0) {
                        mbar_wait(empty + stage, phase ^ 1);
                        mbar_expect_tx(full + stage, (a.BW + 2) * 128);
                        tma_load_4d(stg + stage * kStageBytes, &tmapA, full + stage, c * 64, x0 - 1, r, n0);
                        if (++stage == a.nstages) { stage = 0; phase ^= 1; }
                    } else {
                        for (int dx = 0; dx < 3; ++dx) {
                            mbar_wait(empty + stage, phase ^ 1);
                            mbar_expect_tx(full + stage, 128 * 128);
                            tma_load_4d(stg + stage * kStageBytes, &tmapA, full + stage, c * 64, x0 + dx - 1, r, n0);
                            if (++stage == a.nstages) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
            g += yb - ya;
        }
    } else if (threadIdx.x == 32) {
        // ------------------------------------------------------------------ MMA issuer
        mbar_wait(wbar, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_f16(a.fmt_in, 128, NT);
        const uint32_t w_addr = smem_u32(wsm);
        const uint32_t s_addr = smem_u32(stg);
        int stage = 0;
        uint32_t phase = 0;
        uint32_t it = 0;
        for (long long g = rr.g0; g < rr.g1;) {
            const int ya = static_cast<int>(g % H);
            const int yb = static_cast<int>(min(static_cast<long long>(H), ya + (rr.g1 - g)));
            const int ra = max(ya - 1, 0), rb = min(yb, H - 1);
            for (int r = ra; r <= rb; ++r, ++it) {
                const uint32_t slot = it % kTmemSlots;
                mbar_wait(tempty + slot, ((it / kTmemSlots) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_addr = tbase + slot * kTmemSlotCols;
                uint32_t acc = 0;
                for (int c = 0; c < a.nchunks; ++c) {
                    if (a.mode == 0) {
                        mbar_wait(full + stage, phase);
                        tc_fence_after();
                        const uint32_t abase = s_addr + stage * kStageBytes;
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const uint32_t bbase = w_addr + (c * 3 + dx) * WTILE;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                umma_f16(d_addr, make_smem_desc(abase + dx * 128 + ks * 32, 1024, 128),
                                         make_smem_desc(bbase + ks * 32, 1024, 128), idesc, acc);
                                acc = 1;
                            }
                        }
                        umma_commit(empty + stage);
                        if (++stage == a.nstages) { stage = 0; phase ^= 1; }
                    } else {
                        for (int dx = 0; dx < 3; ++dx) {
                            mbar_wait(full + stage, phase);
                            tc_fence_after();
                            const uint32_t abase = s_addr + stage * kStageBytes;
                            const uint32_t bbase = w_addr + (c * 3 + dx) * WTILE;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                umma_f16(d_addr, make_smem_desc(abase + ks * 32, 1024, 128),
                                         make_smem_desc(bbase + ks * 32, 1024, 128), idesc, acc);
                                acc = 1;
                            }
                            umma_commit(empty + stage);
                            if (++stage == a.nstages) { stage = 0; phase ^= 1; }
                        }
                    }
                }
                umma_commit(tfull + slot);
            }
            g += yb - ya;
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue warps
        const int q = warp - 4;            // TMEM lane quadrant
        const int m = q * 32 + lane;       // M row == TMEM lane == pixel of the tile
        const int img_in_tile = m / a.BW;
        const int x_in_tile = m % a.BW;
        const Emit<NOUT> emit{a, bias_s, slice};
        uint32_t it = 0;
        for (long long g = rr.g0; g < rr.g1;) {
            const int cg = static_cast<int>(g / H);
            const int ya = static_cast<int>(g % H);
            const int yb = static_cast<int>(min(static_cast<long long>(H), ya + (rr.g1 - g)));
            const int n = (cg / a.nxs) * a.BN + img_in_tile;
            const int x = (cg % a.nxs) * a.BW + x_in_tile;
            const bool valid = (n < a.N) && (x < a.W);
            const int ra = max(ya - 1, 0), rb = min(yb, H - 1);
            float s0[NOUT], s1[NOUT];
#pragma unroll
            for (int i = 0; i < NOUT; ++i) { s0[i] = 0.f; s1[i] = 0.f; }
            for (int r = ra; r <= rb; ++r, ++it) {
                const uint32_t slot = it % kTmemSlots;
                mbar_wait(tfull + slot, (it / kTmemSlots) & 1);
                tc_fence_after();
                const uint32_t taddr = tbase + (static_cast<uint32_t>(q * 32) << 16) + slot * kTmemSlotCols;
                float fin[NOUT], t[NOUT];
                if (NOUT == 32) tmem_ld32(taddr + 2 * NOUT, fin); else tmem_ld16(taddr + 2 * NOUT, fin);
                if (NOUT == 32) tmem_ld32(taddr + NOUT, t); else tmem_ld16(taddr + NOUT, t);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < NOUT; ++i) {
                    fin[i] = __fadd_rn(s1[i], fin[i]);   // out[r-1] = (D'_{r-2}[0] + D'_{r-1}[1]) + D'_r[2]
                    s1[i] = __fadd_rn(s0[i], t[i]);      // partial of out[r]
                }
                if (NOUT == 32) tmem_ld32(taddr, s0); else tmem_ld16(taddr, s0);   // partial of out[r+1]
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(tempty + slot);
                if (valid && r - 1 >= ya) emit(n, r - 1, x, fin);
            }
            if (valid && yb == H && rb >= ra) emit(n, H - 1, x, s1);
            g += yb - ya;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tbase, 512);
}

// --------------------------------------------------------------------------------------------- host side

int conv3x3_pick_stages(int nchunks, int cout_slice) {
    const int wbytes = nchunks * 3 * (3 * cout_slice) * 128;
    int ns = (kSmemMax - 1024 /*alignment slack*/ - kMiscBytes - wbytes) / kStageBytes;
    if (ns > kMaxStages) ns = kMaxStages;
    return ns;
}

void conv3x3_pick_tile(int W, int* BW, int* BN) {
    if (W == 64 || W == 32 || W == 16 || W == 8) {
        *BW = W;
        *BN = 128 / W;
    } else {
        *BW = 128;
        *BN = 1;
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

int conv3x3_make_tmap(CUtensorMap* out, const void* base, int N, int H, int W, int C, int mode, int BW, int BN) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return -1;
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                                static_cast<cuuint64_t>(N)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2,
                                   static_cast<cuuint64_t>(H) * W * C * 2};
    cuuint32_t box[4];
    box[0] = 64;
    if (mode == 0) {
        box[1] = static_cast<cuuint32_t>(BW + 2);
        box[2] = 1;
        box[3] = 1;
    } else {
        box[1] = static_cast<cuuint32_t>(BW);
        box[2] = 1;
        box[3] = static_cast<cuuint32_t>(BN);
    }
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

cudaError_t conv3x3_launch(const CUtensorMap& tmapA, const ConvArgs& args, int cout_slice, int nslices, int num_sms,
                           cudaStream_t stream) {
    const int wbytes = args.nchunks * 3 * (3 * cout_slice) * 128;
    const int smem = 1024 + wbytes + args.nstages * kStageBytes + kMiscBytes;
    if (args.nstages < 2 || smem > kSmemMax) return cudaErrorInvalidConfiguration;
    // Always request the full 227 KB so that exactly one CTA (one 512-column TMEM allocation) lives on an SM.
    const int smem_req = kSmemMax;
    long long gx = num_sms / nslices;
    const long long min_rows = 4;  // do not shred tiny problems into 1-row strips (2 halo rows each)
    const long long cap = (args.rows_total + min_rows - 1) / min_rows;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(nslices), 1);
    cudaError_t e;
    if (cout_slice == 32) {
        static bool attr32 = false;
        if (!attr32) {
            e = cudaFuncSetAttribute(conv3x3_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
            if (e != cudaSuccess) return e;
            attr32 = true;
        }
        conv3x3_tc_kernel<32><<<grid, 256, smem_req, stream>>>(tmapA, args);
    } else if (cout_slice == 16) {
        static bool attr16 = false;
        if (!attr16) {
            e = cudaFuncSetAttribute(conv3x3_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
            if (e != cudaSuccess) return e;
            attr16 = true;
        }
        conv3x3_tc_kernel<16><<<grid, 256, smem_req, stream>>>(tmapA, args);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace resr
