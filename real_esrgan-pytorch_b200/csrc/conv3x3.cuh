// Internal interface of the tcgen05 3x3 convolution ("row-rolling implicit GEMM").
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace resr {

// Epilogue arithmetic selector (what the reference does after each nn.Conv2d).
enum EpMode : int {
    EP_PLAIN = 0,  // v = acc + b                       (model.py:90-93 conv1..4, :258, :264-269)
    EP_RDB = 1,    // v = (acc + b) * 0.2 + res1         (model.py:94-96)
    EP_RRDB = 2,   // v = ((acc + b) * 0.2 + res1) * 0.2 + res2   (model.py:94-96 then :129-130)
    EP_SKIP = 3,   // v = res1 + (acc + b)               (model.py:261-262)
    EP_ADD2 = 4,   // v = (acc + b) [+ res1] + res2_scale * res2   (backward: gradient accumulation)
};

struct ConvArgs {
    // geometry of this convolution (input resolution == output resolution, stride 1, zero pad 1)
    int N, H, W;
    int BW, BN;     // one 128-lane M tile = BW consecutive x  *  BN consecutive images
    int nxs;        // x segments per image row = ceil(W / BW)
    int ncg;        // column groups = ceil(N / BN) * nxs
    int nchunks;    // input channels / 64 (zero-padded weights cover the remainder)
    int tail_ksteps;  // K16 steps of the LAST chunk that carry real channels (1..4); the zero-weight rest is not issued
    int mode;       // 0: one TMA load of BW+2 px per (row, chunk), dx taken by shifting the smem descriptor
                    // 1: three TMA loads per (row, chunk), one per dx (any BW x BN split)
    int nstages;    // activation ring depth
    int nepi;       // epilogue warp groups (1 .. 3), rows alternate between them
    int reverse;    // pair kernel: traverse the work back to front (rows bottom-up, column groups last to first, dy taps
                    // flipped when the weights are gathered). Consecutive layers alternate direction, so a launch starts
                    // on the data its predecessor touched last -- what is still in the 126 MB L2.
    int epi_bufs;   // pair kernel: staging tiles per epilogue group (2 = a pass never waits for the previous pass's TMA store)
    int fmt_in;     // MMA operand format: 0 fp16, 1 bf16
    long long rows_total;  // ncg * H
    long long rows_total_pair;  // ceil(ncg / 2) * H: work units of the CTA-pair kernel (filled by its launcher)
    const uint8_t* wpack;  // [slice][chunk][dx][(dy,co) x 64ch] 16-bit, 128B-swizzled, ready for a bulk copy
    const float* bias;     // [nslices * NOUT]
    // epilogue
    int ep_mode, lrelu, clamp01;
    int has_out16;         // 16-bit NHWC output through tmapO16 (TMA store from a staging tile)
    int out16_fmt;         // 0 fp16, 1 bf16
    int out16_choff;       // first destination channel of slice 0
    int out16_up2;         // 1: destination is [N,2H,2W,*]; every pixel is written to its 2x2 nearest-upsampled sites
    int has_outf;          // fp32 NHWC output: each epilogue thread stores its pixel's 128 contiguous bytes directly
    int outf_choff;
    float* outf;           // base of the fp32 NHWC output, outf_cstride channels per pixel
    int outf_cstride;
    int has_res1;          // fp32 NHWC residual: each epilogue thread loads its pixel's channels straight into registers
    int res_choff;         //   before it waits for the accumulator (no shared-memory tile, no dependency on the TMA queue)
    const void* res1;
    int res1_cstride;
    int res16;             // 0: res1 / res2 are fp32; 1: both are 16-bit NHWC tensors in format res16_fmt (the residual
    int res16_fmt;         //    stream of the inference trunk IS the fp16 conv input: no fp32 master is kept)
    const void* res2;      // second residual (EP_RRDB / EP_ADD2): plain global loads
    int res2_cstride;
    float* out_nchw;       // NCHW fp32 output with out_nchw_c channels (or null)
    int out_nchw_c;
    unsigned char* out_u8; // NHWC u8 image output with out_nchw_c channels (or null): tensor_to_image fused into the last conv
                           // (imgproc.py:1594: mul(255).clamp(0, 255) then astype(uint8) = truncation)
    // ---- backward (data-gradient) extensions; all zero in the forward pass
    unsigned slice_nores_mask;  // bit s: Cout slice s does not read res1
    unsigned slice_noutf_mask;  // bit s: slice s does not write the fp32 output
    unsigned slice_no16_mask;   // bit s: slice s does not write the 16-bit output
    int out16_slice_fixed;      // 1: every writing slice stores at out16_choff (instead of out16_choff + slice * cout_slice)
    const void* mask16;         // LeakyReLU' mask source: NHWC 16-bit activations; v *= (act > 0 ? 1 : 0.2) before out16
    int mask16_cstride, mask16_choff;
    float res2_scale;           // EP_ADD2
    float out16_scale;          // != 0: the 16-bit output stores out16_scale * v (the fp32 output keeps v): the next dense block's dY5
    float* out_nchw_raw;        // pre-clamp copy of the NCHW output (training: clamp backward needs it)
    int dbg_flags;            // experiments only: 1 = skip output stores, 2 = producer re-reads row 0, 4 = issue 1/4 of the MMAs
    unsigned long long* dbg;  // optional: CTA (0,0) writes phase timestamps (globaltimer ns) here, 16 slots
};

// K16 steps of the last 64-channel chunk that hold real input channels.
inline int conv3x3_tail_ksteps(int cin) {
#ifdef RESR_NO_TAILSKIP
    (void)cin;
    return 4;
#else
    const int r = cin % 64;
    return r == 0 || r > 32 ? 4 : (r > 16 ? 2 : 1);
#endif
}

struct ConvMaps {
    CUtensorMap a;    // activations (load)
    CUtensorMap o16;  // 16-bit output (store), 4-D or 5-D (up2)
    CUtensorMap of;   // fp32 output (store)
};

// Launch one convolution. `cout_slice` is 32 or 16 (channels per CTA slice), `nslices` slices cover Cout.
cudaError_t conv3x3_launch(const ConvMaps& maps, const ConvArgs& args, int cout_slice, int nslices, int num_sms,
                           cudaStream_t stream);

// Fills nstages / nepi from the shared-memory budget. Returns false if the configuration does not fit.
bool conv3x3_plan_smem(ConvArgs* args, int cout_slice);

// ---- CTA-pair kernel (conv3x3_pair.cu): tcgen05 cta_group::2, M = 256; reads the same weight packs.
// nout: accumulator columns per output row of a pair (64 = two 32-channel pack slices per pair, 32, or 16).
bool conv3x3_pair_plan_smem(ConvArgs* args, int nout);
cudaError_t conv3x3_pair_launch(const ConvMaps& maps, const ConvArgs& args, int nout, int nslices, int num_sms,
                                cudaStream_t stream);

// Development builds (-DRESR_PROFILE_WAITS): copies / resets the pair kernel's wait-time counters (16 x u64 cycles).
int conv3x3_wait_profile(unsigned long long* out16, int reset);

// Kernel choice for one convolution whose weights are packed in `pack_nout`-channel slices (`pack_nslices` of them).
struct ConvLaunchCfg {
    int pair;     // 1: CTA-pair kernel, 0: single-CTA kernel
    int nout;     // channels per launch slice
    int nslices;  // blockIdx.y extent
};
// Picks the kernel (pairs whenever there are at least two column groups to pair up; RESR_CONV_PAIR=0 forces the
// single-CTA kernel), fills args->nstages / nepi. Returns false if the configuration does not fit in shared memory.
int conv3x3_set_pair_policy(int policy);  // returns the previous policy; -1 only queries
bool conv3x3_choose(ConvArgs* args, int pack_nout, int pack_nslices, ConvLaunchCfg* cfg);
cudaError_t conv3x3_run(const ConvMaps& maps, const ConvArgs& args, const ConvLaunchCfg& cfg, int num_sms, cudaStream_t stream);

// Tensor maps. base: NHWC tensor [N,H,W,C].
int conv3x3_make_tmap_act(CUtensorMap* out, const void* base, int N, int H, int W, int C, int mode, int BW, int BN,
                          int cvalid = 0);  // cvalid: channels the convolution reads (0 = all C)
int conv3x3_make_tmap_out16(CUtensorMap* out, const void* base, int N, int H, int W, int C, int nout, int BW, int BN,
                            int up2);  // N,H,W = geometry of the CONVOLUTION (destination is 2H x 2W when up2)
int conv3x3_make_tmap_f32(CUtensorMap* out, const void* base, int N, int H, int W, int C, int BW, int BN);

// Chooses the lane split for an image width.
void conv3x3_pick_tile(int W, int* BW, int* BN);

}  // namespace resr
