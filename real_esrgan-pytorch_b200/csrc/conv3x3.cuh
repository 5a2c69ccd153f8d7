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
};

struct ConvArgs {
    // geometry of this convolution (input resolution == output resolution, stride 1, zero pad 1)
    int N, H, W;
    int BW, BN;     // one 128-lane M tile = BW consecutive x  *  BN consecutive images
    int nxs;        // x segments per image row = ceil(W / BW)
    int ncg;        // column groups = ceil(N / BN) * nxs
    int nchunks;    // input channels / 64 (zero-padded weights cover the remainder)
    int mode;       // 0: one TMA load of BW+2 px per (row, chunk), dx taken by shifting the smem descriptor
                    // 1: three TMA loads per (row, chunk), one per dx (any BW x BN split)
    int nstages;    // activation ring depth
    int fmt_in;     // MMA operand format: 0 fp16, 1 bf16
    long long rows_total;  // ncg * H
    const uint8_t* wpack;  // [slice][chunk][dx][(dy,co) x 64ch] 16-bit, 128B-swizzled, ready for a bulk copy
    const float* bias;     // [nslices * NOUT]
    // epilogue
    int ep_mode, lrelu, clamp01;
    void* out16;           // NHWC 16-bit output (or null)
    int out16_fmt;         // 0 fp16, 1 bf16
    int out16_cstride;     // channels per pixel of the destination tensor
    int out16_choff;       // first destination channel of slice 0
    int out16_up2;         // 1: destination is [N,2H,2W,*]; every pixel is written to its 2x2 nearest-upsampled sites
    float* outf;           // NHWC fp32 output (or null)
    int outf_cstride, outf_choff;
    const float* res1;     // NHWC fp32 residual inputs
    const float* res2;
    int res_cstride, res_choff;
    float* out_nchw;       // NCHW fp32 output with out_nchw_c channels (or null)
    int out_nchw_c;
};

// Launch one convolution. `cout_slice` is 32 or 16 (channels per CTA slice), `nslices` slices cover Cout.
cudaError_t conv3x3_launch(const CUtensorMap& tmapA, const ConvArgs& args, int cout_slice, int nslices, int num_sms,
                           cudaStream_t stream);

// Shared-memory / pipeline planning for a convolution with `nchunks` K-chunks.
int conv3x3_pick_stages(int nchunks, int cout_slice);

// Builds the activation tensor map for `conv3x3_launch`. base: NHWC 16-bit tensor [N,H,W,C].
int conv3x3_make_tmap(CUtensorMap* out, const void* base, int N, int H, int W, int C, int mode, int BW, int BN);

// Chooses the lane split for an image width.
void conv3x3_pick_tile(int W, int* BW, int* BN);

}  // namespace resr
