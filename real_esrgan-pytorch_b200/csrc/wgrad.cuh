// Internal interface of the tensor-core weight-gradient kernel (wgrad_tc.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace resr {

struct WgradArgs {
    int N, H, W;
    int cin, cout;
    int n_mb;             // input-channel blocks of 128
    int n_cs;             // output-channel slices of 32
    int segs_per_row;     // ceil(W / 64)
    long long kstages_total;  // N * H * segs_per_row (one stage = 64 pixels of one row)
    float* partial;       // [nsplit][n_mb][n_cs][3 dy][128 ci][96 = (dx, co)] fp32
    int nstages;
    int dy_rows;          // channel rows of one shifted copy of dY^T
};

size_t wgrad_partial_bytes(int num_sms);

// xt: channels-first bf16 activations [x_channels][N][H][W]; dyt: three x-shifted channels-first bf16 copies of the
// output gradient [3][dy_channels][N][H][W] (dy_channels >= ceil(cout/32)*32; copy dx holds dY[.., x - dx + 1]). dw: OIHW fp32 [cout][cin][3][3]; db: [cout] or null.
int wgrad_launch(const uint16_t* xt, int x_channels, const uint16_t* dyt, int dy_channels, int N, int H, int W, int cin, int cout,
                 float* partial, float* dw, float* db, int num_sms, cudaStream_t s);

}  // namespace resr
