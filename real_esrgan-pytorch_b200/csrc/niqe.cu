// NIQE feature extraction on the device (reference /root/reference/image_quality_assessment.py:886-998
// `_niqe_torch` and the helpers it calls, imgproc.py:1815-1840 for the Y channel). Everything after the Y channel is
// float64, as in the reference. Stages for each of the two scales (block = 96, then 48 on the half-size image):
//   Y * 255 rounded (fp32 arithmetic like the reference, then float64)
//   MSCN coefficients: 7 x 7 Gaussian (sigma 7/6, replicate padding) local mean / deviation, (y - mu) / (sigma + 1)
//   per block and per map (the block itself + its products with four circularly shifted copies, torch.roll): six sums
//   AGGD fit per (block, map): shape parameter by table search over 0.2 : 0.001 : 10, left / right scale, mean -> 18 features
//   MATLAB-style antialiased bicubic x0.5 (ten taps, one weight set, symmetric padding) for the second scale
// The 36-dimensional Gaussian fit over the blocks (nanmean, covariance, pseudo-inverse: a 36 x 36 problem) is left to the
// caller (resr_b200/iqa.py does it with torch.linalg on the device).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

#include "../../include/resr.h"
#include "device_state.h"
#include "errors.h"

namespace resr {

static constexpr int kGamN = 9801;   // torch.arange(0.2, 10.001, 0.001)
__constant__ double c_niqe_win[49];
__constant__ double c_niqe_rw[10];   // resize weights

// Y channel of rgb2ycbcr_torch(only_use_y_channel=True) in fp32, * 255, round (half to even), as float64; the crop of
// `crop_border` pixels and the crop to whole blocks are folded into the indexing.
__global__ void __launch_bounds__(256) niqe_y_kernel(const float* __restrict__ x, double* __restrict__ y, int B, int H, int W, int border,
                                                     int h, int w) {
    const size_t total = static_cast<size_t>(B) * h * w;
    const size_t plane = static_cast<size_t>(H) * W;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int xx = static_cast<int>(i % w), yy = static_cast<int>((i / w) % h), b = static_cast<int>(i / (static_cast<size_t>(w) * h));
        const float* p = x + static_cast<size_t>(b) * 3 * plane + static_cast<size_t>(yy + border) * W + xx + border;
        float v = __fadd_rn(__fadd_rn(__fmul_rn(p[0], 65.481f), __fmul_rn(p[plane], 128.553f)), __fmul_rn(p[2 * plane], 24.966f));
        v = __fdiv_rn(__fadd_rn(v, 16.0f), 255.0f);            // imgproc.py:1830, 1838
        y[i] = static_cast<double>(rintf(__fmul_rn(v, 255.0f)));   // image_quality_assessment.py:984-985
    }
}

// structdis = (y - mu) / (sqrt(|E[y^2] - mu^2| + 1e-8) + 1), 7 x 7 window, replicate padding (image_quality_assessment.py:869-873)
__global__ void __launch_bounds__(256) niqe_mscn_kernel(const double* __restrict__ y, double* __restrict__ sd, int B, int h, int w) {
    const size_t total = static_cast<size_t>(B) * h * w;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int xx = static_cast<int>(i % w), yy = static_cast<int>((i / w) % h);
        const double* img = y + (i - static_cast<size_t>(yy) * w - xx);
        double mu = 0.0, e2 = 0.0;
        for (int dy = 0; dy < 7; ++dy) {
            const int sy = min(max(yy + dy - 3, 0), h - 1);
            for (int dx = 0; dx < 7; ++dx) {
                const int sx = min(max(xx + dx - 3, 0), w - 1);
                const double v = img[static_cast<size_t>(sy) * w + sx], k = c_niqe_win[dy * 7 + dx];
                mu += k * v;
                e2 += k * (v * v);
            }
        }
        const double sigma = sqrt(fabs(e2 - mu * mu) + 1e-8);
        sd[i] = (img[static_cast<size_t>(yy) * w + xx] - mu) / (sigma + 1.0);
    }
}

// One CUDA block per image block of s x s MSCN coefficients. Map m = 0: the block; m = 1..4: block * roll(block, shift_m)
// with shifts (0,1), (1,0), (1,1), (1,-1) (torch.roll: circular inside the block, image_quality_assessment.py:851-853).
// stats[(block * 5 + m) * 6 + {count<0, count>0, sum v^2 over v<0, sum v^2 over v>0, sum |v|, sum v^2}].
__global__ void __launch_bounds__(256) niqe_block_stats_kernel(const double* __restrict__ sd, double* __restrict__ stats, int h, int w, int s,
                                                               int nbh, int nbw) {
    __shared__ double red[8][30];
    const int blk = blockIdx.x;                       // (image, block row, block column)
    const int bw_i = blk % nbw, bh_i = (blk / nbw) % nbh, img = blk / (nbw * nbh);
    const double* base = sd + (static_cast<size_t>(img) * h + static_cast<size_t>(bh_i) * s) * w + static_cast<size_t>(bw_i) * s;
    double acc[30];
#pragma unroll
    for (int i = 0; i < 30; ++i) acc[i] = 0.0;
    const int sy[5] = {0, 0, 1, 1, 1}, sx[5] = {0, 1, 0, 1, -1};
    for (int p = threadIdx.x; p < s * s; p += blockDim.x) {
        const int i = p / s, j = p % s;
        const double v0 = base[static_cast<size_t>(i) * w + j];
#pragma unroll
        for (int m = 0; m < 5; ++m) {
            double v = v0;
            if (m > 0) {
                const int ii = (i - sy[m] + s) % s, jj = (j - sx[m] + s) % s;
                v = v0 * base[static_cast<size_t>(ii) * w + jj];
            }
            const double v2 = v * v;
            if (v < 0) { acc[m * 6 + 0] += 1.0; acc[m * 6 + 2] += v2; }
            if (v > 0) { acc[m * 6 + 1] += 1.0; acc[m * 6 + 3] += v2; }
            acc[m * 6 + 4] += fabs(v);
            acc[m * 6 + 5] += v2;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 30; ++i) {
        double v = acc[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 30) {
        double v = 0.0;
        for (int k = 0; k < 8; ++k) v += red[k][threadIdx.x];
        stats[static_cast<size_t>(blk) * 30 + threadIdx.x] = v;
    }
}

// AGGD fit of one (block, map) (image_quality_assessment.py:790-836 with get_sigma=True) and its slot of the 18 features
// (:845-858): one CUDA block; the table search is a parallel arg-min (first minimum wins, like torch.argmin).
__global__ void __launch_bounds__(256) niqe_aggd_kernel(const double* __restrict__ stats, const double* __restrict__ r_gam,
                                                        double* __restrict__ feat, int s, int col0) {
    __shared__ double s_val[256];
    __shared__ int s_idx[256];
    const int bm = blockIdx.x, blk = bm / 5, m = bm % 5;
    const double* st = stats + static_cast<size_t>(bm) * 6;
    const double n = static_cast<double>(s) * s;
    const double left_std = sqrt(st[2] / (st[0] + 1e-8)), right_std = sqrt(st[3] / (st[1] + 1e-8));
    const double gamma_hat = left_std / right_std;
    const double mean_abs = st[4] / n;
    const double rhat = mean_abs * mean_abs / (st[5] / n);
    const double g2 = gamma_hat * gamma_hat;
    const double rhat_norm = (rhat * (g2 * gamma_hat + 1.0) * (gamma_hat + 1.0)) / ((g2 + 1.0) * (g2 + 1.0));
    double best = INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < kGamN; i += blockDim.x) {
        const double d = fabs(r_gam[i] - rhat_norm);
        if (d < best) { best = d; bi = i; }   // i increases: the first minimum of this thread's entries is kept
    }
    s_val[threadIdx.x] = best; s_idx[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (static_cast<int>(threadIdx.x) < o) {
            const double v = s_val[threadIdx.x + o];
            const int ix = s_idx[threadIdx.x + o];
            if (v < s_val[threadIdx.x] || (v == s_val[threadIdx.x] && ix < s_idx[threadIdx.x])) { s_val[threadIdx.x] = v; s_idx[threadIdx.x] = ix; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // a NaN distance (a constant block: 0 / 0) loses every comparison: torch.argmin returns the first NaN's index (0)
        const int pos = s_idx[0] == 0x7fffffff ? 0 : s_idx[0];
        const double alpha = 0.2 + 0.001 * pos;
        const double sc = sqrt(exp(lgamma(1.0 / alpha) - lgamma(3.0 / alpha)));
        const double lb = left_std * sc, rb = right_std * sc;
        double* f = feat + static_cast<size_t>(blk) * 36 + col0;
        if (m == 0) {
            f[0] = alpha; f[1] = (lb + rb) / 2.0;
        } else {
            const double mean = (rb - lb) * exp(lgamma(2.0 / alpha) - lgamma(1.0 / alpha));
            double* q = f + 2 + (m - 1) * 4;
            q[0] = alpha; q[1] = mean; q[2] = lb; q[3] = rb;
        }
    }
}

// MATLAB imresize(x / 255, 0.5) * 255 with antialiasing along one axis (image_quality_assessment.py:517-585 for scale 0.5:
// ten taps, pos = 2 i + 0.5, base = 2 i - 4, one weight set, symmetric padding of four: the edge element is used twice).
// axis 0: rows (h -> h / 2), axis 1: columns. div_in / scale_out fold the / 255 and * 255 of :877-878 into the passes.
__global__ void __launch_bounds__(256) niqe_resize_half_kernel(const double* __restrict__ in, double* __restrict__ out, int B, int h, int w,
                                                               int axis, double div_in, double scale_out) {
    const int ho = axis == 0 ? h / 2 : h, wo = axis == 1 ? w / 2 : w;
    const size_t total = static_cast<size_t>(B) * ho * wo;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int xx = static_cast<int>(i % wo), yy = static_cast<int>((i / wo) % ho), b = static_cast<int>(i / (static_cast<size_t>(wo) * ho));
        const double* img = in + static_cast<size_t>(b) * h * w;
        const int n = axis == 0 ? h : w, o = axis == 0 ? yy : xx;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            int idx = 2 * o - 4 + k;
            idx = idx < 0 ? -idx - 1 : (idx >= n ? 2 * n - idx - 1 : idx);
            const double v = axis == 0 ? img[static_cast<size_t>(idx) * w + xx] : img[static_cast<size_t>(yy) * w + idx];
            acc += (v / div_in) * c_niqe_rw[k];
        }
        out[i] = acc * scale_out;
    }
}

static int grid_of(size_t total) {
    size_t g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    return static_cast<int>(g < 1 ? 1 : g);
}

static double cubic_w(double x) {   // image_quality_assessment.py:388-404, a = -0.5
    const double a = -0.5, ax = fabs(x), ax2 = ax * ax, ax3 = ax * ax2;
    if (ax <= 1) return (a + 2) * ax3 - (a + 3) * ax2 + 1;
    if (ax <= 2) return a * ax3 - 5 * a * ax2 + 8 * a * ax - 4 * a;
    return 0.0;
}

struct NiqeTables { bool ready = false; double* r_gam = nullptr; };

static int niqe_tables(double** r_gam_out, cudaStream_t s) {
    static PerDevice<NiqeTables> tabs;
    NiqeTables& t = tabs.cur();
    if (!t.ready) {
        // Gaussian window: fspecial('gaussian', 7, 7/6) in double, kept as float32 by the reference (:215-240)
        double win[49], sum = 0.0, mx = 0.0;
        const double sigma = 7.0 / 6.0;
        for (int i = 0; i < 49; ++i) {
            const double yy = i / 7 - 3.0, xx = i % 7 - 3.0;
            win[i] = exp(-(xx * xx + yy * yy) / (2.0 * sigma * sigma));
            mx = fmax(mx, win[i]);
        }
        for (int i = 0; i < 49; ++i) { if (win[i] < 2.220446049250313e-16 * mx) win[i] = 0.0; sum += win[i]; }
        for (int i = 0; i < 49; ++i) win[i] = static_cast<double>(static_cast<float>(win[i] / sum));
        double rw[10], rs = 0.0;
        for (int k = 0; k < 10; ++k) { rw[k] = cubic_w((4.5 - k) * 0.5); rs += rw[k]; }
        for (int k = 0; k < 10; ++k) rw[k] /= rs;
        double* host = new double[kGamN];
        for (int i = 0; i < kGamN; ++i) {
            const double a = 0.2 + 0.001 * i;
            host[i] = exp(2.0 * lgamma(2.0 / a) - (lgamma(1.0 / a) + lgamma(3.0 / a)));
        }
        cudaError_t e = cudaMemcpyToSymbol(c_niqe_win, win, sizeof(win));
        if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_niqe_rw, rw, sizeof(rw));
        if (e == cudaSuccess) e = cudaMalloc(&t.r_gam, kGamN * sizeof(double));
        if (e == cudaSuccess) e = cudaMemcpyAsync(t.r_gam, host, kGamN * sizeof(double), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        delete[] host;
        if (e != cudaSuccess) return set_error(RESR_E_CUDA, "NIQE tables: %s", cudaGetErrorString(e));
        t.ready = true;
    }
    *r_gam_out = t.r_gam;
    return RESR_OK;
}

}  // namespace resr

using namespace resr;

extern "C" {

int resr_niqe_num_blocks(int h, int w, int crop_border, int block) {
    if (block <= 0 || (block & 1) || crop_border < 0) return 0;
    const int hh = h - 2 * crop_border, ww = w - 2 * crop_border;
    if (hh < block || ww < block) return 0;
    return (hh / block) * (ww / block);
}

size_t resr_niqe_workspace_bytes(int b, int h, int w, int crop_border, int block) {
    const int nb = resr_niqe_num_blocks(h, w, crop_border, block);
    if (b <= 0 || nb == 0) return 0;
    const int hh = (h - 2 * crop_border) / block * block, ww = (w - 2 * crop_border) / block * block;
    const size_t img = static_cast<size_t>(b) * hh * ww * sizeof(double);
    return 3 * img + static_cast<size_t>(b) * nb * 30 * sizeof(double) + 1024;
}

int resr_niqe_features(const float* image_rgb, double* features, int b, int h, int w, int crop_border, int block, void* workspace,
                       size_t workspace_bytes, void* stream) {
    if (!image_rgb || !features || !workspace) return set_error(RESR_E_INVALID, "null argument");
    const int nb = resr_niqe_num_blocks(h, w, crop_border, block);
    if (b <= 0 || nb == 0) return set_error(RESR_E_INVALID, "image smaller than one %d x %d block after cropping (or odd block size)", block, block);
    if (workspace_bytes < resr_niqe_workspace_bytes(b, h, w, crop_border, block)) return set_error(RESR_E_NOMEM, "NIQE workspace too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double* r_gam = nullptr;
    const int rc = niqe_tables(&r_gam, s);
    if (rc != RESR_OK) return rc;
    const int nbh = (h - 2 * crop_border) / block, nbw = (w - 2 * crop_border) / block;
    int hh = nbh * block, ww = nbw * block;
    const size_t img = static_cast<size_t>(b) * hh * ww;
    double* y = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
    double* sd = y + img;
    double* tmp = sd + img;
    double* stats = tmp + img;
    niqe_y_kernel<<<grid_of(img), 256, 0, s>>>(image_rgb, y, b, h, w, crop_border, hh, ww);
    int bs = block;
    for (int scale = 1; scale <= 2; ++scale) {
        const size_t cur = static_cast<size_t>(b) * hh * ww;
        niqe_mscn_kernel<<<grid_of(cur), 256, 0, s>>>(y, sd, b, hh, ww);
        niqe_block_stats_kernel<<<b * nbh * nbw, 256, 0, s>>>(sd, stats, hh, ww, bs, nbh, nbw);
        niqe_aggd_kernel<<<b * nbh * nbw * 5, 256, 0, s>>>(stats, r_gam, features, bs, (scale - 1) * 18);
        if (scale == 1) {   // y = imresize(y / 255, 0.5) * 255: rows first, then columns (:877-878, :517-585)
            niqe_resize_half_kernel<<<grid_of(cur / 2), 256, 0, s>>>(y, tmp, b, hh, ww, 0, 255.0, 1.0);
            niqe_resize_half_kernel<<<grid_of(cur / 4), 256, 0, s>>>(tmp, y, b, hh / 2, ww, 1, 1.0, 255.0);
            hh /= 2; ww /= 2; bs /= 2;
        }
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(RESR_E_CUDA, "NIQE launch: %s", cudaGetErrorString(e));
    return RESR_OK;
}

}  // extern "C"
