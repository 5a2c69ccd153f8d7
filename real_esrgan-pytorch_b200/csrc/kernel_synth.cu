// Blur-kernel synthesis on the device in float64.
// Reference: /root/reference/imgproc.py:72-90 (_mesh_grid), :170-204 (sigma matrix, density), :225-327 (Gaussian,
// generalized Gaussian, plateau), :576-603 (sinc via Bessel J1); sequencing and zero padding to 21: dataset.py:81-141.
#include <cmath>

#include "../../include/resr.h"
#include "errors.h"

namespace resr {

// One block per kernel. type: 0 Gaussian, 1 generalized Gaussian, 2 plateau, 3 sinc, 4 delta.
__global__ void __launch_bounds__(1024) kernel_synth_kernel(const resr_kernel_params* __restrict__ prm, double* __restrict__ out64,
                                                           float* __restrict__ out32, int pad) {
    __shared__ double red[32];
    __shared__ double s_total;
    const resr_kernel_params p = prm[blockIdx.x];
    const int k = p.kernel_size;
    const int t = threadIdx.x;
    const int i = t / k, j = t % k;  // row, column
    double v = 0.0;
    if (t < k * k) {
        const double c = (k - 1) * 0.5;
        if (p.type == 3) {  // imgproc.py:590-597
            const double dx = i - c, dy = j - c;
            const double r = sqrt(dx * dx + dy * dy);
            v = (i == (k - 1) / 2 && j == (k - 1) / 2) ? p.cutoff * p.cutoff / (4.0 * M_PI)
                                                      : p.cutoff * j1(p.cutoff * r) / (2.0 * M_PI * r);
        } else if (p.type == 4) {
            v = (i == (k - 1) / 2 && j == (k - 1) / 2) ? 1.0 : 0.0;
        } else {
            // grid (imgproc.py:85-88): first coordinate = column offset, second = row offset
            const double gx = j - (k / 2), gy = i - (k / 2);
            double s00, s01, s11;
            if (p.isotropic) {  // imgproc.py:245-246
                s00 = p.sigma_x * p.sigma_x; s01 = 0.0; s11 = s00;
            } else {            // imgproc.py:182-184: U diag(sx^2, sy^2) U^T
                const double cs = cos(p.theta), sn = sin(p.theta);
                const double a = p.sigma_x * p.sigma_x, b = p.sigma_y * p.sigma_y;
                s00 = cs * cs * a + sn * sn * b;
                s01 = cs * sn * a - sn * cs * b;
                s11 = sn * sn * a + cs * cs * b;
            }
            const double det = s00 * s11 - s01 * s01;
            const double i00 = s11 / det, i01 = -s01 / det, i11 = s00 / det;
            const double q = gx * (i00 * gx + i01 * gy) + gy * (i01 * gx + i11 * gy);  // g Sigma^-1 g^T
            if (p.type == 0) v = exp(-0.5 * q);                    // imgproc.py:201
            else if (p.type == 1) v = exp(-0.5 * pow(q, p.beta));  // imgproc.py:287
            else v = 1.0 / (pow(q, p.beta) + 1.0);                 // imgproc.py:323
        }
    }
    // block sum in double
    double s = v;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((t & 31) == 0) red[t >> 5] = s;
    __syncthreads();
    if (t < 32) {
        double w = (t < (blockDim.x + 31) / 32) ? red[t] : 0.0;
        for (int o = 16; o > 0; o >>= 1) w += __shfl_down_sync(0xffffffffu, w, o);
        if (t == 0) s_total = w;
    }
    __syncthreads();
    const int P = pad > k ? pad : k;
    const size_t base = static_cast<size_t>(blockIdx.x) * P * P;
    for (int e = t; e < P * P; e += blockDim.x) {  // zero padding (dataset.py:102-103, imgproc.py:599-601)
        if (out64) out64[base + e] = 0.0;
        if (out32) out32[base + e] = 0.f;
    }
    __syncthreads();
    if (t < k * k) {
        const int off = (P - k) / 2;
        const double nv = v / s_total;
        const size_t o = base + static_cast<size_t>(i + off) * P + j + off;
        if (out64) out64[o] = nv;
        if (out32) out32[o] = static_cast<float>(nv);
    }
}

}  // namespace resr

extern "C" int resr_synthesize_kernels(const resr_kernel_params* params_host, int count, int pad, double* out_f64, float* out_f32,
                                       void* stream) {
    using namespace resr;
    if (!params_host || count <= 0 || (!out_f64 && !out_f32)) return set_error(RESR_E_INVALID, "bad argument");
    for (int i = 0; i < count; ++i) {
        const int k = params_host[i].kernel_size;
        if (k % 2 != 1 || k < 1 || k > 31) return set_error(RESR_E_INVALID, "Kernel size must be an odd number.");  // imgproc.py:350, 588
        if (pad > 0 && pad < k) return set_error(RESR_E_INVALID, "pad %d smaller than kernel size %d", pad, k);
        if (params_host[i].type < 0 || params_host[i].type > 4) return set_error(RESR_E_INVALID, "bad kernel type");
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    resr_kernel_params* dp = nullptr;
    if (cudaMallocAsync(&dp, sizeof(resr_kernel_params) * count, s) != cudaSuccess) return set_error(RESR_E_CUDA, "cudaMallocAsync failed");
    cudaMemcpyAsync(dp, params_host, sizeof(resr_kernel_params) * count, cudaMemcpyHostToDevice, s);
    int kmax = 1;
    for (int i = 0; i < count; ++i) kmax = params_host[i].kernel_size > kmax ? params_host[i].kernel_size : kmax;
    int threads = (kmax * kmax + 31) / 32 * 32;
    if (pad > 0 && pad != kmax) {
        // all kernels share one padded size
    } else if (pad <= 0) {
        for (int i = 0; i < count; ++i)
            if (params_host[i].kernel_size != kmax) { cudaFreeAsync(dp, s); return set_error(RESR_E_INVALID, "mixed sizes need pad > 0"); }
    }
    kernel_synth_kernel<<<count, threads, 0, s>>>(dp, out_f64, out_f32, pad > 0 ? pad : kmax);
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(dp, s);
    if (e != cudaSuccess) return set_error(RESR_E_CUDA, "kernel_synth: %s", cudaGetErrorString(e));
    return RESR_OK;
}
