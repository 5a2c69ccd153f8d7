// Blur-kernel synthesis on the device in float64.
// Reference: /root/reference/imgproc.py:72-90 (_mesh_grid), :170-204 (sigma matrix, density), :225-327 (Gaussian,
// generalized Gaussian, plateau), :576-603 (sinc via Bessel J1); sequencing and zero padding to 21: dataset.py:81-141.
#include <curand_kernel.h>
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

// The random decisions of dataset.py:81-141 (kernel1, kernel2, final sinc / delta of every sample) drawn on the device:
// one thread per sample, Philox4x32-10 (subsequence = sample, offset advances per call), the reference's distributions
// (uniform size choice, sinc with probability p, categorical kernel family, uniform sigma / theta, the two-sided beta
// draw of imgproc.py:405-409 / 466-470). Not the reference's RNG STREAM -- host-fed parameters remain the parity path.
__device__ __forceinline__ double draw_u(curandStatePhilox4_32_10_t* st, double lo, double hi) {
    return lo + (hi - lo) * curand_uniform_double(st);   // (0, 1]
}

__device__ void draw_mixed(curandStatePhilox4_32_10_t* st, const resr_kernel_draw_config& c, int which, int ks, resr_kernel_params* o) {
    const double* prob = which == 0 ? c.prob1 : c.prob2;
    const double* srange = which == 0 ? c.sigma_range1 : c.sigma_range2;
    const double* grange = which == 0 ? c.gen_beta_range1 : c.gen_beta_range2;
    const double* prange = which == 0 ? c.plateau_beta_range1 : c.plateau_beta_range2;
    double tot = 0.0;
    for (int i = 0; i < 6; ++i) tot += prob[i];
    double u = curand_uniform_double(st) * tot, acc = 0.0;
    int kt = 5;
    for (int i = 0; i < 6; ++i) {
        acc += prob[i];
        if (u <= acc) { kt = i; break; }
    }
    // order of config.py:24-25: isotropic, anisotropic, generalized_isotropic, generalized_anisotropic, plateau_iso, plateau_aniso
    const bool iso = (kt % 2) == 0;
    o->type = kt / 2;  // 0 Gaussian, 1 generalized, 2 plateau
    o->kernel_size = ks;
    o->isotropic = iso ? 1 : 0;
    o->reserved = 0;
    o->sigma_x = draw_u(st, srange[0], srange[1]);
    if (iso) { o->sigma_y = o->sigma_x; o->theta = 0.0; }
    else { o->sigma_y = draw_u(st, srange[0], srange[1]); o->theta = draw_u(st, -M_PI, M_PI); }
    o->beta = 1.0;
    if (o->type != 0) {
        const double* br = o->type == 1 ? grange : prange;
        o->beta = curand_uniform_double(st) < 0.5 ? draw_u(st, br[0], 1.0) : draw_u(st, 1.0, br[1]);
    }
    o->cutoff = 1.0;
}

__global__ void draw_kernel_params_kernel(const resr_kernel_draw_config cfg, int batch, unsigned long long seed,
                                          unsigned long long* __restrict__ call_state, resr_kernel_params* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long call = call_state ? *call_state : 0ull;
    if (b < batch) {
        curandStatePhilox4_32_10_t st;
        curand_init(seed, static_cast<unsigned long long>(b), call * 64ull, &st);
        for (int which = 0; which < 2; ++which) {
            resr_kernel_params* o = out + 3 * b + which;
            const int ks = cfg.sizes[min(static_cast<int>(curand_uniform_double(&st) * cfg.n_sizes), cfg.n_sizes - 1)];
            if (curand_uniform_double(&st) < (which == 0 ? cfg.sinc_prob1 : cfg.sinc_prob2)) {   // dataset.py:84-91, 108-115
                o->type = 3; o->kernel_size = ks; o->isotropic = 1; o->reserved = 0;
                o->sigma_x = o->sigma_y = 1.0; o->theta = 0.0; o->beta = 1.0;
                o->cutoff = ks < cfg.sinc_size_split ? draw_u(&st, M_PI / 3, M_PI) : draw_u(&st, M_PI / 5, M_PI);
            } else {
                draw_mixed(&st, cfg, which, ks, o);
            }
        }
        resr_kernel_params* o = out + 3 * b + 2;   // dataset.py:131-139
        o->isotropic = 1; o->reserved = 0; o->sigma_x = o->sigma_y = 1.0; o->theta = 0.0; o->beta = 1.0; o->cutoff = 1.0;
        if (curand_uniform_double(&st) < cfg.sinc_prob3) {
            o->type = 3;
            o->kernel_size = cfg.sizes[min(static_cast<int>(curand_uniform_double(&st) * cfg.n_sizes), cfg.n_sizes - 1)];
            o->cutoff = draw_u(&st, M_PI / 3, M_PI);
        } else {
            o->type = 4;
            o->kernel_size = cfg.final_size;
        }
    }
}

__global__ void bump_counter_kernel(unsigned long long* c) { *c += 1ull; }  // stream-ordered after every reader of the old value

}  // namespace resr

extern "C" int resr_draw_degradation_kernel_params(const resr_kernel_draw_config* cfg, int batch, unsigned long long seed,
                                                   unsigned long long* call_state, resr_kernel_params* params_dev, void* stream) {
    using namespace resr;
    if (!cfg || !params_dev || batch <= 0) return set_error(RESR_E_INVALID, "bad argument");
    if (cfg->n_sizes < 1 || cfg->n_sizes > 16) return set_error(RESR_E_INVALID, "n_sizes must be 1..16");
    for (int i = 0; i < cfg->n_sizes; ++i)
        if (cfg->sizes[i] % 2 != 1 || cfg->sizes[i] < 1 || cfg->sizes[i] > 31) return set_error(RESR_E_INVALID, "Kernel size must be an odd number.");
    draw_kernel_params_kernel<<<(batch + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(*cfg, batch, seed, call_state, params_dev);
    if (call_state) bump_counter_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(call_state);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(RESR_E_CUDA, "draw_kernel_params: %s", cudaGetErrorString(e));
    return RESR_OK;
}

extern "C" int resr_synthesize_kernels_device(const resr_kernel_params* params_dev, int count, int kmax, int pad, double* out_f64,
                                              float* out_f32, void* stream) {
    using namespace resr;
    if (!params_dev || count <= 0 || (!out_f64 && !out_f32)) return set_error(RESR_E_INVALID, "bad argument");
    if (kmax % 2 != 1 || kmax < 1 || kmax > 31 || pad < kmax) return set_error(RESR_E_INVALID, "kmax must be odd, <= 31 and <= pad");
    const int threads = (kmax * kmax + 31) / 32 * 32;
    kernel_synth_kernel<<<count, threads, 0, static_cast<cudaStream_t>(stream)>>>(params_dev, out_f64, out_f32, pad);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(RESR_E_CUDA, "kernel_synth: %s", cudaGetErrorString(e));
    return RESR_OK;
}

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
