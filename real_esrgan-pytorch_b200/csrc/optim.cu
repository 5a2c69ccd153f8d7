// the optimizer side of the training step on the flat parameter vector.
//   torch.optim.Adam(model.parameters(), lr, betas)   train_realesrnet.py:197-200 (eps 1e-8, no weight decay, no amsgrad)
//   EMA.update()                                       model.py:42-49            shadow = (1 - d) * p + d * shadow
// One elementwise kernel over 16.7 M parameters instead of ~702 x 6 small launches per step; HBM bound:
// reads p, g, m, v, shadow and writes p, m, v, shadow = 36 B per parameter.
#include <cmath>

#include <cuda_runtime.h>

#include "../../include/resr.h"
#include "errors.h"

namespace resr {

__global__ void __launch_bounds__(256) adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                       float* __restrict__ v, float* __restrict__ shadow, size_t n, float beta1,
                                                       float beta2, float eps, float step_size, float inv_bc2_sqrt,
                                                       float ema_decay, float grad_scale) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += stride) {
        const float gi = __fmul_rn(g[i], grad_scale);
        // torch/optim/adam.py (_single_tensor_adam): exp_avg.lerp_(grad, 1 - beta1);
        // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2);
        // denom = (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps); param.addcdiv_(exp_avg, denom, value=-step_size)
        const float mi = fmaf(1.f - beta1, gi - m[i], m[i]);
        const float vi = fmaf(__fmul_rn(gi, gi), 1.f - beta2, __fmul_rn(v[i], beta2));
        const float denom = __fadd_rn(__fmul_rn(sqrtf(vi), inv_bc2_sqrt), eps);
        const float pi = fmaf(-step_size, __fdiv_rn(mi, denom), p[i]);
        m[i] = mi;
        v[i] = vi;
        p[i] = pi;
        if (shadow) shadow[i] = __fadd_rn(__fmul_rn(1.f - ema_decay, pi), __fmul_rn(ema_decay, shadow[i]));  // model.py:47
    }
}

}  // namespace resr

extern "C" int resr_adam_ema_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* ema_shadow,
                                  size_t n, float lr, float beta1, float beta2, float eps, long long step, float ema_decay,
                                  float grad_scale, void* stream) {
    using namespace resr;
    if (!params || !grads || !exp_avg || !exp_avg_sq) return set_error(RESR_E_INVALID, "null argument");
    if (step < 1) return set_error(RESR_E_INVALID, "Adam step counter starts at 1");
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
    const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
    const float step_size = static_cast<float>(static_cast<double>(lr) / bc1);
    const float inv_bc2_sqrt = static_cast<float>(1.0 / sqrt(bc2));
    int blocks = static_cast<int>((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    adam_ema_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(params, grads, exp_avg, exp_avg_sq, ema_shadow, n, beta1,
                                                                          beta2, eps, step_size, inv_bc2_sqrt, ema_decay, grad_scale);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(RESR_E_CUDA, "adam_ema_step: %s", cudaGetErrorString(e));
    return RESR_OK;
}
