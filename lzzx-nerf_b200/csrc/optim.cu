// optim.cu — AdamW over ONE flat parameter buffer: the reference steps ~70 tensors through torch's AdamW after a GradScaler
// unscale pass (TrainerUtil.py:1040-1056, network.py:315-357: tables lr 1e-2 / default weight decay, networks lr 1e-3 / no decay,
// betas (0, 0.99), eps 1e-8 — train.py:274).  Here parameters, gradients and both moments each live in one contiguous fp32 buffer (the
// gradient buffer is the one the data-parallel all-reduce uses), split in two hyper-parameter groups at `n0`; one launch does unscale +
// skip-on-overflow + decoupled weight decay + moment update + parameter update.  The step counter, the loss scale and the overflow flag
// stay on the device, so nothing in the optimizer step synchronises or blocks CUDA-graph capture.
#include "common.cuh"

namespace b2n {

struct AdamGroup { float lr, weight_decay; };

__global__ void k_adamw_step_count(float *step, const float *found_inf) {
    if (found_inf == nullptr || found_inf[0] == 0.0f) step[0] += 1.0f;
}

__global__ void __launch_bounds__(256) k_adamw_flat(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, uint32_t n,
                                                    uint32_t n0, AdamGroup g0, AdamGroup g1, float beta1, float beta2, float eps,
                                                    const float *__restrict__ step, const float *__restrict__ grad_scale, const float *__restrict__ found_inf) {
    if (found_inf != nullptr && found_inf[0] != 0.0f) return;                 // overflow in this step's gradients: skip (GradScaler semantics)
    const float t = step[0];                                                   // already incremented by k_adamw_step_count
    const float inv_scale = grad_scale != nullptr ? 1.0f / grad_scale[0] : 1.0f;
    const float bias1 = 1.0f - powf(beta1, t), bias2 = 1.0f - powf(beta2, t);
    const float rsqrt_bias2 = 1.0f / sqrtf(bias2);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const AdamGroup h = i < n0 ? g0 : g1;
        const float grad = g[i] * inv_scale;
        float w = p[i];
        w -= h.lr * h.weight_decay * w;                                        // decoupled weight decay
        const float mi = beta1 * m[i] + (1.0f - beta1) * grad;
        const float vi = beta2 * v[i] + (1.0f - beta2) * grad * grad;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) * rsqrt_bias2 + eps;
        p[i] = w - (h.lr / bias1) * (mi / denom);
    }
}

}  // namespace b2n

using namespace b2n;

extern "C" int b2n_adamw_flat(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, uint32_t n, uint32_t n_group0, float lr0,
                              float weight_decay0, float lr1, float weight_decay1, float beta1, float beta2, float eps, float *step,
                              const float *grad_scale, const float *found_inf, void *stream) {
    B2N_REQUIRE(params && grads && exp_avg && exp_avg_sq && step, "adamw_flat: null pointer");
    B2N_REQUIRE(n_group0 <= n, "adamw_flat: group boundary %u beyond n=%u", n_group0, n);
    if (n == 0) return 0;
    cudaStream_t st = as_stream(stream);
    k_adamw_step_count<<<1, 1, 0, st>>>(step, found_inf);
    if (check_launch("adamw_flat(step)")) return 1;
    uint32_t blocks = ceil_div<uint32_t>(n, 256);
    const uint32_t cap = (uint32_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    k_adamw_flat<<<blocks, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, n_group0, AdamGroup{lr0, weight_decay0}, AdamGroup{lr1, weight_decay1}, beta1,
                                         beta2, eps, step, grad_scale, found_inf);
    return check_launch("adamw_flat");
}
