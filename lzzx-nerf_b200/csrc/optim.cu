// optim.cu — AdamW over ONE flat parameter buffer: the reference steps ~70 tensors through torch's AdamW after a GradScaler
// unscale pass (TrainerUtil.py:1040-1056, network.py:315-357: tables lr 1e-2 / default weight decay, networks lr 1e-3 / no decay,
// betas (0, 0.99), eps 1e-8 — train.py:274).  Here parameters, gradients and both moments each live in one contiguous fp32 buffer (the
// gradient buffer is the one the data-parallel all-reduce uses), split into contiguous hyper-parameter groups (network.py:332-356: tables lr / AdamW's
// default decay 0.01, networks lr_net / wd, audio_att_net 5 lr_net / 1e-4); one launch does unscale + skip-on-overflow + decoupled weight decay + moment
// update + parameter update and, on the steps the trainer asks for it (TrainerUtil.py:1055-1056, every 1000 steps), the weight EMA of torch_ema.  The step counter, the loss scale and the overflow flag
// stay on the device, so nothing in the optimizer step synchronises or blocks CUDA-graph capture.
#include "common.cuh"

namespace b2n {

__global__ void k_adamw_step_count(float *step, const float *found_inf) {
    if (found_inf == nullptr || found_inf[0] == 0.0f) step[0] += 1.0f;
}

__global__ void __launch_bounds__(256) k_adamw_flat(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, uint32_t n,
                                                    const __grid_constant__ b2n_adam_groups grp, float beta1, float beta2, float eps,
                                                    const float *__restrict__ step, const float *__restrict__ grad_scale, const float *__restrict__ found_inf,
                                                    float *__restrict__ ema, float ema_decay) {
    const float one_minus_decay = 1.0f - ema_decay;
    if (found_inf != nullptr && found_inf[0] != 0.0f) {                        // overflow in this step's gradients: skip (GradScaler semantics) ...
        if (ema != nullptr)                                                    // ... but the reference's ema.update() still runs on the unchanged weights
            for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { const float s = ema[i]; ema[i] = s - one_minus_decay * (s - p[i]); }
        return;
    }
    const float t = step[0];                                                   // already incremented by k_adamw_step_count
    const float inv_scale = grad_scale != nullptr ? 1.0f / grad_scale[0] : 1.0f;
    const float bias1 = 1.0f - powf(beta1, t), bias2 = 1.0f - powf(beta2, t);
    const float rsqrt_bias2 = 1.0f / sqrtf(bias2);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t k = 0;
        while (k + 1 < grp.n_groups && i >= grp.end[k]) k++;                  // <= 8 contiguous hyper-parameter groups
        const float lr = grp.lr[k], wd = grp.weight_decay[k];
        const float grad = g[i] * inv_scale;
        float w = p[i];
        w -= lr * wd * w;                                                      // decoupled weight decay
        const float mi = beta1 * m[i] + (1.0f - beta1) * grad;
        const float vi = beta2 * v[i] + (1.0f - beta2) * grad * grad;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) * rsqrt_bias2 + eps;
        w -= (lr / bias1) * (mi / denom);
        p[i] = w;
        if (ema != nullptr) { const float s = ema[i]; ema[i] = s - one_minus_decay * (s - w); }     // torch_ema: shadow -= (1 - decay) (shadow - param)
    }
}

// GradScaler on the device (torch.amp.GradScaler's _amp_foreach_non_finite_check_and_unscale_ + _amp_update_scale_, TrainerUtil.py:1046-1047): the non-finite check
// is ONE pass over the flat gradient buffer instead of a multi-tensor launch over ~60 views, the unscale is folded into the AdamW kernel, and the scale update
// is a one-thread kernel behind it.  scaler_state: float[4] = {scale, growth tracker, found_inf, -}.
__global__ void __launch_bounds__(256) k_flat_nonfinite(const float *__restrict__ g, uint32_t n, float *__restrict__ found_inf) {
    bool bad = false;
    const uint32_t n4 = n / 4;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(g) + i);
        bad |= !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w));
    }
    if (blockIdx.x == 0 && threadIdx.x < n - 4 * n4) bad |= !isfinite(g[4 * n4 + threadIdx.x]);
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) found_inf[0] = 1.0f;
}
__global__ void k_scaler_update(float *state, float growth, float backoff, float interval) {
    if (state[2] != 0.0f) { state[0] *= backoff; state[1] = 0.0f; }
    else {
        const float t = state[1] + 1.0f;
        if (t >= interval) { const float grown = state[0] * growth; if (isfinite(grown)) state[0] = grown; state[1] = 0.0f; }
        else state[1] = t;
    }
}

}  // namespace b2n

using namespace b2n;

extern "C" int b2n_adamw_flat_groups(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, uint32_t n, const b2n_adam_groups *groups, float beta1,
                                     float beta2, float eps, float *step, const float *grad_scale, const float *found_inf, float *ema, float ema_decay, void *stream) {
    B2N_REQUIRE(params && grads && exp_avg && exp_avg_sq && step && groups, "adamw_flat: null pointer");
    B2N_REQUIRE(groups->n_groups >= 1 && groups->n_groups <= 8, "adamw_flat: %u groups (1..8)", groups->n_groups);
    for (uint32_t k = 0; k < groups->n_groups; k++)
        B2N_REQUIRE(groups->end[k] <= n && (k == 0 || groups->end[k] >= groups->end[k - 1]), "adamw_flat: group %u ends at %u (n = %u; ends must ascend)", k, groups->end[k], n);
    B2N_REQUIRE(groups->end[groups->n_groups - 1] == n, "adamw_flat: the last group must end at n");
    if (n == 0) return 0;
    cudaStream_t st = as_stream(stream);
    k_adamw_step_count<<<1, 1, 0, st>>>(step, found_inf);
    if (check_launch("adamw_flat(step)")) return 1;
    uint32_t blocks = ceil_div<uint32_t>(n, 256);
    const uint32_t cap = (uint32_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    k_adamw_flat<<<blocks, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, *groups, beta1, beta2, eps, step, grad_scale, found_inf, ema, ema_decay);
    return check_launch("adamw_flat");
}

extern "C" int b2n_adamw_flat(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, uint32_t n, uint32_t n_group0, float lr0,
                              float weight_decay0, float lr1, float weight_decay1, float beta1, float beta2, float eps, float *step,
                              const float *grad_scale, const float *found_inf, void *stream) {
    B2N_REQUIRE(n_group0 <= n, "adamw_flat: group boundary %u beyond n=%u", n_group0, n);
    b2n_adam_groups g = {};
    g.n_groups = 2; g.end[0] = n_group0; g.end[1] = n; g.lr[0] = lr0; g.lr[1] = lr1; g.weight_decay[0] = weight_decay0; g.weight_decay[1] = weight_decay1;
    return b2n_adamw_flat_groups(params, grads, exp_avg, exp_avg_sq, n, &g, beta1, beta2, eps, step, grad_scale, found_inf, nullptr, 0.0f, stream);
}

// AdamW + the whole GradScaler protocol in four small launches: non-finite check of the (all-reduced) flat gradients -> step counter -> AdamW with the unscale and
// the overflow skip folded in -> scale update (x backoff on overflow, x growth after `growth_interval` clean steps).  scaler_state: device float[4]
// {scale, growth tracker, found_inf (written here), unused}.  `grads` must be 16-byte aligned.
extern "C" int b2n_adamw_flat_scaled(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, uint32_t n, const b2n_adam_groups *groups, float beta1,
                                     float beta2, float eps, float *step, float *scaler_state, float growth_factor, float backoff_factor, uint32_t growth_interval,
                                     float *ema, float ema_decay, void *stream) {
    B2N_REQUIRE(grads && scaler_state, "adamw_flat_scaled: null pointer");
    B2N_REQUIRE(((uintptr_t)grads & 15) == 0, "adamw_flat_scaled: the gradient buffer must be 16-byte aligned");
    if (n == 0) return 0;
    cudaStream_t st = as_stream(stream);
    B2N_CUDA(cudaMemsetAsync(scaler_state + 2, 0, sizeof(float), st));
    uint32_t blocks = ceil_div<uint32_t>(n / 4 + 1, 256);
    const uint32_t cap = (uint32_t)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    k_flat_nonfinite<<<blocks, 256, 0, st>>>(grads, n, scaler_state + 2);
    if (check_launch("adamw_flat_scaled(check)")) return 1;
    if (int rc = b2n_adamw_flat_groups(params, grads, exp_avg, exp_avg_sq, n, groups, beta1, beta2, eps, step, scaler_state, scaler_state + 2, ema, ema_decay, stream)) return rc;
    k_scaler_update<<<1, 1, 0, st>>>(scaler_state, growth_factor, backoff_factor, (float)growth_interval);
    return check_launch("adamw_flat_scaled(update)");
}
