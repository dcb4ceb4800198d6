// loss.cu — the head branch's training loss and its gradient in two launches (TrainerUtil.py:238-300; renderer.py:559-561 for the background blend):
//     img  = clamp(image + (1 - weights_sum) * bg, 0, 1)
//     loss = mean_n mean_c (img - gt)^2  +  lambda_ent * mean_n H2(clamp(weights_sum, 1e-5, 1 - 1e-5))  +  lambda_amb * (mean_n aud_sum + mean_n eye_sum)
// with H2(a) = -a log2 a - (1 - a) log2 (1 - a).  The reference builds this from ~40 elementwise / reduction kernels over 65 536 rays and autograd
// replays as many backwards; both directions are one pass over the rays here.
#include "common.cuh"

namespace b2n {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256) k_head_loss_fwd(const float *__restrict__ image, const float *__restrict__ ws, const float *__restrict__ aud, const float *__restrict__ eye,
                                                        const float *__restrict__ gt, const float *__restrict__ bg, uint32_t bg_per_ray, uint32_t N, float lambda_ent,
                                                        float lambda_amb, float *__restrict__ loss) {
    float acc = 0.0f;
    const float inv_n = 1.0f / (float)N;
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const float w = ws[n], k = 1.0f - w;
        float se = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float b = bg[bg_per_ray ? 3 * (size_t)n + c : c];
            const float v = fminf(fmaxf(image[3 * (size_t)n + c] + k * b, 0.0f), 1.0f);
            const float d = v - gt[3 * (size_t)n + c];
            se += d * d;
        }
        const float a = fminf(fmaxf(w, 1e-5f), 1.0f - 1e-5f);
        const float ent = -a * log2f(a) - (1.0f - a) * log2f(1.0f - a);
        acc += se * (1.0f / 3.0f) + lambda_ent * ent + lambda_amb * (aud[n] + eye[n]);
    }
    __shared__ float s[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < 8 ? s[threadIdx.x] : 0.0f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(loss, v * inv_n);
    }
}

// gradients of the loss w.r.t. the composite's outputs, times the upstream scalar *g (the GradScaler's scale)
__global__ void __launch_bounds__(256) k_head_loss_bwd(const float *__restrict__ image, const float *__restrict__ ws, const float *__restrict__ gt, const float *__restrict__ bg,
                                                        uint32_t bg_per_ray, uint32_t N, float lambda_ent, float lambda_amb, const float *__restrict__ g,
                                                        float *__restrict__ d_image, float *__restrict__ d_ws, float *__restrict__ d_aud, float *__restrict__ d_eye) {
    const float up = g[0], inv_n = 1.0f / (float)N;
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const float w = ws[n], k = 1.0f - w;
        float dw = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float b = bg[bg_per_ray ? 3 * (size_t)n + c : c];
            const float pre = image[3 * (size_t)n + c] + k * b;
            const float v = fminf(fmaxf(pre, 0.0f), 1.0f);
            // clamp passes the gradient where min <= x <= max (torch.clamp backward)
            const float dv = (pre >= 0.0f && pre <= 1.0f) ? up * (2.0f / 3.0f) * inv_n * (v - gt[3 * (size_t)n + c]) : 0.0f;
            d_image[3 * (size_t)n + c] = dv;
            dw -= dv * b;
        }
        if (w >= 1e-5f && w <= 1.0f - 1e-5f) dw += up * lambda_ent * inv_n * (log2f(1.0f - w) - log2f(w));       // d H2 / d a = log2((1 - a) / a)
        d_ws[n] = dw;
        d_aud[n] = up * lambda_amb * inv_n;
        d_eye[n] = up * lambda_amb * inv_n;
    }
}

}  // namespace b2n

using namespace b2n;

extern "C" int b2n_head_loss_forward(const float *image, const float *weights_sum, const float *aud_sum, const float *eye_sum, const float *gt_rgb,
                                     const float *bg_color, int bg_per_ray, uint32_t N, float lambda_ent, float lambda_amb, float *loss, void *stream) {
    B2N_REQUIRE(image && weights_sum && aud_sum && eye_sum && gt_rgb && bg_color && loss, "head_loss_forward: null pointer");
    B2N_REQUIRE(N > 0, "head_loss_forward: empty batch");
    cudaStream_t st = as_stream(stream);
    B2N_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
    uint32_t g = ceil_div<uint32_t>(N, 256);
    if (g > (uint32_t)sm_count() * 4) g = (uint32_t)sm_count() * 4;
    k_head_loss_fwd<<<g, 256, 0, st>>>(image, weights_sum, aud_sum, eye_sum, gt_rgb, bg_color, bg_per_ray ? 1u : 0u, N, lambda_ent, lambda_amb, loss);
    return check_launch("head_loss_forward");
}

extern "C" int b2n_head_loss_backward(const float *image, const float *weights_sum, const float *gt_rgb, const float *bg_color, int bg_per_ray, uint32_t N,
                                      float lambda_ent, float lambda_amb, const float *grad_loss, float *grad_image, float *grad_weights_sum,
                                      float *grad_aud_sum, float *grad_eye_sum, void *stream) {
    B2N_REQUIRE(image && weights_sum && gt_rgb && bg_color && grad_loss && grad_image && grad_weights_sum && grad_aud_sum && grad_eye_sum, "head_loss_backward: null pointer");
    if (N == 0) return 0;
    uint32_t g = ceil_div<uint32_t>(N, 256);
    if (g > (uint32_t)sm_count() * 4) g = (uint32_t)sm_count() * 4;
    k_head_loss_bwd<<<g, 256, 0, as_stream(stream)>>>(image, weights_sum, gt_rgb, bg_color, bg_per_ray ? 1u : 0u, N, lambda_ent, lambda_amb, grad_loss, grad_image,
                                                      grad_weights_sum, grad_aud_sum, grad_eye_sum);
    return check_launch("head_loss_backward");
}
