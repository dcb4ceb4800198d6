// loss.cu — the head branch's training loss and its gradient (TrainerUtil.py:238-334; renderer.py:559-561 for the background blend), per ray n:
//     img   = clamp(image + (1 - weights_sum) * bg, 0, 1)                                   mse_n = mean_c (img - gt)^2
//     unc_loss (TrainerUtil.py:256-277), u = uncertainty_sum, sf = step_factor = min(global_step / iters, 1), face = face_mask:
//         w_n    = 0.2 + 0.8 * clamp((1 - sf) + sf * softmax(u)_n * N, 0, 10)                (detached)
//         l_n    = mse_n * w_n + sf * face * (|img - gt|_2 / (2 (u + 1)^2) + log(u + 1)^2 / 2) + 1e-3 * sf * (1 - face) * u       (the norm detached)
//     loss  = mean_n l_n + lambda_ent * mean_n H2(clamp(weights_sum, 1e-5, 1 - 1e-5))                                            (:313-316, 1e-4)
//           + sf * lambda_amb * mean_n (aud_sum * (1 - face))                                                                    (:318-324)
//           + sf * lambda_amb * mean_n (eye_sum / max_steps * aud_sum.detach() * face)                                           (:326-331)
// with H2(a) = -a log2 a - (1 - a) log2 (1 - a).  The reference builds this from ~60 elementwise / reduction kernels over 65 536 rays and autograd
// replays as many backwards; here: a 64-CTA pass for the partial softmax statistics (folded by every CTA of the next two kernels), one pass for the loss,
// one for all five gradients.
// step_factor is read from the device so a CUDA graph of the step can be replayed while it ramps.
#include "common.cuh"

namespace b2n {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// softmax over the rays of the batch (TrainerUtil.py:262), in two levels: LS_BLOCKS CTAs each leave (max, sum exp(u - max)) of their share in
// stats[4 + 2 b], and every CTA of the loss kernels folds the LS_BLOCKS partials (softmax_fold) — no single-CTA pass over the batch, no extra launch.
constexpr uint32_t LS_BLOCKS = 64;
__global__ void __launch_bounds__(256) k_loss_softmax_stats(const float *__restrict__ u, uint32_t N, float *__restrict__ stats) {
    __shared__ float s[8];
    float mx = -INFINITY;
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) mx = fmaxf(mx, u[n]);
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = warp_max(s[threadIdx.x & 7]);
    __syncthreads();
    float se = 0.0f;
    if (mx > -INFINITY)
        for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) se += expf(u[n] - mx);
    se = warp_sum(se);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = se;
    __syncthreads();
    if (threadIdx.x < 32) {
        se = warp_sum(threadIdx.x < 8 ? s[threadIdx.x] : 0.0f);
        if (threadIdx.x == 0) { stats[4 + 2 * blockIdx.x] = mx; stats[5 + 2 * blockIdx.x] = se; }
    }
}
// every thread of the CTA gets (max, sum exp(u - max)) of the whole batch; call from all threads (one barrier)
__device__ __forceinline__ void softmax_fold(const float *__restrict__ stats, float &mx, float &se) {
    __shared__ float s_ms[2];
    if (threadIdx.x < 32) {
        const float m0 = stats[4 + 2 * threadIdx.x], m1 = stats[4 + 2 * (threadIdx.x + 32)];
        const float m = warp_max(fmaxf(m0, m1));
        float v = 0.0f;
        if (m0 > -INFINITY) v += stats[5 + 2 * threadIdx.x] * expf(m0 - m);
        if (m1 > -INFINITY) v += stats[5 + 2 * (threadIdx.x + 32)] * expf(m1 - m);
        v = warp_sum(v);
        if (threadIdx.x == 0) { s_ms[0] = m; s_ms[1] = v; }
    }
    __syncthreads();
    mx = s_ms[0]; se = s_ms[1];
}

struct RayTerms { float img[3], pre[3], d[3], mse, w, sf, face; };

__device__ __forceinline__ float step_factor_of(const b2n_loss_args &a) { return a.step_factor ? a.step_factor[0] : a.step_factor_host; }

__device__ __forceinline__ RayTerms ray_terms(const b2n_loss_args &a, uint32_t n, uint32_t N, float sm_max, float sm_sum) {
    RayTerms t;
    const float k = 1.0f - a.weights_sum[n];
    t.mse = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float b = a.bg_color[a.bg_per_ray ? 3 * (size_t)n + c : c];
        t.pre[c] = a.image[3 * (size_t)n + c] + k * b;
        t.img[c] = fminf(fmaxf(t.pre[c], 0.0f), 1.0f);
        t.d[c] = t.img[c] - a.gt_rgb[3 * (size_t)n + c];
        t.mse += t.d[c] * t.d[c];
    }
    t.mse *= (1.0f / 3.0f);
    t.sf = step_factor_of(a);
    t.face = a.face_mask ? (a.face_mask[n] ? 1.0f : 0.0f) : 1.0f;
    t.w = 1.0f;
    if (a.unc_sum) {
        const float uw = expf(a.unc_sum[n] - sm_max) / sm_sum * (float)N;
        t.w = 0.2f + 0.8f * fminf(fmaxf((1.0f - t.sf) + t.sf * uw, 0.0f), 10.0f);
    }
    return t;
}

__global__ void __launch_bounds__(256) k_head_loss_fwd(const __grid_constant__ b2n_loss_args a, uint32_t N, float *__restrict__ stats) {
    float acc = 0.0f, sm_max = 0.0f, sm_sum = 1.0f;
    if (a.unc_sum) softmax_fold(stats, sm_max, sm_sum);
    const float inv_n = 1.0f / (float)N;
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const RayTerms t = ray_terms(a, n, N, sm_max, sm_sum);
        float l = t.mse * t.w;
        if (a.unc_sum) {
            const float u = a.unc_sum[n], beta = u + 1.0f, lb = logf(beta);
            const float nrm = sqrtf(t.d[0] * t.d[0] + t.d[1] * t.d[1] + t.d[2] * t.d[2]);
            l += t.sf * t.face * (nrm / (2.0f * beta * beta) + 0.5f * lb * lb) + 1e-3f * t.sf * (1.0f - t.face) * u;
        }
        const float w = a.weights_sum[n];
        const float al = fminf(fmaxf(w, 1e-5f), 1.0f - 1e-5f);
        l += a.lambda_ent * (-al * log2f(al) - (1.0f - al) * log2f(1.0f - al));
        const float lam = t.sf * a.lambda_amb;
        if (a.amb_aud_loss) l += lam * a.aud_sum[n] * (1.0f - t.face);
        if (a.amb_eye_loss) l += lam * (a.eye_sum[n] * a.inv_max_steps) * a.aud_sum[n] * t.face;
        acc += l;
    }
    __shared__ float s[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < 8 ? s[threadIdx.x] : 0.0f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(stats, v * inv_n);
    }
}

// gradients of the loss w.r.t. the composite's outputs, times the upstream scalar *g (the GradScaler's scale)
__global__ void __launch_bounds__(256) k_head_loss_bwd(const __grid_constant__ b2n_loss_args a, uint32_t N, const float *__restrict__ stats, const float *__restrict__ g,
                                                        float *__restrict__ d_image, float *__restrict__ d_ws, float *__restrict__ d_aud, float *__restrict__ d_eye,
                                                        float *__restrict__ d_unc) {
    const float up = g[0] / (float)N;
    float sm_max = 0.0f, sm_sum = 1.0f;
    if (a.unc_sum) softmax_fold(stats, sm_max, sm_sum);
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const RayTerms t = ray_terms(a, n, N, sm_max, sm_sum);
        float dw = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float b = a.bg_color[a.bg_per_ray ? 3 * (size_t)n + c : c];
            // clamp passes the gradient where min <= x <= max (torch.clamp backward)
            const float dv = (t.pre[c] >= 0.0f && t.pre[c] <= 1.0f) ? up * t.w * (2.0f / 3.0f) * t.d[c] : 0.0f;
            d_image[3 * (size_t)n + c] = dv;
            dw -= dv * b;
        }
        const float w = a.weights_sum[n];
        if (w >= 1e-5f && w <= 1.0f - 1e-5f) dw += up * a.lambda_ent * (log2f(1.0f - w) - log2f(w));       // d H2 / d a = log2((1 - a) / a)
        d_ws[n] = dw;
        const float lam = t.sf * a.lambda_amb;
        d_aud[n] = a.amb_aud_loss ? up * lam * (1.0f - t.face) : 0.0f;                                     // the cross term sees aud_sum detached
        d_eye[n] = a.amb_eye_loss ? up * lam * a.inv_max_steps * a.aud_sum[n] * t.face : 0.0f;
        if (d_unc) {
            float du = 0.0f;
            if (a.unc_sum) {
                const float beta = a.unc_sum[n] + 1.0f;
                const float nrm = sqrtf(t.d[0] * t.d[0] + t.d[1] * t.d[1] + t.d[2] * t.d[2]);
                du = up * (t.sf * t.face * (-nrm / (beta * beta * beta) + logf(beta) / beta) + 1e-3f * t.sf * (1.0f - t.face));
            }
            d_unc[n] = du;
        }
    }
}

}  // namespace b2n

using namespace b2n;

static int check_args(const b2n_loss_args *a, const char *who) {
    B2N_REQUIRE(a && a->image && a->weights_sum && a->aud_sum && a->eye_sum && a->gt_rgb && a->bg_color, "%s: null pointer", who);
    return 0;
}

extern "C" int b2n_head_loss_forward(const b2n_loss_args *a, uint32_t N, float *stats, void *stream) {
    if (int rc = check_args(a, "head_loss_forward")) return rc;
    B2N_REQUIRE(stats, "head_loss_forward: null stats");
    B2N_REQUIRE(N > 0, "head_loss_forward: empty batch");
    cudaStream_t st = as_stream(stream);
    B2N_CUDA(cudaMemsetAsync(stats, 0, sizeof(float), st));
    if (a->unc_sum) {
        k_loss_softmax_stats<<<LS_BLOCKS, 256, 0, st>>>(a->unc_sum, N, stats);
        if (check_launch("head_loss_forward(stats)")) return 1;
    }
    uint32_t g = ceil_div<uint32_t>(N, 256);
    if (g > (uint32_t)sm_count() * 4) g = (uint32_t)sm_count() * 4;
    k_head_loss_fwd<<<g, 256, 0, st>>>(*a, N, stats);
    return check_launch("head_loss_forward");
}

extern "C" int b2n_head_loss_backward(const b2n_loss_args *a, uint32_t N, const float *stats, const float *grad_loss, float *grad_image, float *grad_weights_sum,
                                      float *grad_aud_sum, float *grad_eye_sum, float *grad_unc_sum, void *stream) {
    if (int rc = check_args(a, "head_loss_backward")) return rc;
    B2N_REQUIRE(stats && grad_loss && grad_image && grad_weights_sum && grad_aud_sum && grad_eye_sum, "head_loss_backward: null pointer");
    if (N == 0) return 0;
    uint32_t g = ceil_div<uint32_t>(N, 256);
    if (g > (uint32_t)sm_count() * 4) g = (uint32_t)sm_count() * 4;
    k_head_loss_bwd<<<g, 256, 0, as_stream(stream)>>>(*a, N, stats, grad_loss, grad_image, grad_weights_sum, grad_aud_sum, grad_eye_sum, grad_unc_sum);
    return check_launch("head_loss_backward");
}
