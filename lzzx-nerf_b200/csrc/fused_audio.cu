// fused_audio.cu — the per-frame audio prologue (AudioNet + AudioAttNet, nerf_triplane/network.py:9-70, encode_audio :226-240)
// as ONE kernel: a thread-block cluster of 8 CTAs, one CTA per frame of the 8-frame audio window.
//
// The reference runs ~25 tiny cuDNN/cuBLAS launches plus ~50 autocast cast kernels per frame for this (≈0.2 ms of a ≈0.8 ms frame in
// the first profile).  Here every CTA pushes its frame through the four stride-2 Conv1d + LeakyReLU layers and the two Linear layers in
// shared memory (one warp per output element, lanes split the reduction, coalesced weight reads); the 8 resulting 32-vectors are
// exchanged through distributed shared memory (cluster.map_shared_rank) and CTA 0 runs the attention net (5 Conv1d over the 8 frames,
// Linear(8,8), softmax) and the weighted sum.  Numerics follow autocast(fp16): operands rounded to fp16, fp32 accumulation, layer outputs
// rounded to fp16, softmax and the final weighted sum in fp32.
#include "common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace b2n {

constexpr int AU_THREADS = 1024;
constexpr int AU_WARPS = AU_THREADS / 32;
constexpr int AU_FRAMES = 8;

__device__ __forceinline__ float rh(float v) { return __half2float(__float2half_rn(v)); }
__device__ __forceinline__ float leaky_h(float v) { return v > 0.0f ? v : rh(0.02f * v); }       // LeakyReLU(0.02) on a half tensor

// Conv1d(ci -> co, kernel 3, given stride, padding 1) + bias + LeakyReLU on a [ci, li] activation held in shared memory (values already
// fp16-rounded).  One warp per output element (co, lo); lanes split the ci*3 reduction; weights [co, ci, 3] are read coalesced.
// SW = true: w / b point to SHARED memory copies already rounded to fp16 (k_audio_encode stages every small layer at kernel start: a layer then
// costs shared-memory latency instead of one L2 round trip for its weights plus one for its bias).
template <bool SW = false>
__device__ void conv3_layer(const float *in_s, uint32_t ci, uint32_t li, const float *__restrict__ w, const float *__restrict__ b, float *out_s, uint32_t co,
                            uint32_t stride, bool act) {
    const uint32_t lo = (li + 2 - 3) / stride + 1;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t o = warp; o < co * lo; o += AU_WARPS) {
        const uint32_t c = o / lo, x = o - c * lo;
        const float *wc = w + (size_t)c * ci * 3;
        float acc = 0.0f;
        // weights are fetched eight at a time, unconditionally (the tap test guarded the load before: no two loads of a lane were ever in flight, and
        // the first convolution of a HuBERT window — 96 dependent L2 round trips per lane — was most of this kernel's 45 us); same summation order
        const uint32_t n = ci * 3;
        for (uint32_t r0 = lane; r0 < n; r0 += 32 * 8) {
            float wv[8];
#pragma unroll
            for (uint32_t u = 0; u < 8; u++) { const uint32_t r = r0 + 32 * u; wv[u] = r < n ? (SW ? wc[r] : __ldg(wc + r)) : 0.0f; }
#pragma unroll
            for (uint32_t u = 0; u < 8; u++) {
                const uint32_t r = r0 + 32 * u;
                if (r < n) {
                    const uint32_t cin = r / 3, k = r - cin * 3;
                    const int xi = (int)(x * stride + k) - 1;
                    if (xi >= 0 && xi < (int)li) acc = fmaf(SW ? wv[u] : rh(wv[u]), in_s[cin * li + xi], acc);
                }
            }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
        if (lane == 0) {
            float v = rh(acc + (SW ? b[c] : rh(__ldg(b + c))));
            out_s[c * lo + x] = act ? leaky_h(v) : v;
        }
    }
    __syncthreads();
}

// Linear(ci -> co) + bias (+ LeakyReLU); one warp per output
template <bool SW = false>
__device__ void linear_layer(const float *in_s, uint32_t ci, const float *__restrict__ w, const float *__restrict__ b, float *out_s, uint32_t co, bool act) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t o = warp; o < co; o += AU_WARPS) {
        float acc = 0.0f;
        for (uint32_t r = lane; r < ci; r += 32) acc = fmaf(SW ? w[(size_t)o * ci + r] : rh(__ldg(w + (size_t)o * ci + r)), in_s[r], acc);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
        if (lane == 0) {
            float v = rh(acc + (SW ? b[o] : rh(__ldg(b + o))));
            out_s[o] = act ? leaky_h(v) : v;
        }
    }
    __syncthreads();
}

// shared-memory weight cache of k_audio_encode (floats): everything except the first convolution (dim_in x 32 x 3: 393 KB for HuBERT features)
constexpr uint32_t AW_C1 = 0, AW_C2 = AW_C1 + 32 * 32 * 3, AW_C3 = AW_C2 + 64 * 32 * 3, AW_F0 = AW_C3 + 64 * 64 * 3, AW_F1 = AW_F0 + 64 * 64,
                   AW_B = AW_F1 + 32 * 64,                      // biases: conv 32, 32, 64, 64, fc 64, 32
                   AW_AC = AW_B + 288,                         // attention convs 32->16->8->4->2->1
                   AW_AB = AW_AC + (16 * 32 + 8 * 16 + 4 * 8 + 2 * 4 + 1 * 2) * 3,      // their biases 16, 8, 4, 2, 1 (+1 pad)
                   AW_AF = AW_AB + 32, AW_AFB = AW_AF + 64, AW_TOTAL = AW_AFB + 8;

__constant__ uint32_t c_att_w_off[5] = {0, 16 * 32 * 3, (16 * 32 + 8 * 16) * 3, (16 * 32 + 8 * 16 + 4 * 8) * 3, (16 * 32 + 8 * 16 + 4 * 8 + 2 * 4) * 3};
__constant__ uint32_t c_att_b_off[5] = {0, 16, 24, 28, 30};

__device__ __forceinline__ void stage_weights(float *dst, const float *__restrict__ src, uint32_t n) {
#pragma unroll 8
    for (uint32_t i = threadIdx.x; i < n; i += AU_THREADS) dst[i] = rh(__ldg(src + i));
}

struct AudioArgs {
    b2n_audio_weights w;
    const float *auds;      // [8, dim_in, L]
    uint32_t L;             // samples per frame (the reference slices x[:, :, 8-8:8+8], i.e. min(L, 16))
    float *enc_a;           // [32]
    float *lips_state;      // optional [33]: the previous frame's (smoothed) enc_a + a valid flag — smooth_lips, renderer.py:456-460
    float lips_lambda, lips_one_minus;
};

__global__ void __cluster_dims__(AU_FRAMES, 1, 1) __launch_bounds__(AU_THREADS) k_audio_encode(const __grid_constant__ AudioArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ float sm[];
    const uint32_t frame = cluster.block_rank();
    const uint32_t dim_in = a.w.dim_in, Lw = a.L < 16 ? a.L : 16;
    float *s_in = sm;                                   // [dim_in, Lw]
    float *s_a = s_in + dim_in * Lw;                    // ping  (<= 64 * 8)
    float *s_b = s_a + 512;                             // pong
    __shared__ float s_feat[32];                        // this frame's AudioNet output (read by CTA 0 through DSMEM)
    __shared__ float s_all[AU_FRAMES * 32];             // CTA 0: all frames [8, 32]
    __shared__ float s_y[AU_FRAMES];

    // every layer but the first convolution: weights and biases into shared memory (fp16-rounded), all loads issued back to back — the ~13 dependent
    // layers that follow then pay shared-memory latency instead of two L2 round trips each (the kernel is a latency chain; with the batched weight loads of
    // conv3_layer: 45 -> 33 us in the training step, the backward twin 130 -> 112 us)
    float *s_w = s_b + 512;
    stage_weights(s_w + AW_C1, a.w.conv_w[1], 32 * 32 * 3); stage_weights(s_w + AW_C2, a.w.conv_w[2], 64 * 32 * 3); stage_weights(s_w + AW_C3, a.w.conv_w[3], 64 * 64 * 3);
    stage_weights(s_w + AW_F0, a.w.fc_w[0], 64 * 64); stage_weights(s_w + AW_F1, a.w.fc_w[1], 32 * 64);
    stage_weights(s_w + AW_B, a.w.conv_b[0], 32); stage_weights(s_w + AW_B + 32, a.w.conv_b[1], 32); stage_weights(s_w + AW_B + 64, a.w.conv_b[2], 64);
    stage_weights(s_w + AW_B + 128, a.w.conv_b[3], 64); stage_weights(s_w + AW_B + 192, a.w.fc_b[0], 64); stage_weights(s_w + AW_B + 256, a.w.fc_b[1], 32);
    if (frame == 0) {
        const uint32_t chans_[6] = {32, 16, 8, 4, 2, 1};
#pragma unroll
        for (int l = 0; l < 5; l++) {
            stage_weights(s_w + AW_AC + c_att_w_off[l], a.w.att_conv_w[l], chans_[l + 1] * chans_[l] * 3);
            stage_weights(s_w + AW_AB + c_att_b_off[l], a.w.att_conv_b[l], chans_[l + 1]);
        }
        stage_weights(s_w + AW_AF, a.w.att_fc_w, 64); stage_weights(s_w + AW_AFB, a.w.att_fc_b, 8);
    }
    // window slice + fp16 rounding of the input (autocast casts the conv input to half)
    const float *x = a.auds + (size_t)frame * dim_in * a.L;
    for (uint32_t i = threadIdx.x; i < dim_in * Lw; i += AU_THREADS) {
        const uint32_t c = i / Lw, t = i - c * Lw;
        s_in[i] = rh(__ldg(x + (size_t)c * a.L + t));
    }
    __syncthreads();
    // AudioNet.encoder_conv: dim_in -> 32 -> 32 -> 64 -> 64, stride 2 (network.py:45-54)
    uint32_t li = Lw;
    conv3_layer<false>(s_in, dim_in, li, a.w.conv_w[0], a.w.conv_b[0], s_a, 32, 2, true);                     li = (li + 2 - 3) / 2 + 1;      // weights too large to stage
    conv3_layer<true>(s_a, 32, li, s_w + AW_C1, s_w + AW_B + 32, s_b, 32, 2, true);                            li = (li + 2 - 3) / 2 + 1;
    conv3_layer<true>(s_b, 32, li, s_w + AW_C2, s_w + AW_B + 64, s_a, 64, 2, true);                            li = (li + 2 - 3) / 2 + 1;
    conv3_layer<true>(s_a, 64, li, s_w + AW_C3, s_w + AW_B + 128, s_b, 64, 2, true);                           li = (li + 2 - 3) / 2 + 1;
    // li == 1 here (squeeze(-1), network.py:66); encoder_fc1: 64 -> 64 -> 32 (network.py:55-59)
    linear_layer<true>(s_b, 64, s_w + AW_F0, s_w + AW_B + 192, s_a, 64, true);
    linear_layer<true>(s_a, 64, s_w + AW_F1, s_w + AW_B + 256, s_feat, 32, false);
    cluster.sync();                                     // every frame's feature vector is in its CTA's shared memory
    if (frame == 0) {
        for (uint32_t i = threadIdx.x; i < AU_FRAMES * 32; i += AU_THREADS) {
            const float *remote = cluster.map_shared_rank(s_feat, i >> 5);       // distributed shared memory read
            s_all[i] = remote[i & 31];
        }
        __syncthreads();
    }
    cluster.sync();                                     // peers may exit only after CTA 0 has read their shared memory
    if (frame != 0) return;
    // AudioAttNet (network.py:9-36): y = x^T [32, 8] -> conv 32->16->8->4->2->1 (k3, s1, p1, LeakyReLU) -> Linear(8,8) -> softmax -> sum_s y[s] x[s,:]
    float *t0 = s_a, *t1 = s_b;
    for (uint32_t i = threadIdx.x; i < AU_FRAMES * 32; i += AU_THREADS) t0[(i & 31) * AU_FRAMES + (i >> 5)] = s_all[i];    // [32 channels, 8 frames]
    __syncthreads();
    const uint32_t chans[6] = {32, 16, 8, 4, 2, 1};
#pragma unroll 1
    for (int l = 0; l < 5; l++) {
        conv3_layer<true>(t0, chans[l], AU_FRAMES, s_w + AW_AC + c_att_w_off[l], s_w + AW_AB + c_att_b_off[l], t1, chans[l + 1], 1, true);
        float *tmp = t0; t0 = t1; t1 = tmp;
    }
    linear_layer<true>(t0, AU_FRAMES, s_w + AW_AF, s_w + AW_AFB, s_y, AU_FRAMES, false);
    if (threadIdx.x < 32) {
        // softmax over the 8 frames in fp32 (autocast runs softmax in float), then the weighted sum in fp32
        float mx = -INFINITY;
        for (int s = 0; s < AU_FRAMES; s++) mx = fmaxf(mx, s_y[s]);
        float den = 0.0f, e[AU_FRAMES];
        for (int s = 0; s < AU_FRAMES; s++) { e[s] = expf(s_y[s] - mx); den += e[s]; }
        float acc = 0.0f;
        for (int s = 0; s < AU_FRAMES; s++) acc = fmaf(e[s] / den, s_all[s * 32 + threadIdx.x], acc);
        if (a.lips_state != nullptr) {
            // enc_a = lambda * self.enc_a + (1 - lambda) * enc_a; self.enc_a = enc_a  (fp32, two products and a sum like the torch expression)
            const bool have = a.lips_state[32] != 0.0f;
            __syncwarp();
            if (have) acc = __fadd_rn(__fmul_rn(a.lips_lambda, a.lips_state[threadIdx.x]), __fmul_rn(a.lips_one_minus, acc));
            a.lips_state[threadIdx.x] = acc;
            __syncwarp();
            if (threadIdx.x == 0) a.lips_state[32] = 1.0f;
        }
        a.enc_a[threadIdx.x] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------
// backward (training): d enc_a [32] -> gradients of every AudioNet / AudioAttNet parameter
// ---------------------------------------------------------------------------------------------------
// Same cluster of 8 CTAs.  Every CTA recomputes its frame's forward keeping all activations in shared memory; CTA 0 gathers the eight feature
// vectors (DSMEM), runs the attention net forward + backward and leaves d features [8, 32] in its shared memory; after a cluster barrier every CTA
// reads its row and back-propagates through its AudioNet copy, adding its weight-gradient contributions with red.global (8-way contention at
// most).  Gradients are computed in fp32 from the forward's fp16-rounded operands (autocast would also round the inter-layer gradients to fp16).
__device__ __forceinline__ float leaky_grad(float a) { return a > 0.0f ? 1.0f : 0.02f; }      // a = LeakyReLU output: same sign as its input

// conv backward.  in_s [ci, li] forward input, out_s [co, lo] forward OUTPUT (post-activation, for the LeakyReLU slope; NULL = no activation),
// dout_s [co, lo] gradient w.r.t. that output (overwritten with the pre-activation gradient), din_s [ci, li] (NULL = not needed)
__device__ void conv3_backward(const float *in_s, uint32_t ci, uint32_t li, const float *__restrict__ w, const float *out_s, float *dout_s, uint32_t co, uint32_t stride,
                               float *__restrict__ gw, float *__restrict__ gb, float *din_s) {
    const uint32_t lo = (li + 2 - 3) / stride + 1;
    if (out_s) {
        for (uint32_t o = threadIdx.x; o < co * lo; o += AU_THREADS) dout_s[o] *= leaky_grad(out_s[o]);
        __syncthreads();
    }
    for (uint32_t r = threadIdx.x; r < co * ci * 3; r += AU_THREADS) {          // d W[c, cin, k] = sum_x dz[c, x] in[cin, x * stride + k - 1]
        const uint32_t c = r / (ci * 3), rem = r - c * ci * 3, cin = rem / 3, k = rem - cin * 3;
        float acc = 0.0f;
        for (uint32_t x = 0; x < lo; x++) {
            const int xi = (int)(x * stride + k) - 1;
            if (xi >= 0 && xi < (int)li) acc = fmaf(dout_s[c * lo + x], in_s[cin * li + xi], acc);
        }
        if (acc != 0.0f) atomicAdd(gw + r, acc);
    }
    for (uint32_t c = threadIdx.x; c < co; c += AU_THREADS) {
        float acc = 0.0f;
        for (uint32_t x = 0; x < lo; x++) acc += dout_s[c * lo + x];
        atomicAdd(gb + c, acc);
    }
    if (din_s) {
        // d in[cin, xi] = sum_c sum_k dz[c, x] W[c, cin, k] with x * stride + k - 1 == xi: one warp per element, lanes split the (c, k) pairs
        const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (uint32_t o = warp; o < ci * li; o += AU_WARPS) {
            const uint32_t cin = o / li, xi = o - cin * li;
            float acc = 0.0f;
            for (uint32_t r = lane; r < co * 3; r += 32) {
                const uint32_t c = r / 3, k = r - c * 3;
                const int num = (int)xi + 1 - (int)k;
                if (num >= 0 && num % (int)stride == 0 && (uint32_t)num / stride < lo)
                    acc = fmaf(dout_s[c * lo + (uint32_t)num / stride], rh(__ldg(w + ((size_t)c * ci + cin) * 3 + k)), acc);
            }
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
            if (lane == 0) din_s[o] = acc;
        }
    }
    __syncthreads();
}
// linear backward: in_s [ci], out_s [co] post-activation output or NULL, dout_s [co] (overwritten), din_s [ci] or NULL
__device__ void linear_backward(const float *in_s, uint32_t ci, const float *__restrict__ w, const float *out_s, float *dout_s, uint32_t co, float *__restrict__ gw,
                                float *__restrict__ gb, float *din_s) {
    if (out_s) {
        for (uint32_t o = threadIdx.x; o < co; o += AU_THREADS) dout_s[o] *= leaky_grad(out_s[o]);
        __syncthreads();
    }
    for (uint32_t r = threadIdx.x; r < co * ci; r += AU_THREADS) {
        const uint32_t o = r / ci, i = r - o * ci;
        const float v = dout_s[o] * in_s[i];
        if (v != 0.0f) atomicAdd(gw + r, v);
    }
    for (uint32_t o = threadIdx.x; o < co; o += AU_THREADS) atomicAdd(gb + o, dout_s[o]);
    if (din_s) {
        const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (uint32_t i = warp; i < ci; i += AU_WARPS) {              // one warp per input element, lanes split the outputs
            float acc = 0.0f;
            for (uint32_t o = lane; o < co; o += 32) acc = fmaf(dout_s[o], rh(__ldg(w + (size_t)o * ci + i)), acc);
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
            if (lane == 0) din_s[i] = acc;
        }
    }
    __syncthreads();
}

struct AudioBwdArgs {
    b2n_audio_weights w;
    b2n_audio_grads g;
    const float *auds;
    uint32_t L;
    const float *d_enc_a;       // [32]
};

__global__ void __cluster_dims__(AU_FRAMES, 1, 1) __launch_bounds__(AU_THREADS) k_audio_backward(const __grid_constant__ AudioBwdArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ float sm[];
    const uint32_t frame = cluster.block_rank();
    const uint32_t dim_in = a.w.dim_in, Lw = a.L < 16 ? a.L : 16;
    uint32_t l1 = (Lw + 2 - 3) / 2 + 1, l2 = (l1 + 2 - 3) / 2 + 1, l3 = (l2 + 2 - 3) / 2 + 1;      // l4 == 1 (checked on the host)
    // activations of this frame (all kept)                      sizes with Lw = 16: 32*8, 32*4, 64*2, 64, 64, 32
    float *s_in = sm, *s_a1 = s_in + dim_in * Lw, *s_a2 = s_a1 + 256, *s_a3 = s_a2 + 128, *s_a4 = s_a3 + 128, *s_g1 = s_a4 + 64, *s_ga = s_g1 + 64, *s_gb = s_ga + 512;
    __shared__ float s_feat[32];
    __shared__ float s_all[AU_FRAMES * 32];             // CTA 0: features of all frames [8, 32]
    __shared__ float s_dX[AU_FRAMES * 32];              // CTA 0: d features
    __shared__ float s_c[5][16 * AU_FRAMES];            // CTA 0: attention conv outputs (post-activation)
    __shared__ float s_t0[32 * AU_FRAMES];
    __shared__ float s_y[AU_FRAMES], s_p[AU_FRAMES], s_dy[AU_FRAMES];

    // ---- forward recompute (same arithmetic as k_audio_encode) ----------------------------------------------------------------
    const float *x = a.auds + (size_t)frame * dim_in * a.L;
    for (uint32_t i = threadIdx.x; i < dim_in * Lw; i += AU_THREADS) {
        const uint32_t c = i / Lw, t = i - c * Lw;
        s_in[i] = rh(__ldg(x + (size_t)c * a.L + t));
    }
    __syncthreads();
    conv3_layer(s_in, dim_in, Lw, a.w.conv_w[0], a.w.conv_b[0], s_a1, 32, 2, true);
    conv3_layer(s_a1, 32, l1, a.w.conv_w[1], a.w.conv_b[1], s_a2, 32, 2, true);
    conv3_layer(s_a2, 32, l2, a.w.conv_w[2], a.w.conv_b[2], s_a3, 64, 2, true);
    conv3_layer(s_a3, 64, l3, a.w.conv_w[3], a.w.conv_b[3], s_a4, 64, 2, true);
    linear_layer(s_a4, 64, a.w.fc_w[0], a.w.fc_b[0], s_g1, 64, true);
    linear_layer(s_g1, 64, a.w.fc_w[1], a.w.fc_b[1], s_feat, 32, false);
    cluster.sync();
    if (frame == 0) {
        for (uint32_t i = threadIdx.x; i < AU_FRAMES * 32; i += AU_THREADS) {
            const float *remote = cluster.map_shared_rank(s_feat, i >> 5);
            s_all[i] = remote[i & 31];
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < AU_FRAMES * 32; i += AU_THREADS) s_t0[(i & 31) * AU_FRAMES + (i >> 5)] = s_all[i];      // [32 channels, 8 frames]
        __syncthreads();
        const uint32_t chans[6] = {32, 16, 8, 4, 2, 1};
        const float *prev = s_t0;
#pragma unroll 1
        for (int l = 0; l < 5; l++) {
            conv3_layer(prev, chans[l], AU_FRAMES, a.w.att_conv_w[l], a.w.att_conv_b[l], s_c[l], chans[l + 1], 1, true);
            prev = s_c[l];
        }
        linear_layer(s_c[4], AU_FRAMES, a.w.att_fc_w, a.w.att_fc_b, s_y, AU_FRAMES, false);
        if (threadIdx.x == 0) {
            float mx = -INFINITY, den = 0.0f;
            for (int q = 0; q < AU_FRAMES; q++) mx = fmaxf(mx, s_y[q]);
            for (int q = 0; q < AU_FRAMES; q++) { s_p[q] = expf(s_y[q] - mx); den += s_p[q]; }
            for (int q = 0; q < AU_FRAMES; q++) s_p[q] /= den;
            // enc_a[j] = sum_s p[s] X[s, j]:  d p[s] = sum_j d_out[j] X[s, j];  softmax backward: d y = p * (d p - sum p d p)
            float dp[AU_FRAMES], dot = 0.0f;
            for (int q = 0; q < AU_FRAMES; q++) {
                float acc = 0.0f;
                for (int j = 0; j < 32; j++) acc = fmaf(a.d_enc_a[j], s_all[q * 32 + j], acc);
                dp[q] = acc; dot = fmaf(s_p[q], acc, dot);
            }
            for (int q = 0; q < AU_FRAMES; q++) s_dy[q] = s_p[q] * (dp[q] - dot);
        }
        __syncthreads();
        // ---- attention backward ------------------------------------------------------------------------------------------------
        float *d_cur = s_ga, *d_nxt = s_gb;                       // gradient w.r.t. the current layer's output / input
        linear_backward(s_c[4], AU_FRAMES, a.w.att_fc_w, nullptr, s_dy, AU_FRAMES, a.g.att_fc_w, a.g.att_fc_b, d_cur);      // d c5 [1, 8]
#pragma unroll 1
        for (int l = 4; l >= 0; l--) {
            const float *in_l = l == 0 ? s_t0 : s_c[l - 1];
            conv3_backward(in_l, chans[l], AU_FRAMES, a.w.att_conv_w[l], s_c[l], d_cur, chans[l + 1], 1, a.g.att_conv_w[l], a.g.att_conv_b[l], d_nxt);
            float *tmp = d_cur; d_cur = d_nxt; d_nxt = tmp;
        }
        // d_cur = d t0 [32, 8];  d X[s, j] = p[s] d_out[j] + d t0[j, s]
        for (uint32_t i = threadIdx.x; i < AU_FRAMES * 32; i += AU_THREADS) {
            const uint32_t q = i >> 5, j = i & 31;
            s_dX[i] = fmaf(s_p[q], a.d_enc_a[j], d_cur[j * AU_FRAMES + q]);
        }
        __syncthreads();
    }
    cluster.sync();                                               // d X is ready in CTA 0
    {
        const float *remote = cluster.map_shared_rank(s_dX, 0);
        if (threadIdx.x < 32) s_ga[threadIdx.x] = remote[frame * 32 + threadIdx.x];      // d feat of this frame
    }
    __syncthreads();
    cluster.sync();                                               // CTA 0 may not exit / reuse its shared memory before every peer has read
    // ---- AudioNet backward for this frame -------------------------------------------------------------------------------------------
    float *d_cur = s_ga, *d_nxt = s_gb;
    linear_backward(s_g1, 64, a.w.fc_w[1], nullptr, d_cur, 32, a.g.fc_w[1], a.g.fc_b[1], d_nxt);                  // -> d g1 [64]
    linear_backward(s_a4, 64, a.w.fc_w[0], s_g1, d_nxt, 64, a.g.fc_w[0], a.g.fc_b[0], d_cur);                     // -> d a4 [64]
    conv3_backward(s_a3, 64, l3, a.w.conv_w[3], s_a4, d_cur, 64, 2, a.g.conv_w[3], a.g.conv_b[3], d_nxt);         // -> d a3 [64, l3]
    conv3_backward(s_a2, 32, l2, a.w.conv_w[2], s_a3, d_nxt, 64, 2, a.g.conv_w[2], a.g.conv_b[2], d_cur);         // -> d a2 [32, l2]
    conv3_backward(s_a1, 32, l1, a.w.conv_w[1], s_a2, d_cur, 32, 2, a.g.conv_w[1], a.g.conv_b[1], d_nxt);         // -> d a1 [32, l1]
    conv3_backward(s_in, dim_in, Lw, a.w.conv_w[0], s_a1, d_nxt, 32, 2, a.g.conv_w[0], a.g.conv_b[0], nullptr);
}

}  // namespace b2n

using namespace b2n;

extern "C" int b2n_audio_encode(const b2n_audio_weights *w, const float *auds, uint32_t L, float *enc_a, void *stream) {
    return b2n_audio_encode_smooth(w, auds, L, enc_a, nullptr, 0.0f, stream);
}

extern "C" int b2n_audio_encode_smooth(const b2n_audio_weights *w, const float *auds, uint32_t L, float *enc_a, float *lips_state, float lambda, void *stream) {
    B2N_REQUIRE(w && auds && enc_a, "audio_encode: null pointer");
    B2N_REQUIRE(w->dim_in >= 1 && w->dim_in <= 4096 && L >= 1, "audio_encode: dim_in=%u / L=%u unsupported", w->dim_in, L);
    for (int i = 0; i < 4; i++) B2N_REQUIRE(w->conv_w[i] && w->conv_b[i], "audio_encode: null conv weight");
    for (int i = 0; i < 5; i++) B2N_REQUIRE(w->att_conv_w[i] && w->att_conv_b[i], "audio_encode: null attention conv weight");
    B2N_REQUIRE(w->fc_w[0] && w->fc_w[1] && w->fc_b[0] && w->fc_b[1] && w->att_fc_w && w->att_fc_b, "audio_encode: null linear weight");
    const uint32_t Lw = L < 16 ? L : 16;
    // the conv stack must collapse the window to length 1 (the reference squeezes the last dim, network.py:66)
    uint32_t li = Lw;
    for (int i = 0; i < 4; i++) li = (li + 2 - 3) / 2 + 1;
    B2N_REQUIRE(li == 1, "audio_encode: window length %u does not reduce to 1 after four stride-2 convolutions", L);
    AudioArgs a = {*w, auds, L, enc_a, lips_state, lambda, (float)(1.0 - (double)lambda)};
    const size_t smem = sizeof(float) * ((size_t)w->dim_in * Lw + 1024 + AW_TOTAL);       // input slice + ping / pong + the weight cache (121 KB)
    B2N_SMEM(k_audio_encode, smem);
    k_audio_encode<<<AU_FRAMES, AU_THREADS, smem, as_stream(stream)>>>(a);
    return check_launch("audio_encode");
}

extern "C" int b2n_audio_backward(const b2n_audio_weights *w, const float *auds, uint32_t L, const float *d_enc_a, const b2n_audio_grads *g, void *stream) {
    B2N_REQUIRE(w && auds && d_enc_a && g, "audio_backward: null pointer");
    B2N_REQUIRE(w->dim_in >= 1 && w->dim_in <= 4096 && L >= 1, "audio_backward: dim_in=%u / L=%u unsupported", w->dim_in, L);
    for (int i = 0; i < 4; i++) B2N_REQUIRE(w->conv_w[i] && w->conv_b[i] && g->conv_w[i] && g->conv_b[i], "audio_backward: null conv weight / gradient");
    for (int i = 0; i < 5; i++) B2N_REQUIRE(w->att_conv_w[i] && w->att_conv_b[i] && g->att_conv_w[i] && g->att_conv_b[i], "audio_backward: null attention conv weight / gradient");
    B2N_REQUIRE(w->fc_w[0] && w->fc_w[1] && w->fc_b[0] && w->fc_b[1] && w->att_fc_w && w->att_fc_b && g->fc_w[0] && g->fc_w[1] && g->fc_b[0] && g->fc_b[1] &&
                g->att_fc_w && g->att_fc_b, "audio_backward: null linear weight / gradient");
    const uint32_t Lw = L < 16 ? L : 16;
    uint32_t li = Lw;
    for (int i = 0; i < 4; i++) li = (li + 2 - 3) / 2 + 1;
    B2N_REQUIRE(li == 1, "audio_backward: window length %u does not reduce to 1 after four stride-2 convolutions", L);
    AudioBwdArgs a = {*w, *g, auds, L, d_enc_a};
    const size_t smem = sizeof(float) * ((size_t)w->dim_in * Lw + 256 + 128 + 128 + 64 + 64 + 512 + 512);
    B2N_SMEM(k_audio_backward, smem);
    k_audio_backward<<<AU_FRAMES, AU_THREADS, smem, as_stream(stream)>>>(a);
    return check_launch("audio_backward");
}
