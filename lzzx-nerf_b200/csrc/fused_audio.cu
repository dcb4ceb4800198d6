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

constexpr int AU_THREADS = 256;
constexpr int AU_WARPS = AU_THREADS / 32;
constexpr int AU_FRAMES = 8;

__device__ __forceinline__ float rh(float v) { return __half2float(__float2half_rn(v)); }
__device__ __forceinline__ float leaky_h(float v) { return v > 0.0f ? v : rh(0.02f * v); }       // LeakyReLU(0.02) on a half tensor

// Conv1d(ci -> co, kernel 3, given stride, padding 1) + bias + LeakyReLU on a [ci, li] activation held in shared memory (values already
// fp16-rounded).  One warp per output element (co, lo); lanes split the ci*3 reduction; weights [co, ci, 3] are read coalesced.
__device__ void conv3_layer(const float *in_s, uint32_t ci, uint32_t li, const float *__restrict__ w, const float *__restrict__ b, float *out_s, uint32_t co,
                            uint32_t stride, bool act) {
    const uint32_t lo = (li + 2 - 3) / stride + 1;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t o = warp; o < co * lo; o += AU_WARPS) {
        const uint32_t c = o / lo, x = o - c * lo;
        const float *wc = w + (size_t)c * ci * 3;
        float acc = 0.0f;
        for (uint32_t r = lane; r < ci * 3; r += 32) {
            const uint32_t cin = r / 3, k = r - cin * 3;
            const int xi = (int)(x * stride + k) - 1;
            if (xi >= 0 && xi < (int)li) acc = fmaf(rh(__ldg(wc + r)), in_s[cin * li + xi], acc);
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
        if (lane == 0) {
            float v = rh(acc + rh(__ldg(b + c)));
            out_s[c * lo + x] = act ? leaky_h(v) : v;
        }
    }
    __syncthreads();
}

// Linear(ci -> co) + bias (+ LeakyReLU); one warp per output
__device__ void linear_layer(const float *in_s, uint32_t ci, const float *__restrict__ w, const float *__restrict__ b, float *out_s, uint32_t co, bool act) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t o = warp; o < co; o += AU_WARPS) {
        float acc = 0.0f;
        for (uint32_t r = lane; r < ci; r += 32) acc = fmaf(rh(__ldg(w + (size_t)o * ci + r)), in_s[r], acc);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
        if (lane == 0) {
            float v = rh(acc + rh(__ldg(b + o)));
            out_s[o] = act ? leaky_h(v) : v;
        }
    }
    __syncthreads();
}

struct AudioArgs {
    b2n_audio_weights w;
    const float *auds;      // [8, dim_in, L]
    uint32_t L;             // samples per frame (the reference slices x[:, :, 8-8:8+8], i.e. min(L, 16))
    float *enc_a;           // [32]
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

    // window slice + fp16 rounding of the input (autocast casts the conv input to half)
    const float *x = a.auds + (size_t)frame * dim_in * a.L;
    for (uint32_t i = threadIdx.x; i < dim_in * Lw; i += AU_THREADS) {
        const uint32_t c = i / Lw, t = i - c * Lw;
        s_in[i] = rh(__ldg(x + (size_t)c * a.L + t));
    }
    __syncthreads();
    // AudioNet.encoder_conv: dim_in -> 32 -> 32 -> 64 -> 64, stride 2 (network.py:45-54)
    uint32_t li = Lw;
    conv3_layer(s_in, dim_in, li, a.w.conv_w[0], a.w.conv_b[0], s_a, 32, 2, true); li = (li + 2 - 3) / 2 + 1;
    conv3_layer(s_a, 32, li, a.w.conv_w[1], a.w.conv_b[1], s_b, 32, 2, true);      li = (li + 2 - 3) / 2 + 1;
    conv3_layer(s_b, 32, li, a.w.conv_w[2], a.w.conv_b[2], s_a, 64, 2, true);      li = (li + 2 - 3) / 2 + 1;
    conv3_layer(s_a, 64, li, a.w.conv_w[3], a.w.conv_b[3], s_b, 64, 2, true);      li = (li + 2 - 3) / 2 + 1;
    // li == 1 here (squeeze(-1), network.py:66); encoder_fc1: 64 -> 64 -> 32 (network.py:55-59)
    linear_layer(s_b, 64, a.w.fc_w[0], a.w.fc_b[0], s_a, 64, true);
    linear_layer(s_a, 64, a.w.fc_w[1], a.w.fc_b[1], s_feat, 32, false);
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
        conv3_layer(t0, chans[l], AU_FRAMES, a.w.att_conv_w[l], a.w.att_conv_b[l], t1, chans[l + 1], 1, true);
        float *tmp = t0; t0 = t1; t1 = tmp;
    }
    linear_layer(t0, AU_FRAMES, a.w.att_fc_w, a.w.att_fc_b, s_y, AU_FRAMES, false);
    if (threadIdx.x < 32) {
        // softmax over the 8 frames in fp32 (autocast runs softmax in float), then the weighted sum in fp32
        float mx = -INFINITY;
        for (int s = 0; s < AU_FRAMES; s++) mx = fmaxf(mx, s_y[s]);
        float den = 0.0f, e[AU_FRAMES];
        for (int s = 0; s < AU_FRAMES; s++) { e[s] = expf(s_y[s] - mx); den += e[s]; }
        float acc = 0.0f;
        for (int s = 0; s < AU_FRAMES; s++) acc = fmaf(e[s] / den, s_all[s * 32 + threadIdx.x], acc);
        a.enc_a[threadIdx.x] = acc;
    }
}

}  // namespace b2n

using namespace b2n;

extern "C" int b2n_audio_encode(const b2n_audio_weights *w, const float *auds, uint32_t L, float *enc_a, void *stream) {
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
    AudioArgs a = {*w, auds, L, enc_a};
    const size_t smem = sizeof(float) * ((size_t)w->dim_in * Lw + 1024);
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) { B2N_CUDA(cudaFuncSetAttribute(k_audio_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); smem_set = smem; }
    k_audio_encode<<<AU_FRAMES, AU_THREADS, smem, as_stream(stream)>>>(a);
    return check_launch("audio_encode");
}
