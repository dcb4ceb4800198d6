// fused_torso.cu — the torso branch of one frame as ONE kernel.
//
// Replaces NeRFRenderer.run_torso (nerf_triplane/renderer.py:572-631) + NeRFNetwork.forward_torso (network.py:170-205): per pixel
//   occupancy = bilinear sample of the 2-D torso density grid (F.grid_sample, align_corners=True) > threshold            (renderer.py:603-606)
//   enc_x = FreqEncoder(2-D, degree 8)(x * torso_shrink)                                   34 values                     (network.py:176,185)
//   dx = torso_deform_net([enc_x | enc_anchor 42 | ind code 8])        84 -> 32 -> 32 -> 2                              (network.py:187-192)
//   g  = torso_encoder(clamp(x + dx, -1, 1)): tiled grid D=2 L=16 C=2, fp16 under autocast (grid.py:36-39)   32 values   (network.py:194-196)
//   h  = torso_net([g | enc_x | enc_anchor | ind code])                116 -> 32 -> 32 -> 4                             (network.py:199-201)
//   alpha, color = sigmoid(h) * 1.002 - 0.001;  bg = color * alpha + bg * (1 - alpha)                                   (network.py:203-204, renderer.py:620)
// ~25 launches (grid_sample, mask compaction, two encoders, the table cast to half, six GEMMs, cats / repeats, scatter back) in the reference.
//
// Numerics = the reference under autocast(fp16): Linear inputs / weights / outputs rounded to fp16 with fp32 accumulation, ReLU on the fp16 values,
// the frequency encoding in fp32 (custom_fwd(cast_inputs=float32)) with the per-op kernel's expression, the grid with the per-op kernel's half
// arithmetic (every corner product and every partial sum narrowed to half, gridencoder.cu:161-175 with scalar_t = at::Half), sigmoid in fp16.
//
// Organisation: the MLPs are 5.4 k MACs per pixel with hidden width 32 — CUDA-core work (a 512x512 frame is 1.4 GMAC).  Thread = pixel; the weights sit
// in shared memory TRANSPOSED ([k][32 outputs], pre-rounded to fp16) so that one input value feeds 32 FFMAs from 8 broadcast LDS.128; the per-frame
// constant part of both first layers (enc_anchor, individual code: 50 of the 84 / 116 inputs) is folded into a bias once per CTA; hidden
// activations go through a per-thread shared-memory column (conflict-free [k][thread]) so the k loops stay rolled (small code); the grid features
// (four levels = 16 table reads in flight at a time) and the encoding are produced inside the first-layer loops and never stored.  Persistent CTAs
// take 128-pixel tiles from a device counter: the torso covers part of the image and warps without a torso pixel skip the network, so static
// shares of the image are unbalanced.
#include "common.cuh"

namespace b2n {

constexpr uint32_t TS_THREADS = 128;
constexpr uint32_t TS_ENC = 34, TS_CONST = 50, TS_GRID = 32, TS_H = 32, TS_LEVELS = 16;

struct TorsoArgs {
    const float *bg_coords;
    uint32_t N;
    const float *dgrid;
    uint32_t G;
    float thresh, shrink;
    const float *img;           // packed weight image (k_torso_pack)
    const float *hconst;
    const float *table;
    const int32_t *offsets;
    float S;
    uint32_t H;
    const float *bg_color;
    int bg_per_ray;
    float *bg_out, *alpha_out, *deform_out;
    uint32_t *tile_counter;     // zeroed by k_torso_pack: tiles are handed out dynamically (the torso covers part of the image: static shares are unbalanced)
};

struct TorsoLvl { float scale; uint32_t stride, size, off, mask; };      // mask: 0xffffffff = no wrap (dense level), size - 1 = power-of-two wrap, 0 = generic modulo

__device__ __forceinline__ float rh(float v) { return __half2float(__float2half_rn(v)); }

// value k of FreqEncoder(input_dim = 2, degree = 8)(x): freqencoder.cu:48-57 — same expression as encoders.cu:k_freq_fwd
__device__ __forceinline__ float freq2(float x0, float x1, uint32_t k) {
    if (k < 2) return k ? x1 : x0;
    const uint32_t col = (k >> 1) - 1u, f = col >> 1;
    return __sinf(__fadd_rn(scalbnf((k & 1u) ? x1 : x0, (int)f), (float)(col & 1u) * (3.141592653589793f / 2)));
}

// acc[0..31] += w[0..31] * v   (w: 32 consecutive floats in shared memory, same address for every lane -> broadcast)
__device__ __forceinline__ void axpy32(float (&acc)[TS_H], const float *w, float v) {
#pragma unroll
    for (uint32_t q = 0; q < TS_H / 4; q++) {
        const float4 c = reinterpret_cast<const float4 *>(w)[q];
        acc[4 * q] = __fmaf_rn(c.x, v, acc[4 * q]); acc[4 * q + 1] = __fmaf_rn(c.y, v, acc[4 * q + 1]);
        acc[4 * q + 2] = __fmaf_rn(c.z, v, acc[4 * q + 2]); acc[4 * q + 3] = __fmaf_rn(c.w, v, acc[4 * q + 3]);
    }
}

// packed weight image (floats): every matrix transposed to [input k][outputs] and rounded to fp16; the constant-input rows of both first layers last
constexpr uint32_t TI_WD0 = 0, TI_WD1 = TI_WD0 + TS_ENC * TS_H, TI_WT0 = TI_WD1 + TS_H * TS_H, TI_WT1 = TI_WT0 + (TS_GRID + TS_ENC) * TS_H,
                   TI_WD2 = TI_WT1 + TS_H * TS_H, TI_WT2 = TI_WD2 + TS_H * 2, TI_MAIN = TI_WT2 + TS_H * 4,
                   TI_CD = TI_MAIN, TI_CT = TI_CD + TS_CONST * TS_H, TI_TOTAL = TI_CT + TS_CONST * TS_H;
static_assert(TI_MAIN % 4 == 0, "image copied in 16-byte pieces");

__global__ void __launch_bounds__(256) k_torso_pack(const float *__restrict__ wd0, const float *__restrict__ wd1, const float *__restrict__ wd2,
                                                     const float *__restrict__ wt0, const float *__restrict__ wt1, const float *__restrict__ wt2,
                                                     float *__restrict__ img, uint32_t *__restrict__ tile_counter) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *tile_counter = 0u;
    if (i >= TI_TOTAL) return;
    constexpr uint32_t LD_D = TS_ENC + TS_CONST, LD_T = TS_GRID + TS_ENC + TS_CONST;
    float v;
    if (i < TI_WD1) { const uint32_t k = i / TS_H, j = i % TS_H; v = wd0[j * LD_D + k]; }
    else if (i < TI_WT0) { const uint32_t r = i - TI_WD1, k = r / TS_H, j = r % TS_H; v = wd1[j * TS_H + k]; }
    else if (i < TI_WT1) { const uint32_t r = i - TI_WT0, k = r / TS_H, j = r % TS_H; v = wt0[j * LD_T + k]; }
    else if (i < TI_WD2) { const uint32_t r = i - TI_WT1, k = r / TS_H, j = r % TS_H; v = wt1[j * TS_H + k]; }
    else if (i < TI_WT2) { const uint32_t r = i - TI_WD2; v = wd2[(r & 1u) * TS_H + (r >> 1)]; }
    else if (i < TI_MAIN) { const uint32_t r = i - TI_WT2; v = wt2[(r & 3u) * TS_H + (r >> 2)]; }
    else if (i < TI_CT) { const uint32_t r = i - TI_CD, c = r / TS_H, j = r % TS_H; v = wd0[j * LD_D + TS_ENC + c]; }
    else { const uint32_t r = i - TI_CT, c = r / TS_H, j = r % TS_H; v = wt0[j * LD_T + TS_GRID + TS_ENC + c]; }
    img[i] = rh(v);
}

__global__ void __launch_bounds__(TS_THREADS) k_torso_frame(const __grid_constant__ TorsoArgs a) {
    // transposed, fp16-rounded weights (row k = the output weights of input k), copied as they lie in the packed image (k_torso_pack)
    __shared__ __align__(16) float s_img[TI_MAIN];
    float (*s_wd0)[TS_H] = reinterpret_cast<float (*)[TS_H]>(s_img + TI_WD0);
    float (*s_wd1)[TS_H] = reinterpret_cast<float (*)[TS_H]>(s_img + TI_WD1);
    float (*s_wt0)[TS_H] = reinterpret_cast<float (*)[TS_H]>(s_img + TI_WT0);
    float (*s_wt1)[TS_H] = reinterpret_cast<float (*)[TS_H]>(s_img + TI_WT1);
    float (*s_wd2)[2] = reinterpret_cast<float (*)[2]>(s_img + TI_WD2);
    float (*s_wt2)[4] = reinterpret_cast<float (*)[4]>(s_img + TI_WT2);
    __shared__ float s_bias_d[TS_H], s_bias_t[TS_H];
    __shared__ TorsoLvl s_lvl[TS_LEVELS];
    __shared__ float s_h[TS_H][TS_THREADS];              // hidden activations, column = thread
    const uint32_t tid = threadIdx.x;

    // ---- per-CTA prologue: 22 KB of weights (coalesced 16-byte copies), the two bias vectors, the level geometry ----------------------------
    for (uint32_t i = tid; i < TI_MAIN / 4; i += TS_THREADS) reinterpret_cast<float4 *>(s_img)[i] = __ldg(reinterpret_cast<const float4 *>(a.img) + i);
    if (tid < 2 * TS_H) {                                   // constant inputs folded into a bias (fp32 sum of fp16 products, like the GEMM's)
        const float *cw = a.img + (tid < TS_H ? TI_CD : TI_CT) + (tid & (TS_H - 1u));
        float b = 0.0f;
#pragma unroll 10
        for (uint32_t c = 0; c < TS_CONST; c++) b = __fmaf_rn(__ldg(cw + c * TS_H), rh(__ldg(a.hconst + c)), b);
        (tid < TS_H ? s_bias_d : s_bias_t)[tid & (TS_H - 1u)] = b;
    }
    if (tid >= 2 * TS_H && tid < 2 * TS_H + TS_LEVELS) {    // level geometry with the per-op kernel's arithmetic (gridenc.cu:level_geom)
        const uint32_t l = tid - 2 * TS_H;
        TorsoLvl g;
        g.off = (uint32_t)a.offsets[l];
        g.size = (uint32_t)a.offsets[l + 1] - g.off;
        g.scale = __fmaf_rn(exp2f(__fmul_rn((float)l, a.S)), (float)a.H, -1.0f);
        g.stride = (uint32_t)ceilf(g.scale) + 2u;           // resolution + 1
        g.mask = (g.stride <= g.size && (uint64_t)g.stride * g.stride <= g.size) ? 0xffffffffu : ((g.size & (g.size - 1u)) == 0u ? g.size - 1u : 0u);
        s_lvl[l] = g;
    }
    __syncthreads();

    const float Gm1 = (float)(a.G - 1u);
    __shared__ uint32_t s_tile;
    const uint32_t n_tiles = (a.N + TS_THREADS - 1) / TS_THREADS;
    for (;;) {
        __syncthreads();                                    // everybody has read the previous s_tile
        if (tid == 0) s_tile = atomicAdd(a.tile_counter, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= n_tiles) break;
        const uint32_t n = tile * TS_THREADS + tid;
        const bool live = n < a.N;
        float cx = 0.0f, cy = 0.0f;
        if (live) { cx = __ldcs(a.bg_coords + 2 * (size_t)n); cy = __ldcs(a.bg_coords + 2 * (size_t)n + 1); }
        // ---- occupancy: F.grid_sample(bilinear, zeros padding, align_corners=True) on [1,1,G,G]; grid[..., 0] indexes the width --------------
        bool on = false;
        if (live) {
            const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(cx, 1.0f), 2.0f), Gm1), iy = __fmul_rn(__fdiv_rn(__fadd_rn(cy, 1.0f), 2.0f), Gm1);
            const float fx = floorf(ix), fy = floorf(iy);
            const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
            const float nw = __fmul_rn(__fsub_rn(fx + 1.0f, ix), __fsub_rn(fy + 1.0f, iy)), ne = __fmul_rn(__fsub_rn(ix, fx), __fsub_rn(fy + 1.0f, iy));
            const float sw = __fmul_rn(__fsub_rn(fx + 1.0f, ix), __fsub_rn(iy, fy)), se = __fmul_rn(__fsub_rn(ix, fx), __fsub_rn(iy, fy));
            const int G = (int)a.G;
            auto tap = [&](int x, int y) { return (x >= 0 && x < G && y >= 0 && y < G) ? __ldg(a.dgrid + (size_t)y * G + x) : 0.0f; };
            float occ = 0.0f;
            occ = __fmaf_rn(tap(x0, y0), nw, occ); occ = __fmaf_rn(tap(x1, y0), ne, occ);
            occ = __fmaf_rn(tap(x0, y1), sw, occ); occ = __fmaf_rn(tap(x1, y1), se, occ);
            on = occ > a.thresh;
        }
        float alpha = 0.0f, col[3] = {0.0f, 0.0f, 0.0f}, dx0 = 0.0f, dx1 = 0.0f;
        if (__any_sync(0xffffffffu, on)) {                  // warps entirely outside the torso skip the network
            const float x0 = __fmul_rn(cx, a.shrink), x1 = __fmul_rn(cy, a.shrink);
            float acc[TS_H];
            // ---- deform net layer 0: enc_x part + constant bias
#pragma unroll
            for (uint32_t j = 0; j < TS_H; j++) acc[j] = 0.0f;
#pragma unroll 1
            for (uint32_t k = 0; k < TS_ENC; k++) axpy32(acc, s_wd0[k], rh(freq2(x0, x1, k)));
#pragma unroll
            for (uint32_t j = 0; j < TS_H; j++) s_h[j][tid] = fmaxf(rh(__fadd_rn(acc[j], s_bias_d[j])), 0.0f);
            // ---- layer 1
#pragma unroll
            for (uint32_t j = 0; j < TS_H; j++) acc[j] = 0.0f;
#pragma unroll 1
            for (uint32_t k = 0; k < TS_H; k++) axpy32(acc, s_wd1[k], s_h[k][tid]);
#pragma unroll
            for (uint32_t j = 0; j < TS_H; j++) s_h[j][tid] = fmaxf(rh(acc[j]), 0.0f);       // own column: no barrier needed
            // ---- layer 2 -> dx (fp16)
            float d0 = 0.0f, d1 = 0.0f;
#pragma unroll 4
            for (uint32_t k = 0; k < TS_H; k++) { const float v = s_h[k][tid]; d0 = __fmaf_rn(s_wd2[k][0], v, d0); d1 = __fmaf_rn(s_wd2[k][1], v, d1); }
            dx0 = rh(d0); dx1 = rh(d1);
            // x = (x + dx).clamp(-1, 1); GridEncoder: (x + bound) / (2 bound), bound = 1   (network.py:194, grid.py:143)
            const float u0 = __fdiv_rn(__fadd_rn(fminf(fmaxf(__fadd_rn(x0, dx0), -1.0f), 1.0f), 1.0f), 2.0f);
            const float u1 = __fdiv_rn(__fadd_rn(fminf(fmaxf(__fadd_rn(x1, dx1), -1.0f), 1.0f), 1.0f), 2.0f);
            // ---- torso net layer 0: grid part (features produced level by level), enc_x part, constant bias
#pragma unroll
            for (uint32_t j = 0; j < TS_H; j++) acc[j] = 0.0f;
#pragma unroll 1
            for (uint32_t l0 = 0; l0 < TS_LEVELS; l0 += 4) {                 // four levels = 16 table reads in flight before the first is used
                float2 t[4][4];
                float wgt[4][4];
#pragma unroll
                for (uint32_t q = 0; q < 4; q++) {
                    const TorsoLvl g = s_lvl[l0 + q];
                    float p0 = __fmaf_rn(u0, g.scale, 0.5f), p1 = __fmaf_rn(u1, g.scale, 0.5f);
                    const uint32_t i0 = (uint32_t)floorf(p0), i1 = (uint32_t)floorf(p1);
                    p0 = __fsub_rn(p0, (float)i0); p1 = __fsub_rn(p1, (float)i1);
                    const float2 *tab = reinterpret_cast<const float2 *>(a.table) + g.off;
#pragma unroll
                    for (uint32_t idx = 0; idx < 4; idx++) {
                        // w = 1 * (1 - p0 | p0) * (1 - p1 | p1) in this order (gridencoder.cu:150-160); tiled index, gridencoder.cu:54-72 with gridtype = tiled
                        wgt[q][idx] = __fmul_rn((idx & 1u) ? p0 : __fsub_rn(1.0f, p0), (idx & 2u) ? p1 : __fsub_rn(1.0f, p1));
                        const uint32_t c0 = i0 + (idx & 1u), c1 = i1 + (idx >> 1);
                        uint32_t e = c0;
                        if (g.stride <= g.size) e += c1 * g.stride;
                        e = g.mask ? (e & g.mask) : (e % g.size);            // index % hashmap_size (uniform branch)
                        t[q][idx] = __ldg(tab + e);
                    }
                }
#pragma unroll
                for (uint32_t q = 0; q < 4; q++) {
                    float r0 = 0.0f, r1 = 0.0f;
#pragma unroll
                    for (uint32_t idx = 0; idx < 4; idx++) {
                        // scalar_t = half: product narrowed, sum narrowed (c10::Half operator+=)
                        r0 = rh(__fadd_rn(r0, rh(__fmul_rn(wgt[q][idx], rh(t[q][idx].x)))));
                        r1 = rh(__fadd_rn(r1, rh(__fmul_rn(wgt[q][idx], rh(t[q][idx].y)))));
                    }
                    axpy32(acc, s_wt0[2 * (l0 + q)], r0);
                    axpy32(acc, s_wt0[2 * (l0 + q) + 1], r1);
                }
            }
#pragma unroll 1
            for (uint32_t k = 0; k < TS_ENC; k++) axpy32(acc, s_wt0[TS_GRID + k], rh(freq2(x0, x1, k)));
#pragma unroll
            for (uint32_t j = 0; j < TS_H; j++) s_h[j][tid] = fmaxf(rh(__fadd_rn(acc[j], s_bias_t[j])), 0.0f);
            // ---- layer 1
#pragma unroll
            for (uint32_t j = 0; j < TS_H; j++) acc[j] = 0.0f;
#pragma unroll 1
            for (uint32_t k = 0; k < TS_H; k++) axpy32(acc, s_wt1[k], s_h[k][tid]);
#pragma unroll
            for (uint32_t j = 0; j < TS_H; j++) s_h[j][tid] = fmaxf(rh(acc[j]), 0.0f);
            // ---- layer 2 -> [alpha, r, g, b] logits (fp16), sigmoid * 1.002 - 0.001 on half tensors (network.py:203-204)
            float o[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 4
            for (uint32_t k = 0; k < TS_H; k++) {
                const float v = s_h[k][tid];
                const float4 c = *reinterpret_cast<const float4 *>(s_wt2[k]);
                o[0] = __fmaf_rn(c.x, v, o[0]); o[1] = __fmaf_rn(c.y, v, o[1]); o[2] = __fmaf_rn(c.z, v, o[2]); o[3] = __fmaf_rn(c.w, v, o[3]);
            }
            if (on) {
#pragma unroll
                for (uint32_t q = 0; q < 4; q++) {
                    const float s = rh(1.0f / (1.0f + expf(-rh(o[q]))));
                    const float v = rh(rh(s * 1.002f) - 0.001f);
                    if (q == 0) alpha = v; else col[q - 1] = v;
                }
            } else { dx0 = 0.0f; dx1 = 0.0f; }
        }
        if (live) {
            // bg_color = torso_color * torso_alpha + bg_color * (1 - torso_alpha), fp32   (renderer.py:620)
            const float om = __fsub_rn(1.0f, alpha);
#pragma unroll
            for (uint32_t q = 0; q < 3; q++) {
                const float bg = a.bg_color ? __ldg(a.bg_color + (a.bg_per_ray ? 3 * (size_t)n + q : q)) : 1.0f;
                __stcs(a.bg_out + 3 * (size_t)n + q, __fadd_rn(__fmul_rn(col[q], alpha), __fmul_rn(bg, om)));
            }
            if (a.alpha_out) __stcs(a.alpha_out + n, alpha);
            if (a.deform_out) { __stcs(a.deform_out + 2 * (size_t)n, dx0); __stcs(a.deform_out + 2 * (size_t)n + 1, dx1); }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------------------
// training backward of the same branch (network.py:170-205 + renderer.py:572-631 through autograd), SURVEY §8f-2
// ---------------------------------------------------------------------------------------------------------------------------------------------
// One pass per pixel: the forward is recomputed with k_torso_frame's arithmetic (nothing was kept), then
//     d out -> d (alpha, colour) -> sigmoid -> torso_net (3 layers, ReLU masks held as four 32-bit words) -> d grid features
//     -> table gradients (fp32 vector reductions straight into the parameter's gradient; the reference accumulates them in half) and the gradient of the
//        deformed coordinate (gridencoder.cu:179-222's dy_dx) -> clamp mask -> torso_deform_net (3 layers).
// The weight gradients dW = dY^T X are not reduced here: the kernel writes every layer's X and dY rows as fp16 operands (zero rows for pixels outside the
// torso mask), and the tcgen05 weight-gradient kernel of the head (b2n_linear_wgrad_batch, csrc/wgrad.cu) multiplies all six pairs in one launch.
// The 50 per-frame constant inputs (anchor encoding, individual code) get their gradient from the column sums of two of those dY (host side: a matvec).
struct TorsoBwdArgs {
    TorsoArgs f;                // forward arguments (bg_out / alpha_out / deform_out unused)
    const float *g_out;         // [N,3] gradient of bg_out (= torso_color)
    const float *g_alpha;       // [N] gradient of torso_alpha or NULL
    float *g_table;             // [sO,2] fp32, accumulated
    __half *x_t0, *x_t1, *x_t2, *x_d0, *x_d1, *x_d2;        // [N,120] [N,32] [N,32] [N,88] [N,32] [N,32]
    __half *dy_t2, *dy_t1, *dy_t0, *dy_d2, *dy_d1, *dy_d0;  // [N,8]  [N,32] [N,32] [N,8]  [N,32] [N,32]
};

// dot of one transposed weight row (32 outputs of input k, shared memory, broadcast) with a 32-vector in registers
__device__ __forceinline__ float dot32(const float *w, const float (&v)[TS_H]) {
    float r = 0.0f;
#pragma unroll
    for (uint32_t q = 0; q < TS_H / 4; q++) {
        const float4 c = reinterpret_cast<const float4 *>(w)[q];
        r = __fmaf_rn(c.x, v[4 * q], r); r = __fmaf_rn(c.y, v[4 * q + 1], r); r = __fmaf_rn(c.z, v[4 * q + 2], r); r = __fmaf_rn(c.w, v[4 * q + 3], r);
    }
    return r;
}
// a thread's row of `n` halves (n a multiple of 8) out of its shared-memory column
__device__ __forceinline__ void row_out(__half *dst, const float (*col)[TS_THREADS], uint32_t tid, uint32_t n) {
    for (uint32_t c = 0; c < n; c += 8) {
        __half2 h[4];
#pragma unroll
        for (uint32_t q = 0; q < 4; q++) h[q] = __floats2half2_rn(col[c + 2 * q][tid], col[c + 2 * q + 1][tid]);
        *reinterpret_cast<uint4 *>(dst + c) = *reinterpret_cast<const uint4 *>(h);
    }
}
__device__ __forceinline__ void row_zero(__half *dst, uint32_t n) {
    for (uint32_t c = 0; c < n; c += 8) *reinterpret_cast<uint4 *>(dst + c) = make_uint4(0u, 0u, 0u, 0u);
}

__global__ void __launch_bounds__(TS_THREADS) k_torso_backward(const __grid_constant__ TorsoBwdArgs b) {
    const TorsoArgs &a = b.f;
    __shared__ __align__(16) float s_img[TI_MAIN];
    float (*s_wd0)[TS_H] = reinterpret_cast<float (*)[TS_H]>(s_img + TI_WD0);
    float (*s_wd1)[TS_H] = reinterpret_cast<float (*)[TS_H]>(s_img + TI_WD1);
    float (*s_wt0)[TS_H] = reinterpret_cast<float (*)[TS_H]>(s_img + TI_WT0);
    float (*s_wt1)[TS_H] = reinterpret_cast<float (*)[TS_H]>(s_img + TI_WT1);
    float (*s_wd2)[2] = reinterpret_cast<float (*)[2]>(s_img + TI_WD2);
    float (*s_wt2)[4] = reinterpret_cast<float (*)[4]>(s_img + TI_WT2);
    __shared__ float s_bias_d[TS_H], s_bias_t[TS_H], s_hc[TS_CONST];
    __shared__ TorsoLvl s_lvl[TS_LEVELS];
    extern __shared__ __align__(16) float tb_dyn[];      // (beyond the 48 KB of static shared memory)
    float (*s_h)[TS_THREADS] = reinterpret_cast<float (*)[TS_THREADS]>(tb_dyn);                        // one 32-vector per thread (column = thread)
    float (*s_e)[TS_THREADS] = reinterpret_cast<float (*)[TS_THREADS]>(tb_dyn + TS_H * TS_THREADS);    // the pixel's frequency encoding [34]
    const uint32_t tid = threadIdx.x;
    for (uint32_t i = tid; i < TI_MAIN / 4; i += TS_THREADS) reinterpret_cast<float4 *>(s_img)[i] = __ldg(reinterpret_cast<const float4 *>(a.img) + i);
    if (tid < 2 * TS_H) {
        const float *cw = a.img + (tid < TS_H ? TI_CD : TI_CT) + (tid & (TS_H - 1u));
        float bb = 0.0f;
#pragma unroll 10
        for (uint32_t c = 0; c < TS_CONST; c++) bb = __fmaf_rn(__ldg(cw + c * TS_H), rh(__ldg(a.hconst + c)), bb);
        (tid < TS_H ? s_bias_d : s_bias_t)[tid & (TS_H - 1u)] = bb;
    }
    if (tid >= 2 * TS_H && tid < 2 * TS_H + TS_LEVELS) {
        const uint32_t l = tid - 2 * TS_H;
        TorsoLvl g;
        g.off = (uint32_t)a.offsets[l];
        g.size = (uint32_t)a.offsets[l + 1] - g.off;
        g.scale = __fmaf_rn(exp2f(__fmul_rn((float)l, a.S)), (float)a.H, -1.0f);
        g.stride = (uint32_t)ceilf(g.scale) + 2u;
        g.mask = (g.stride <= g.size && (uint64_t)g.stride * g.stride <= g.size) ? 0xffffffffu : ((g.size & (g.size - 1u)) == 0u ? g.size - 1u : 0u);
        s_lvl[l] = g;
    }
    for (uint32_t c = tid; c < TS_CONST; c += TS_THREADS) s_hc[c] = rh(__ldg(a.hconst + c));
    __syncthreads();

    const float Gm1 = (float)(a.G - 1u);
    __shared__ uint32_t s_tile;
    const uint32_t n_tiles = (a.N + TS_THREADS - 1) / TS_THREADS;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(a.tile_counter, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= n_tiles) break;
        const uint32_t n = tile * TS_THREADS + tid;
        const bool live = n < a.N;
        float cx = 0.0f, cy = 0.0f;
        if (live) { cx = __ldcs(a.bg_coords + 2 * (size_t)n); cy = __ldcs(a.bg_coords + 2 * (size_t)n + 1); }
        bool on = false;
        if (live) {                                        // occupancy test: k_torso_frame's
            const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(cx, 1.0f), 2.0f), Gm1), iy = __fmul_rn(__fdiv_rn(__fadd_rn(cy, 1.0f), 2.0f), Gm1);
            const float fx = floorf(ix), fy = floorf(iy);
            const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
            const float nw = __fmul_rn(__fsub_rn(fx + 1.0f, ix), __fsub_rn(fy + 1.0f, iy)), ne = __fmul_rn(__fsub_rn(ix, fx), __fsub_rn(fy + 1.0f, iy));
            const float sw = __fmul_rn(__fsub_rn(fx + 1.0f, ix), __fsub_rn(iy, fy)), se = __fmul_rn(__fsub_rn(ix, fx), __fsub_rn(iy, fy));
            const int G = (int)a.G;
            auto tap = [&](int x, int y) { return (x >= 0 && x < G && y >= 0 && y < G) ? __ldg(a.dgrid + (size_t)y * G + x) : 0.0f; };
            float occ = 0.0f;
            occ = __fmaf_rn(tap(x0, y0), nw, occ); occ = __fmaf_rn(tap(x1, y0), ne, occ);
            occ = __fmaf_rn(tap(x0, y1), sw, occ); occ = __fmaf_rn(tap(x1, y1), se, occ);
            on = occ > a.thresh;
        }
        if (!__any_sync(0xffffffffu, on)) {                 // no torso pixel in this warp: zero operand rows, nothing else
            if (live) {
                row_zero(b.x_t0 + (size_t)n * 120, 120); row_zero(b.x_t1 + (size_t)n * 32, 32); row_zero(b.x_t2 + (size_t)n * 32, 32);
                row_zero(b.x_d0 + (size_t)n * 88, 88); row_zero(b.x_d1 + (size_t)n * 32, 32); row_zero(b.x_d2 + (size_t)n * 32, 32);
                row_zero(b.dy_t2 + (size_t)n * 8, 8); row_zero(b.dy_t1 + (size_t)n * 32, 32); row_zero(b.dy_t0 + (size_t)n * 32, 32);
                row_zero(b.dy_d2 + (size_t)n * 8, 8); row_zero(b.dy_d1 + (size_t)n * 32, 32); row_zero(b.dy_d0 + (size_t)n * 32, 32);
            }
            continue;
        }
        const size_t row = live ? (size_t)n : 0;            // (dead lanes of a live warp compute on pixel 0 and write nothing)
        // =============================== forward (k_torso_frame's arithmetic) ===============================
        const float x0 = __fmul_rn(cx, a.shrink), x1 = __fmul_rn(cy, a.shrink);
        for (uint32_t k = 0; k < TS_ENC; k++) s_e[k][tid] = rh(freq2(x0, x1, k));
        float acc[TS_H];
        uint32_t m_d1 = 0, m_d2 = 0, m_t1 = 0, m_t2 = 0;    // ReLU masks
#pragma unroll
        for (uint32_t j = 0; j < TS_H; j++) acc[j] = 0.0f;
#pragma unroll 1
        for (uint32_t k = 0; k < TS_ENC; k++) axpy32(acc, s_wd0[k], s_e[k][tid]);
#pragma unroll
        for (uint32_t j = 0; j < TS_H; j++) { const float v = fmaxf(rh(__fadd_rn(acc[j], s_bias_d[j])), 0.0f); s_h[j][tid] = v; m_d1 |= (v > 0.0f ? 1u : 0u) << j; }
        if (live) row_out(b.x_d1 + row * 32, s_h, tid, 32);
#pragma unroll
        for (uint32_t j = 0; j < TS_H; j++) acc[j] = 0.0f;
#pragma unroll 1
        for (uint32_t k = 0; k < TS_H; k++) axpy32(acc, s_wd1[k], s_h[k][tid]);
#pragma unroll
        for (uint32_t j = 0; j < TS_H; j++) { const float v = fmaxf(rh(acc[j]), 0.0f); s_h[j][tid] = v; m_d2 |= (v > 0.0f ? 1u : 0u) << j; }
        if (live) row_out(b.x_d2 + row * 32, s_h, tid, 32);
        float d0 = 0.0f, d1 = 0.0f;
#pragma unroll 4
        for (uint32_t k = 0; k < TS_H; k++) { const float v = s_h[k][tid]; d0 = __fmaf_rn(s_wd2[k][0], v, d0); d1 = __fmaf_rn(s_wd2[k][1], v, d1); }
        const float dx0 = rh(d0), dx1 = rh(d1);
        const float xs0 = __fadd_rn(x0, dx0), xs1 = __fadd_rn(x1, dx1);
        const bool in0 = xs0 >= -1.0f && xs0 <= 1.0f, in1 = xs1 >= -1.0f && xs1 <= 1.0f;       // clamp passes the gradient inside [-1, 1]
        const float u0 = __fdiv_rn(__fadd_rn(fminf(fmaxf(xs0, -1.0f), 1.0f), 1.0f), 2.0f);
        const float u1 = __fdiv_rn(__fadd_rn(fminf(fmaxf(xs1, -1.0f), 1.0f), 1.0f), 2.0f);
#pragma unroll
        for (uint32_t j = 0; j < TS_H; j++) acc[j] = 0.0f;
#pragma unroll 1
        for (uint32_t l = 0; l < TS_LEVELS; l++) {           // grid features, one level at a time; kept in s_h for the X operand
            const TorsoLvl g = s_lvl[l];
            float p0 = __fmaf_rn(u0, g.scale, 0.5f), p1 = __fmaf_rn(u1, g.scale, 0.5f);
            const uint32_t i0 = (uint32_t)floorf(p0), i1 = (uint32_t)floorf(p1);
            p0 = __fsub_rn(p0, (float)i0); p1 = __fsub_rn(p1, (float)i1);
            const float2 *tab = reinterpret_cast<const float2 *>(a.table) + g.off;
            float r0 = 0.0f, r1 = 0.0f;
#pragma unroll
            for (uint32_t idx = 0; idx < 4; idx++) {
                const float w = __fmul_rn((idx & 1u) ? p0 : __fsub_rn(1.0f, p0), (idx & 2u) ? p1 : __fsub_rn(1.0f, p1));
                uint32_t e = i0 + (idx & 1u);
                if (g.stride <= g.size) e += (i1 + (idx >> 1)) * g.stride;
                e = g.mask ? (e & g.mask) : (e % g.size);
                const float2 t = __ldg(tab + e);
                r0 = rh(__fadd_rn(r0, rh(__fmul_rn(w, rh(t.x)))));
                r1 = rh(__fadd_rn(r1, rh(__fmul_rn(w, rh(t.y)))));
            }
            s_h[2 * l][tid] = r0; s_h[2 * l + 1][tid] = r1;
            axpy32(acc, s_wt0[2 * l], r0);
            axpy32(acc, s_wt0[2 * l + 1], r1);
        }
        if (live) {                                          // X operands of both first layers: [g 32 | enc_x 34 | const 50 | 0 x4] and [enc_x 34 | const 50 | 0 x4]
            __half *xt = b.x_t0 + row * 120, *xd = b.x_d0 + row * 88;
            row_out(xt, s_h, tid, 32);
            for (uint32_t k = 0; k < 88; k += 2) {
                const float v0 = k < TS_ENC ? s_e[k][tid] : (k < TS_ENC + TS_CONST ? s_hc[k - TS_ENC] : 0.0f);
                const float v1 = k + 1 < TS_ENC ? s_e[k + 1][tid] : (k + 1 < TS_ENC + TS_CONST ? s_hc[k + 1 - TS_ENC] : 0.0f);
                const __half2 h = __floats2half2_rn(v0, v1);
                *reinterpret_cast<__half2 *>(xt + 32 + k) = h;
                *reinterpret_cast<__half2 *>(xd + k) = h;
            }
        }
#pragma unroll 1
        for (uint32_t k = 0; k < TS_ENC; k++) axpy32(acc, s_wt0[TS_GRID + k], s_e[k][tid]);
#pragma unroll
        for (uint32_t j = 0; j < TS_H; j++) { const float v = fmaxf(rh(__fadd_rn(acc[j], s_bias_t[j])), 0.0f); s_h[j][tid] = v; m_t1 |= (v > 0.0f ? 1u : 0u) << j; }
        if (live) row_out(b.x_t1 + row * 32, s_h, tid, 32);
#pragma unroll
        for (uint32_t j = 0; j < TS_H; j++) acc[j] = 0.0f;
#pragma unroll 1
        for (uint32_t k = 0; k < TS_H; k++) axpy32(acc, s_wt1[k], s_h[k][tid]);
#pragma unroll
        for (uint32_t j = 0; j < TS_H; j++) { const float v = fmaxf(rh(acc[j]), 0.0f); s_h[j][tid] = v; m_t2 |= (v > 0.0f ? 1u : 0u) << j; }
        if (live) row_out(b.x_t2 + row * 32, s_h, tid, 32);
        float o[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 4
        for (uint32_t k = 0; k < TS_H; k++) {
            const float v = s_h[k][tid];
            const float4 c = *reinterpret_cast<const float4 *>(s_wt2[k]);
            o[0] = __fmaf_rn(c.x, v, o[0]); o[1] = __fmaf_rn(c.y, v, o[1]); o[2] = __fmaf_rn(c.z, v, o[2]); o[3] = __fmaf_rn(c.w, v, o[3]);
        }
        // =============================== backward ===============================
        float d_o[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (on && live) {
            float sg[4], val[4];
#pragma unroll
            for (uint32_t q = 0; q < 4; q++) { sg[q] = rh(1.0f / (1.0f + expf(-rh(o[q])))); val[q] = rh(rh(sg[q] * 1.002f) - 0.001f); }
            const float alpha = val[0];
            float d_alpha = b.g_alpha ? __ldg(b.g_alpha + n) : 0.0f;
#pragma unroll
            for (uint32_t q = 0; q < 3; q++) {
                const float go = __ldg(b.g_out + 3 * (size_t)n + q);
                const float bg = a.bg_color ? __ldg(a.bg_color + (a.bg_per_ray ? 3 * (size_t)n + q : q)) : 1.0f;
                d_alpha = __fmaf_rn(go, val[q + 1] - bg, d_alpha);                 // out = col * alpha + bg * (1 - alpha)
                d_o[q + 1] = go * alpha * 1.002f * sg[q + 1] * (1.0f - sg[q + 1]);
            }
            d_o[0] = d_alpha * 1.002f * sg[0] * (1.0f - sg[0]);
        }
        if (live) {
            const __half2 h01 = __floats2half2_rn(d_o[0], d_o[1]), h23 = __floats2half2_rn(d_o[2], d_o[3]);
            *reinterpret_cast<uint4 *>(b.dy_t2 + row * 8) = make_uint4(*reinterpret_cast<const uint32_t *>(&h01), *reinterpret_cast<const uint32_t *>(&h23), 0u, 0u);
        }
        float dv[TS_H];
        // torso_net layer 2 -> d (layer-1 pre-activation)
#pragma unroll
        for (uint32_t k = 0; k < TS_H; k++) {
            const float4 c = *reinterpret_cast<const float4 *>(s_wt2[k]);
            const float v = c.x * d_o[0] + c.y * d_o[1] + c.z * d_o[2] + c.w * d_o[3];
            dv[k] = ((m_t2 >> k) & 1u) ? v : 0.0f;
        }
#pragma unroll
        for (uint32_t k = 0; k < TS_H; k++) s_h[k][tid] = dv[k];
        if (live) row_out(b.dy_t1 + row * 32, s_h, tid, 32);
        // layer 1 -> d (layer-0 pre-activation)
#pragma unroll 1
        for (uint32_t k = 0; k < TS_H; k++) s_h[k][tid] = ((m_t1 >> k) & 1u) ? dot32(s_wt1[k], dv) : 0.0f;
#pragma unroll
        for (uint32_t k = 0; k < TS_H; k++) dv[k] = s_h[k][tid];
        if (live) row_out(b.dy_t0 + row * 32, s_h, tid, 32);
        // layer 0 -> d grid features (inputs 0..31); table gradients and the gradient of the deformed coordinate, level by level
        float g_u0 = 0.0f, g_u1 = 0.0f;
#pragma unroll 1
        for (uint32_t l = 0; l < TS_LEVELS; l++) {
            const float dg0 = dot32(s_wt0[2 * l], dv), dg1 = dot32(s_wt0[2 * l + 1], dv);
            const TorsoLvl g = s_lvl[l];
            float p0 = __fmaf_rn(u0, g.scale, 0.5f), p1 = __fmaf_rn(u1, g.scale, 0.5f);
            const uint32_t i0 = (uint32_t)floorf(p0), i1 = (uint32_t)floorf(p1);
            p0 = __fsub_rn(p0, (float)i0); p1 = __fsub_rn(p1, (float)i1);
            const float2 *tab = reinterpret_cast<const float2 *>(a.table) + g.off;
            float2 t[4];
            uint32_t e4[4];
#pragma unroll
            for (uint32_t idx = 0; idx < 4; idx++) {
                uint32_t e = i0 + (idx & 1u);
                if (g.stride <= g.size) e += (i1 + (idx >> 1)) * g.stride;
                e = g.mask ? (e & g.mask) : (e % g.size);
                e4[idx] = e;
                t[idx] = __ldg(tab + e);
            }
            if (on && live) {
#pragma unroll
                for (uint32_t idx = 0; idx < 4; idx++) {
                    const float w = ((idx & 1u) ? p0 : 1.0f - p0) * ((idx & 2u) ? p1 : 1.0f - p1);
                    float *dst = b.g_table + 2 * (size_t)(g.off + e4[idx]);
                    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(w * dg0), "f"(w * dg1) : "memory");
                }
                // d feature / d u_d = scale * sum over the other dimension's corners of w_other * (right - left)   (gridencoder.cu:179-222)
                const float dfx0 = (1.0f - p1) * (rh(t[1].x) - rh(t[0].x)) + p1 * (rh(t[3].x) - rh(t[2].x));
                const float dfx1 = (1.0f - p1) * (rh(t[1].y) - rh(t[0].y)) + p1 * (rh(t[3].y) - rh(t[2].y));
                const float dfy0 = (1.0f - p0) * (rh(t[2].x) - rh(t[0].x)) + p0 * (rh(t[3].x) - rh(t[1].x));
                const float dfy1 = (1.0f - p0) * (rh(t[2].y) - rh(t[0].y)) + p0 * (rh(t[3].y) - rh(t[1].y));
                g_u0 = __fmaf_rn(g.scale, dg0 * dfx0 + dg1 * dfx1, g_u0);
                g_u1 = __fmaf_rn(g.scale, dg0 * dfy0 + dg1 * dfy1, g_u1);
            }
        }
        // u = (clamp(x + dx) + 1) / 2  ->  d dx = d u / 2 inside the clamp range
        const float g_dx0 = (on && in0) ? 0.5f * g_u0 : 0.0f, g_dx1 = (on && in1) ? 0.5f * g_u1 : 0.0f;
        if (live) {
            const __half2 h = __floats2half2_rn(g_dx0, g_dx1);
            *reinterpret_cast<uint4 *>(b.dy_d2 + row * 8) = make_uint4(*reinterpret_cast<const uint32_t *>(&h), 0u, 0u, 0u);
        }
        // torso_deform_net layer 2 -> layer 1 -> layer 0 pre-activations
#pragma unroll
        for (uint32_t k = 0; k < TS_H; k++) dv[k] = ((m_d2 >> k) & 1u) ? (s_wd2[k][0] * g_dx0 + s_wd2[k][1] * g_dx1) : 0.0f;
#pragma unroll
        for (uint32_t k = 0; k < TS_H; k++) s_h[k][tid] = dv[k];
        if (live) row_out(b.dy_d1 + row * 32, s_h, tid, 32);
#pragma unroll 1
        for (uint32_t k = 0; k < TS_H; k++) s_h[k][tid] = ((m_d1 >> k) & 1u) ? dot32(s_wd1[k], dv) : 0.0f;
        if (live) row_out(b.dy_d0 + row * 32, s_h, tid, 32);
    }
}

}  // namespace b2n

using namespace b2n;

extern "C" uint64_t b2n_torso_workspace_bytes(void) { return sizeof(float) * (TI_TOTAL + 4); }

extern "C" int b2n_torso_forward(const b2n_torso_weights *w, const float *bg_coords, uint32_t N, const float *density_grid_torso, uint32_t grid_size,
                                 float density_thresh, const float *h_const, const float *bg_color, int bg_per_ray, float *bg_out, float *alpha_out,
                                 float *deform_out, void *workspace, void *stream) {
    B2N_REQUIRE(w && bg_coords && density_grid_torso && h_const && bg_out, "torso_forward: null pointer");
    B2N_REQUIRE(w->deform_w0 && w->deform_w1 && w->deform_w2 && w->torso_w0 && w->torso_w1 && w->torso_w2 && w->table && w->offsets, "torso_forward: null weight pointer");
    B2N_REQUIRE(grid_size >= 2, "torso_forward: grid_size=%u", grid_size);
    B2N_REQUIRE(((uintptr_t)w->table & 7) == 0, "torso_forward: table must be 8-byte aligned");
    if (N == 0) return 0;
    TorsoArgs a = {};
    a.bg_coords = bg_coords; a.N = N; a.dgrid = density_grid_torso; a.G = grid_size; a.thresh = density_thresh; a.shrink = w->torso_shrink;
    a.hconst = h_const; a.table = w->table; a.offsets = w->offsets; a.S = w->S; a.H = w->H;
    a.bg_color = bg_color; a.bg_per_ray = bg_per_ray; a.bg_out = bg_out; a.alpha_out = alpha_out; a.deform_out = deform_out;
    // the weights are re-packed on every call (a 35 KB image + the tile counter in the caller's workspace, one small launch): the caller may have
    // stepped an optimizer in between, and concurrent frames on different streams each bring their own workspace
    B2N_REQUIRE(workspace && ((uintptr_t)workspace & 15) == 0, "torso_forward: workspace must be a 16-byte aligned device buffer of b2n_torso_workspace_bytes()");
    float *img = static_cast<float *>(workspace);
    k_torso_pack<<<ceil_div<uint32_t>(TI_TOTAL, 256), 256, 0, as_stream(stream)>>>(w->deform_w0, w->deform_w1, w->deform_w2, w->torso_w0, w->torso_w1, w->torso_w2, img, reinterpret_cast<uint32_t *>(img + TI_TOTAL));
    if (check_launch("torso_forward(pack)")) return 1;
    a.img = img; a.tile_counter = reinterpret_cast<uint32_t *>(img + TI_TOTAL);
    const uint32_t tiles = ceil_div<uint32_t>(N, TS_THREADS);
    uint32_t ctas = 4u * (uint32_t)sm_count();            // 39 KB of shared memory per CTA: four resident CTAs per SM
    if (ctas > tiles) ctas = tiles;
    k_torso_frame<<<ctas, TS_THREADS, 0, as_stream(stream)>>>(a);
    return check_launch("torso_forward");
}

// Backward of b2n_torso_forward for training.  g_out [N,3] = gradient of bg_out, g_alpha [N] or NULL = gradient of alpha_out.  Produces
//   * g_table [sO,2] fp32: table gradients ACCUMULATED into (zero it first, or pass the parameter's .grad);
//   * the fp16 operands of the six weight gradients dW = dY^T X (every row written; zero rows outside the torso mask), to be multiplied by
//     b2n_linear_wgrad_batch:  torso_net.2: dy_t2 [N,8] x_t2 [N,32];  torso_net.1: dy_t1 x_t1 [N,32];  torso_net.0: dy_t0 [N,32] x_t0 [N,120] =
//     [grid 32 | enc_x 34 | const 50 | 0 x4];  torso_deform_net.2: dy_d2 [N,8] x_d2;  .1: dy_d1 x_d1;  .0: dy_d0 x_d0 [N,88] = [enc_x 34 | const 50 | 0 x4].
//   The gradient of h_const is W_t0[:, 66:116]^T colsum(dy_t0) + W_d0[:, 34:84]^T colsum(dy_d0) (host side).
extern "C" int b2n_torso_backward(const b2n_torso_weights *w, const float *bg_coords, uint32_t N, const float *density_grid_torso, uint32_t grid_size,
                                  float density_thresh, const float *h_const, const float *bg_color, int bg_per_ray, const float *g_out, const float *g_alpha,
                                  float *g_table, const b2n_torso_operands *ops, void *workspace, void *stream) {
    B2N_REQUIRE(w && bg_coords && density_grid_torso && h_const && g_out && g_table && ops, "torso_backward: null pointer");
    B2N_REQUIRE(w->deform_w0 && w->deform_w1 && w->deform_w2 && w->torso_w0 && w->torso_w1 && w->torso_w2 && w->table && w->offsets, "torso_backward: null weight pointer");
    B2N_REQUIRE(ops->x_t0 && ops->x_t1 && ops->x_t2 && ops->x_d0 && ops->x_d1 && ops->x_d2 && ops->dy_t2 && ops->dy_t1 && ops->dy_t0 && ops->dy_d2 && ops->dy_d1 && ops->dy_d0,
                "torso_backward: null operand buffer");
    B2N_REQUIRE(grid_size >= 2, "torso_backward: grid_size=%u", grid_size);
    B2N_REQUIRE(((uintptr_t)w->table & 7) == 0 && ((uintptr_t)g_table & 7) == 0, "torso_backward: tables must be 8-byte aligned");
    B2N_REQUIRE(workspace && ((uintptr_t)workspace & 15) == 0, "torso_backward: workspace must be a 16-byte aligned device buffer of b2n_torso_workspace_bytes()");
    if (N == 0) return 0;
    TorsoBwdArgs b = {};
    TorsoArgs &a = b.f;
    a.bg_coords = bg_coords; a.N = N; a.dgrid = density_grid_torso; a.G = grid_size; a.thresh = density_thresh; a.shrink = w->torso_shrink;
    a.hconst = h_const; a.table = w->table; a.offsets = w->offsets; a.S = w->S; a.H = w->H;
    a.bg_color = bg_color; a.bg_per_ray = bg_per_ray;
    float *img = static_cast<float *>(workspace);
    k_torso_pack<<<ceil_div<uint32_t>(TI_TOTAL, 256), 256, 0, as_stream(stream)>>>(w->deform_w0, w->deform_w1, w->deform_w2, w->torso_w0, w->torso_w1, w->torso_w2, img, reinterpret_cast<uint32_t *>(img + TI_TOTAL));
    if (check_launch("torso_backward(pack)")) return 1;
    a.img = img; a.tile_counter = reinterpret_cast<uint32_t *>(img + TI_TOTAL);
    b.g_out = g_out; b.g_alpha = g_alpha; b.g_table = g_table;
    b.x_t0 = (__half *)ops->x_t0; b.x_t1 = (__half *)ops->x_t1; b.x_t2 = (__half *)ops->x_t2; b.x_d0 = (__half *)ops->x_d0; b.x_d1 = (__half *)ops->x_d1; b.x_d2 = (__half *)ops->x_d2;
    b.dy_t2 = (__half *)ops->dy_t2; b.dy_t1 = (__half *)ops->dy_t1; b.dy_t0 = (__half *)ops->dy_t0; b.dy_d2 = (__half *)ops->dy_d2; b.dy_d1 = (__half *)ops->dy_d1; b.dy_d0 = (__half *)ops->dy_d0;
    const uint32_t tiles = ceil_div<uint32_t>(N, TS_THREADS);
    uint32_t ctas = 3u * (uint32_t)sm_count();
    if (ctas > tiles) ctas = tiles;
    const size_t dyn = sizeof(float) * (TS_H + TS_ENC + 2) * TS_THREADS;
    B2N_SMEM(k_torso_backward, dyn);
    k_torso_backward<<<ctas, TS_THREADS, dyn, as_stream(stream)>>>(b);
    return check_launch("torso_backward");
}
