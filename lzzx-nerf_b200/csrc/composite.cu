// composite.cu — alpha compositing (training forward/backward, inference incremental) for sm_100a.
//
// One templated core instantiates the reference's four training variants and five inference variants
// (raymarching.cu:604-2258): <AMB, NA, UNC> = <ambient mode, #ambient channels, has uncertainty>
//     plain        <1,1,false>   ambient_sum += a              (:604/:712/:1043)
//     sigma        <2,1,false>   ambient_sum += w * a          (:1162/:1270/:1387)
//     uncertainty  <1,1,true>    + uncertainty_sum += w * u    (:1507/:1622/:1754)
//     triplane     <1,2,true>    two plain ambient channels    (:1878/:2000/:2142)
//     rgb only     <0,0,false>   composite_rays                (:943)
//
// The per-ray recurrence (T *= 1-alpha, early stop on T < T_thresh) is kept SERIAL and in the reference's
// operation order, so results are bit-identical to the reference kernels on the same GPU (same MUFU.EX2).
// What changes is data movement: a CTA owns 128 consecutive rows of `rays`; when their sample segments tile
// one contiguous range (always true for b2n_march_rays_train's deterministic allocation) the CTA stages that
// range through shared memory with TMA bulk copies (cp.async.bulk: one thread puts up to 37 KB per pass in flight, six CTAs per SM keep
// > 200 KB outstanding per SM — the latency x bandwidth product HBM3e needs; tiles with more than 1024 samples take several passes),
// threads walk their ray out of shared memory, and the backward writes its per-sample gradients back through
// the same staging buffer with bulk stores — so HBM sees each algorithmic byte exactly once.  Foreign `rays` orderings (the reference's atomic allocation)
// are gathered into the same staging buffers sample-parallel per warp, forward and backward (the backward scatters its gradients back the same way).
#include <stdlib.h>
#include "common.cuh"
#include "tc5.cuh"

namespace b2n {

constexpr int CT_THREADS = 128;
constexpr int CT_CAP_DEFAULT = 1024;    // samples staged per PASS (runtime `cap`): 37 KB for the triplane variant -> six CTAs per SM.  A CTA whose 128 rays hold more
                                        // samples takes several passes over groups of consecutive rows (the reservation used to be the worst case, 128 rays x 16 steps
                                        // = 74 KB, three CTAs per SM, while the average tile needs half of it: ncu showed 12 latency-bound warps per SM)

constexpr int CT_PAD = 4;               // floats of slack per staged array (a span keeps its source's 16-byte phase)

template <int NA, bool UNC> struct Stage {
    // floats per staged sample: sigma 1, deltas 2, rgb 3, ambient NA, unc
    static constexpr int FLOATS = 6 + NA + (UNC ? 1 : 0);
    static size_t bytes(uint32_t cap) { return sizeof(float) * ((size_t)FLOATS * cap + 6 * CT_PAD); }
};
static uint32_t stage_cap() {           // B2N_COMP_CAP: measurement switch
    static const uint32_t cap = [] { const char *e = getenv("B2N_COMP_CAP"); const int v = e ? atoi(e) : 0; return (uint32_t)((v >= 8 && v <= 4096) ? (v & ~3) : CT_CAP_DEFAULT); }();
    return cap;
}

// exp(-sigma*delta) exactly as nvcc emits __expf(-s*d) for the reference: (s*d) * -log2(e) -> ex2.approx
__device__ __forceinline__ float alpha_of(float sigma, float delta) { return __fsub_rn(1.0f, __expf(-__fmul_rn(sigma, delta))); }

// CTA-wide: decide whether the valid segments of this CTA's rows tile [lo, lo+total) in row order.
// Returns true (uniformly) if so and total > 0; fills lo / total / slot (= samples of the CTA's earlier rows).
__device__ __forceinline__ bool cta_tiling(bool valid, uint32_t off, uint32_t num, uint32_t &lo, uint32_t &total, uint32_t &slot) {
    __shared__ uint32_t s_w[CT_THREADS / 32];
    __shared__ uint32_t s_lo;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t cnt = valid ? num : 0u;
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    if (lane == 31) s_w[warp] = inc;
    if (threadIdx.x == 0) s_lo = 0xffffffffu;
    __syncthreads();
    uint32_t before = inc - cnt, tot = 0;
#pragma unroll
    for (int w = 0; w < CT_THREADS / 32; w++) { if (w < (int)warp) before += s_w[w]; tot += s_w[w]; }
    // the first valid row defines lo (it is the one with before == 0 and cnt > 0)
    if (valid && before == 0) s_lo = off;        // several rows may have before==0 only if earlier ones have cnt==0 → invalid
    __syncthreads();
    lo = s_lo; total = tot; slot = before;
    const bool ok = !valid || (off == lo + before);
    return __syncthreads_and(ok) && tot > 0;
}

// The next group of consecutive rows whose samples [base, end) fit `cap` staged samples; returns end (uniform over the CTA).  end == base: the row that starts
// at `base` alone is longer than the staging capacity.
__device__ __forceinline__ uint32_t group_end(bool valid, uint32_t slot, uint32_t cnt, uint32_t base, uint32_t total, uint32_t cap) {
    __shared__ uint32_t s_end;
    if (threadIdx.x == 0) s_end = total;
    __syncthreads();
    if (valid && slot >= base && slot + cnt > base + cap) atomicMin(&s_end, slot);
    __syncthreads();
    const uint32_t e = s_end;
    __syncthreads();                           // (the next call resets s_end)
    return e;
}
// end of the single over-long row that starts at `base` (uniform)
__device__ __forceinline__ uint32_t long_row_end(bool mine, uint32_t slot, uint32_t cnt) {
    __shared__ uint32_t s_next;
    if (mine) s_next = slot + cnt;
    __syncthreads();
    const uint32_t e = s_next;
    __syncthreads();
    return e;
}

// A staged span: `count` floats of global memory mirrored in shared memory AT THE SAME 16-BYTE PHASE, so its interior moves as one TMA
// bulk copy (16 B granules) and at most 3 floats at either end go through ordinary loads / stores.
struct Span {
    float *s;                   // s[i] mirrors g[i]
    uint32_t head, body, count; // floats before the first 16 B boundary; bulk floats (multiple of 4); total
};
__device__ __forceinline__ uint32_t phase_of(const float *g) { return ((uint32_t)(uintptr_t)g >> 2) & 3u; }
__device__ __forceinline__ Span span_of(float *s_base, const float *g, uint32_t count) {
    const uint32_t phase = phase_of(g);
    Span sp;
    sp.s = s_base + phase;
    sp.count = count;
    sp.head = min(count, (4u - phase) & 3u);
    sp.body = (count - sp.head) & ~3u;
    return sp;
}
// global -> shared.  Every thread calls it; thread 0 issues the bulk copy (completion on `bar`, whose expect_tx already counts body * 4).
__device__ __forceinline__ void span_load(const Span &sp, const float *__restrict__ g, uint64_t *bar) {
    const uint32_t edge = sp.head + sp.body, t = threadIdx.x - 32u;
    if (threadIdx.x < sp.head) sp.s[threadIdx.x] = __ldcs(g + threadIdx.x);
    if (t < sp.count - edge) sp.s[edge + t] = __ldcs(g + edge + t);
    if (threadIdx.x == 0 && sp.body) tc5::bulk_g2s(sp.s + sp.head, g + sp.head, sp.body * 4u, bar);
}
// shared -> global (same phase required: phase_of(g) == phase of sp.s); thread 0 issues the bulk store
__device__ __forceinline__ void span_store(const Span &sp, float *__restrict__ g) {
    const uint32_t edge = sp.head + sp.body, t = threadIdx.x - 32u;
    if (threadIdx.x < sp.head) __stcs(g + threadIdx.x, sp.s[threadIdx.x]);
    if (t < sp.count - edge) __stcs(g + edge + t, sp.s[edge + t]);
    if (threadIdx.x == 0 && sp.body) tc5::bulk_s2g(g + sp.head, sp.s + sp.head, sp.body * 4u);
}
// out-of-phase destination: plain coalesced stores
__device__ __forceinline__ void stage_out(float *__restrict__ dst, const float *src, uint32_t count) {
#pragma unroll 4
    for (uint32_t i = threadIdx.x; i < count; i += CT_THREADS) __stcs(dst + i, src[i]);
}
__device__ __forceinline__ void span_store_any(const Span &sp, float *__restrict__ g) {      // uniform branch per CTA
    if (phase_of(g) == (uint32_t)(sp.s - (float *)((uintptr_t)sp.s & ~(uintptr_t)15))) span_store(sp, g); else stage_out(g, sp.s, sp.count);
}

// ---------------------------------------------------------------------------------------------------
// training forward
// ---------------------------------------------------------------------------------------------------
template <int AMB, int NA, bool UNC>
__device__ __forceinline__ void train_fwd_ray(const float *sg, const float *rgb, const float *a0p, const float *a1p, const float *up,
                                              const float *dl, uint32_t num, float T_thresh,
                                              float &ws, float &a0, float &a1, float &u, float &d, float &r, float &g, float &b) {
    float T = 1.0f;
    for (uint32_t k = 0; k < num; k++) {
        const float alpha = alpha_of(sg[k], dl[2 * k]);
        const float w = __fmul_rn(alpha, T);
        r = __fmaf_rn(w, rgb[3 * k], r);
        g = __fmaf_rn(w, rgb[3 * k + 1], g);
        b = __fmaf_rn(w, rgb[3 * k + 2], b);
        d = __fmaf_rn(w, dl[2 * k + 1], d);
        ws = __fadd_rn(ws, w);
        if (NA >= 1) a0 = (AMB == 2) ? __fmaf_rn(w, a0p[k], a0) : __fadd_rn(a0, a0p[k]);
        if (NA >= 2) a1 = (AMB == 2) ? __fmaf_rn(w, a1p[k], a1) : __fadd_rn(a1, a1p[k]);
        if (UNC) u = __fmaf_rn(w, up[k], u);
        T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
        if (T < T_thresh) break;
    }
}

template <int AMB, int NA, bool UNC>
__global__ void __launch_bounds__(CT_THREADS) k_comp_train_fwd(
        const float *__restrict__ sigmas, const float *__restrict__ rgbs, const float *__restrict__ amb0,
        const float *__restrict__ amb1, const float *__restrict__ unc, const float *__restrict__ deltas,
        const int32_t *__restrict__ rays, uint32_t M, uint32_t N, float T_thresh, uint32_t cap,
        float *__restrict__ weights_sum, float *__restrict__ amb0_sum, float *__restrict__ amb1_sum,
        float *__restrict__ unc_sum, float *__restrict__ depth, float *__restrict__ image) {
    extern __shared__ __align__(16) float sm[];
    const uint32_t n = blockIdx.x * CT_THREADS + threadIdx.x;
    uint32_t idx = 0, off = 0, num = 0;
    if (n < N) { idx = (uint32_t)rays[3 * (size_t)n]; off = (uint32_t)rays[3 * (size_t)n + 1]; num = (uint32_t)rays[3 * (size_t)n + 2]; }
    const bool valid = (n < N) && !(num == 0 || off + num > M);
    __shared__ __align__(8) uint64_t s_bar;
    if (threadIdx.x == 0) { tc5::mbar_init(&s_bar, 1); tc5::fence_mbar_init(); }
    uint32_t lo, total, slot;
    const bool tiled = cta_tiling(valid, off, num, lo, total, slot);        // (its barriers also publish the mbarrier init)
    float ws = 0, a0 = 0, a1 = 0, u = 0, d = 0, r = 0, g = 0, b = 0;
    float *b_sg = sm, *b_dl = b_sg + cap + CT_PAD, *b_rgb = b_dl + 2 * cap + CT_PAD, *b_a0 = b_rgb + 3 * cap + CT_PAD;
    float *b_a1 = b_a0 + (NA >= 1 ? cap + CT_PAD : 0), *b_u = b_a1 + (NA >= 2 ? cap + CT_PAD : 0);
    const uint32_t cnt = valid ? num : 0u;
    auto walk_global = [&]() {
        train_fwd_ray<AMB, NA, UNC>(sigmas + off, rgbs + 3 * (size_t)off, NA >= 1 ? amb0 + off : nullptr, NA >= 2 ? amb1 + off : nullptr,
                                    UNC ? unc + off : nullptr, deltas + 2 * (size_t)off, num, T_thresh, ws, a0, a1, u, d, r, g, b);
    };
    // groups of consecutive rows whose samples fit the staging buffers, one pass each (`slot` = samples of the CTA's earlier rows, in row order)
    uint32_t bar_phase = 0;
    for (uint32_t base = 0; base < total;) {
        const uint32_t end = group_end(valid, slot, cnt, base, total, cap);
        if (end == base) {                     // one ray longer than the staging capacity: straight out of global memory
            const bool mine = valid && slot == base;
            if (mine) walk_global();
            base = long_row_end(mine, slot, cnt);
            continue;
        }
        const bool in_group = valid && slot >= base && slot < end;
        const uint32_t gcnt = end - base;
        if (tiled) {
            // the rows' segments tile [lo, lo + total) in row order (b2n_march_rays_train's allocation): the group is ONE contiguous range per array -> TMA bulk copies
            const size_t g0 = (size_t)lo + base;
            const Span p_sg = span_of(b_sg, sigmas + g0, gcnt), p_dl = span_of(b_dl, deltas + 2 * g0, 2 * gcnt), p_rgb = span_of(b_rgb, rgbs + 3 * g0, 3 * gcnt);
            const Span p_a0 = NA >= 1 ? span_of(b_a0, amb0 + g0, gcnt) : Span{b_a0, 0, 0, 0};
            const Span p_a1 = NA >= 2 ? span_of(b_a1, amb1 + g0, gcnt) : Span{b_a1, 0, 0, 0};
            const Span p_u = UNC ? span_of(b_u, unc + g0, gcnt) : Span{b_u, 0, 0, 0};
            if (threadIdx.x == 0) tc5::mbar_expect_tx(&s_bar, 4u * (p_sg.body + p_dl.body + p_rgb.body + p_a0.body + p_a1.body + p_u.body));
            span_load(p_sg, sigmas + g0, &s_bar);
            span_load(p_dl, deltas + 2 * g0, &s_bar);
            span_load(p_rgb, rgbs + 3 * g0, &s_bar);
            if (NA >= 1) span_load(p_a0, amb0 + g0, &s_bar);
            if (NA >= 2) span_load(p_a1, amb1 + g0, &s_bar);
            if (UNC) span_load(p_u, unc + g0, &s_bar);
            __syncthreads();                       // edge floats
            tc5::mbar_wait(&s_bar, bar_phase);     // bulk interior
            bar_phase ^= 1u;
            if (in_group) {
                const uint32_t o = slot - base;
                train_fwd_ray<AMB, NA, UNC>(p_sg.s + o, p_rgb.s + 3 * o, p_a0.s + o, p_a1.s + o, p_u.s + o, p_dl.s + 2 * o, num, T_thresh, ws, a0, a1, u, d, r, g, b);
            }
        } else {
            // Foreign row order (the reference's atomic allocation, raymarching.cu:446-454: consecutive rows own unrelated segments).  A thread walking its ray
            // straight out of global memory touches one sector per sample and array; instead every warp gathers the samples of its rays of this group into the
            // staging arrays, at the slots the exclusive scan of the counts assigns, and the rays are then walked out of shared memory like in the tiled case.
            // Sample-parallel: lane = one staged sample of the warp's rays of this group; its ray is found by a binary search over the lanes' slots
            // (5 shuffles), then the nine floats of the sample are loaded — consecutive lanes read consecutive addresses inside a segment, every lane is
            // busy and all loads of a round are independent (copying ray by ray kept 8 of 32 lanes busy on one dependent load -> store chain per ray).
            const uint32_t lane = threadIdx.x & 31u;
            const uint32_t w_lo = max(__shfl_sync(0xffffffffu, slot, 0), base), w_hi = min(__shfl_sync(0xffffffffu, slot + cnt, 31), end);
#pragma unroll 1
            for (uint32_t i0 = w_lo; i0 < w_hi; i0 += 32) {
                const bool act = i0 + lane < w_hi;
                const uint32_t i = act ? i0 + lane : w_hi - 1;
                uint32_t j = 0;
#pragma unroll
                for (uint32_t step = 16; step; step >>= 1) {
                    const uint32_t sj = __shfl_sync(0xffffffffu, slot, (j + step) & 31u);
                    if (sj <= i) j += step;                      // slots ascend over the lanes; rows without samples share their successor's slot
                }
                const size_t src = (size_t)__shfl_sync(0xffffffffu, off, j) + (i - __shfl_sync(0xffffffffu, slot, j));
                if (act) {
                    const uint32_t dst = i - base;
                    const float v_sg = __ldcs(sigmas + src), v_d0 = __ldcs(deltas + 2 * src), v_d1 = __ldcs(deltas + 2 * src + 1);
                    const float v_r = __ldcs(rgbs + 3 * src), v_g = __ldcs(rgbs + 3 * src + 1), v_b = __ldcs(rgbs + 3 * src + 2);
                    const float v_a0 = NA >= 1 ? __ldcs(amb0 + src) : 0.0f, v_a1 = NA >= 2 ? __ldcs(amb1 + src) : 0.0f, v_u = UNC ? __ldcs(unc + src) : 0.0f;
                    b_sg[dst] = v_sg; b_dl[2 * dst] = v_d0; b_dl[2 * dst + 1] = v_d1;
                    b_rgb[3 * dst] = v_r; b_rgb[3 * dst + 1] = v_g; b_rgb[3 * dst + 2] = v_b;
                    if (NA >= 1) b_a0[dst] = v_a0;
                    if (NA >= 2) b_a1[dst] = v_a1;
                    if (UNC) b_u[dst] = v_u;
                }
            }
            __syncwarp();                          // a warp only reads what it staged itself
            if (in_group) {
                const uint32_t o = slot - base;
                train_fwd_ray<AMB, NA, UNC>(b_sg + o, b_rgb + 3 * o, b_a0 + o, b_a1 + o, b_u + o, b_dl + 2 * o, num, T_thresh, ws, a0, a1, u, d, r, g, b);
            }
        }
        __syncthreads();                           // the next pass overwrites the staging buffers
        base = end;
    }
    if (n >= N) return;
    weights_sum[idx] = ws;
    if (NA >= 1) amb0_sum[idx] = a0;
    if (NA >= 2) amb1_sum[idx] = a1;
    if (UNC) unc_sum[idx] = u;
    depth[idx] = d;
    image[3 * (size_t)idx] = r; image[3 * (size_t)idx + 1] = g; image[3 * (size_t)idx + 2] = b;
}

// ---------------------------------------------------------------------------------------------------
// training backward.  Gradients of samples after the early stop stay 0 (caller pre-zeroes; the staged path
// rewrites the whole tiled range, zeros included).
// ---------------------------------------------------------------------------------------------------
struct RayGrads { float gws, ga0, ga1, gu, gi0, gi1, gi2, wsF, a0F, uF, rF, gF, bF; };

// IN-PLACE capable: gs/grgb/ga0/ga1/gu may alias sg/rgb/a0p/a1p/up (each sample is read before it is written)
template <int AMB, int NA, bool UNC, bool ZERO_TAIL>
__device__ __forceinline__ void train_bwd_ray(const float *sg, const float *rgb, const float *a0p, const float *up, const float *dl,
                                              uint32_t num, float T_thresh, const RayGrads &q,
                                              float *gs, float *grgb, float *ga0, float *ga1, float *gu) {
    float T = 1.0f, r = 0, g = 0, b = 0, a = 0, u = 0;
    uint32_t k = 0;
    for (; k < num; k++) {
        const float sigma = sg[k], delta = dl[2 * k];
        const float c0 = rgb[3 * k], c1 = rgb[3 * k + 1], c2 = rgb[3 * k + 2];
        const float av = (AMB == 2) ? a0p[k] : 0.0f;
        const float uv = UNC ? up[k] : 0.0f;
        const float alpha = alpha_of(sigma, delta);
        const float w = __fmul_rn(alpha, T);
        r = __fmaf_rn(w, c0, r); g = __fmaf_rn(w, c1, g); b = __fmaf_rn(w, c2, b);
        if (AMB == 2) a = __fmaf_rn(w, av, a);
        if (UNC) u = __fmaf_rn(w, uv, u);
        T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
        grgb[3 * k] = __fmul_rn(q.gi0, w); grgb[3 * k + 1] = __fmul_rn(q.gi1, w); grgb[3 * k + 2] = __fmul_rn(q.gi2, w);
        if (NA >= 1) ga0[k] = (AMB == 2) ? __fmul_rn(q.ga0, w) : q.ga0;
        if (NA >= 2) ga1[k] = q.ga1;
        if (UNC) gu[k] = __fmul_rn(q.gu, w);
        // summation order of the reference's sm_100a SASS (all four variants): round gi1*t1 first, fma the others onto it,
        // then ADD the loop-invariant gws*(1-ws_final) product
        float acc = __fmul_rn(q.gi1, __fmaf_rn(T, c1, -__fsub_rn(q.gF, g)));
        acc = __fmaf_rn(q.gi0, __fmaf_rn(T, c0, -__fsub_rn(q.rF, r)), acc);
        acc = __fmaf_rn(q.gi2, __fmaf_rn(T, c2, -__fsub_rn(q.bF, b)), acc);
        if (AMB == 2) acc = __fmaf_rn(q.ga0, __fmaf_rn(T, av, -__fsub_rn(q.a0F, a)), acc);
        if (UNC) acc = __fmaf_rn(q.gu, __fmaf_rn(T, uv, -__fsub_rn(q.uF, u)), acc);
        acc = __fadd_rn(__fmul_rn(q.gws, __fsub_rn(1.0f, q.wsF)), acc);
        gs[k] = __fmul_rn(delta, acc);
        if (T < T_thresh) { k++; break; }
    }
    if (ZERO_TAIL) {
        for (; k < num; k++) {
            gs[k] = 0.0f; grgb[3 * k] = 0.0f; grgb[3 * k + 1] = 0.0f; grgb[3 * k + 2] = 0.0f;
            if (NA >= 1) ga0[k] = 0.0f;
            if (NA >= 2) ga1[k] = 0.0f;
            if (UNC) gu[k] = 0.0f;
        }
    }
}

template <int AMB, int NA, bool UNC>
__global__ void __launch_bounds__(CT_THREADS) k_comp_train_bwd(
        const float *__restrict__ g_ws, const float *__restrict__ g_a0, const float *__restrict__ g_a1,
        const float *__restrict__ g_u, const float *__restrict__ g_img,
        const float *__restrict__ sigmas, const float *__restrict__ rgbs, const float *__restrict__ amb0,
        const float *__restrict__ unc, const float *__restrict__ deltas, const int32_t *__restrict__ rays,
        const float *__restrict__ weights_sum, const float *__restrict__ amb0_sum, const float *__restrict__ unc_sum,
        const float *__restrict__ image, uint32_t M, uint32_t N, float T_thresh, uint32_t cap,
        float *__restrict__ grad_sigmas, float *__restrict__ grad_rgbs, float *__restrict__ grad_a0,
        float *__restrict__ grad_a1, float *__restrict__ grad_u) {
    extern __shared__ __align__(16) float sm[];
    const uint32_t n = blockIdx.x * CT_THREADS + threadIdx.x;
    uint32_t idx = 0, off = 0, num = 0;
    if (n < N) { idx = (uint32_t)rays[3 * (size_t)n]; off = (uint32_t)rays[3 * (size_t)n + 1]; num = (uint32_t)rays[3 * (size_t)n + 2]; }
    const bool valid = (n < N) && !(num == 0 || off + num > M);
    __shared__ __align__(8) uint64_t s_bar;
    if (threadIdx.x == 0) { tc5::mbar_init(&s_bar, 1); tc5::fence_mbar_init(); }
    uint32_t lo, total, slot;
    const bool tiled = cta_tiling(valid, off, num, lo, total, slot);
    RayGrads q = {};
    if (valid) {
        q.gws = g_ws[idx];
        if (NA >= 1) q.ga0 = g_a0[idx];
        if (NA >= 2) q.ga1 = g_a1[idx];
        if (UNC) { q.gu = g_u[idx]; q.uF = unc_sum[idx]; }
        if (AMB == 2) q.a0F = amb0_sum[idx];
        q.gi0 = g_img[3 * (size_t)idx]; q.gi1 = g_img[3 * (size_t)idx + 1]; q.gi2 = g_img[3 * (size_t)idx + 2];
        q.wsF = weights_sum[idx];
        q.rF = image[3 * (size_t)idx]; q.gF = image[3 * (size_t)idx + 1]; q.bF = image[3 * (size_t)idx + 2];
    }
    auto walk_global = [&]() {
        train_bwd_ray<AMB, NA, UNC, false>(sigmas + off, rgbs + 3 * (size_t)off, AMB == 2 ? amb0 + off : nullptr, UNC ? unc + off : nullptr,
                                           deltas + 2 * (size_t)off, num, T_thresh, q, grad_sigmas + off, grad_rgbs + 3 * (size_t)off,
                                           NA >= 1 ? grad_a0 + off : nullptr, NA >= 2 ? grad_a1 + off : nullptr, UNC ? grad_u + off : nullptr);
    };
    // the a0 span doubles as grad_a0 staging (input only when AMB == 2), a1 is output-only (grad_a1 = per-ray constant), u doubles as grad_u
    float *b_sg = sm, *b_dl = b_sg + cap + CT_PAD, *b_rgb = b_dl + 2 * cap + CT_PAD, *b_a0 = b_rgb + 3 * cap + CT_PAD;
    float *b_a1 = b_a0 + (NA >= 1 ? cap + CT_PAD : 0), *b_u = b_a1 + (NA >= 2 ? cap + CT_PAD : 0);
    const uint32_t cnt = valid ? num : 0u;
    uint32_t bar_phase = 0;
    for (uint32_t base = 0; base < total;) {   // groups of consecutive rows that fit the staging buffers (see the forward kernel)
        const uint32_t end = group_end(valid, slot, cnt, base, total, cap);
        if (end == base) {
            const bool mine = valid && slot == base;
            if (mine) {
                // the staged path rewrites its whole range, zeros after the early stop included; the direct path relies on pre-zeroed outputs like the reference
                walk_global();
            }
            base = long_row_end(mine, slot, cnt);
            continue;
        }
        const bool in_group = valid && slot >= base && slot < end;
        const uint32_t gcnt = end - base;
        if (!tiled) {
            // foreign row order (see the forward kernel): the warp gathers the samples of its rays of this group sample-parallel, the rays are walked in shared
            // memory (gradients overwrite the inputs in place, zeros after the early stop), and the gradients are scattered back the same way.
            const uint32_t lane = threadIdx.x & 31u;
            const uint32_t w_lo = max(__shfl_sync(0xffffffffu, slot, 0), base), w_hi = min(__shfl_sync(0xffffffffu, slot + cnt, 31), end);
            auto source_of = [&](uint32_t i) -> size_t {         // global sample index of staged sample i (all lanes call it)
                uint32_t j = 0;
#pragma unroll
                for (uint32_t step = 16; step; step >>= 1) {
                    const uint32_t sj = __shfl_sync(0xffffffffu, slot, (j + step) & 31u);
                    if (sj <= i) j += step;
                }
                return (size_t)__shfl_sync(0xffffffffu, off, j) + (i - __shfl_sync(0xffffffffu, slot, j));
            };
#pragma unroll 1
            for (uint32_t i0 = w_lo; i0 < w_hi; i0 += 32) {
                const bool act = i0 + lane < w_hi;
                const uint32_t i = act ? i0 + lane : w_hi - 1;
                const size_t src = source_of(i);
                if (act) {
                    const uint32_t dst = i - base;
                    const float v_sg = __ldcs(sigmas + src), v_d0 = __ldcs(deltas + 2 * src), v_d1 = __ldcs(deltas + 2 * src + 1);
                    const float v_r = __ldcs(rgbs + 3 * src), v_g = __ldcs(rgbs + 3 * src + 1), v_b = __ldcs(rgbs + 3 * src + 2);
                    const float v_a0 = AMB == 2 ? __ldcs(amb0 + src) : 0.0f, v_u = UNC ? __ldcs(unc + src) : 0.0f;
                    b_sg[dst] = v_sg; b_dl[2 * dst] = v_d0; b_dl[2 * dst + 1] = v_d1;
                    b_rgb[3 * dst] = v_r; b_rgb[3 * dst + 1] = v_g; b_rgb[3 * dst + 2] = v_b;
                    if (AMB == 2) b_a0[dst] = v_a0;
                    if (UNC) b_u[dst] = v_u;
                }
            }
            __syncwarp();                          // a warp only reads what it staged itself
            if (in_group) {
                const uint32_t o = slot - base;
                train_bwd_ray<AMB, NA, UNC, true>(b_sg + o, b_rgb + 3 * o, b_a0 + o, b_u + o, b_dl + 2 * o, num, T_thresh, q,
                                                  b_sg + o, b_rgb + 3 * o, b_a0 + o, b_a1 + o, b_u + o);
            }
            __syncwarp();
#pragma unroll 1
            for (uint32_t i0 = w_lo; i0 < w_hi; i0 += 32) {
                const bool act = i0 + lane < w_hi;
                const uint32_t i = act ? i0 + lane : w_hi - 1;
                const size_t src = source_of(i);
                if (act) {
                    const uint32_t dst = i - base;
                    __stcs(grad_sigmas + src, b_sg[dst]);
                    __stcs(grad_rgbs + 3 * src, b_rgb[3 * dst]); __stcs(grad_rgbs + 3 * src + 1, b_rgb[3 * dst + 1]); __stcs(grad_rgbs + 3 * src + 2, b_rgb[3 * dst + 2]);
                    if (NA >= 1) __stcs(grad_a0 + src, b_a0[dst]);
                    if (NA >= 2) __stcs(grad_a1 + src, b_a1[dst]);
                    if (UNC) __stcs(grad_u + src, b_u[dst]);
                }
            }
            __syncthreads();                       // the next pass overwrites the staging buffers
            base = end;
            continue;
        }
        const size_t g0 = (size_t)lo + base;
        const Span p_sg = span_of(b_sg, sigmas + g0, gcnt), p_dl = span_of(b_dl, deltas + 2 * g0, 2 * gcnt), p_rgb = span_of(b_rgb, rgbs + 3 * g0, 3 * gcnt);
        const Span p_a0 = NA >= 1 ? span_of(b_a0, AMB == 2 ? amb0 + g0 : grad_a0 + g0, gcnt) : Span{b_a0, 0, 0, 0};
        const Span p_a1 = NA >= 2 ? span_of(b_a1, grad_a1 + g0, gcnt) : Span{b_a1, 0, 0, 0};
        const Span p_u = UNC ? span_of(b_u, unc + g0, gcnt) : Span{b_u, 0, 0, 0};
        if (threadIdx.x == 0) tc5::mbar_expect_tx(&s_bar, 4u * (p_sg.body + p_dl.body + p_rgb.body + (AMB == 2 ? p_a0.body : 0u) + p_u.body));
        span_load(p_sg, sigmas + g0, &s_bar);
        span_load(p_dl, deltas + 2 * g0, &s_bar);
        span_load(p_rgb, rgbs + 3 * g0, &s_bar);
        if (AMB == 2) span_load(p_a0, amb0 + g0, &s_bar);
        if (UNC) span_load(p_u, unc + g0, &s_bar);
        __syncthreads();
        tc5::mbar_wait(&s_bar, bar_phase);
        bar_phase ^= 1u;
        if (in_group) {
            const uint32_t o = slot - base;
            train_bwd_ray<AMB, NA, UNC, true>(p_sg.s + o, p_rgb.s + 3 * o, p_a0.s + o, p_u.s + o, p_dl.s + 2 * o, num, T_thresh, q,
                                              p_sg.s + o, p_rgb.s + 3 * o, p_a0.s + o, p_a1.s + o, p_u.s + o);
        }
        tc5::fence_proxy_async();              // this thread's shared-memory writes -> visible to the bulk-store (async) proxy
        __syncthreads();
        span_store_any(p_sg, grad_sigmas + g0);
        span_store_any(p_rgb, grad_rgbs + 3 * g0);
        if (NA >= 1) span_store_any(p_a0, grad_a0 + g0);
        if (NA >= 2) span_store_any(p_a1, grad_a1 + g0);
        if (UNC) span_store_any(p_u, grad_u + g0);
        if (threadIdx.x == 0) { tc5::bulk_commit(); tc5::bulk_wait_read0(); }     // shared memory must outlive the bulk reads
        __syncthreads();                       // ... and the next pass overwrites it
        base = end;
    }
}

// ---------------------------------------------------------------------------------------------------
// inference: incremental, in place (raymarching.cu:943-1029 and variants)
// ---------------------------------------------------------------------------------------------------
template <int AMB, int NA, bool UNC>
__global__ void __launch_bounds__(128) k_comp_infer(
        uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *__restrict__ rays_alive, float *__restrict__ rays_t,
        const float *__restrict__ sigmas, const float *__restrict__ rgbs, const float *__restrict__ deltas,
        const float *__restrict__ amb0, const float *__restrict__ amb1, const float *__restrict__ unc,
        float *__restrict__ weights_sum, float *__restrict__ depth, float *__restrict__ image,
        float *__restrict__ amb0_sum, float *__restrict__ amb1_sum, float *__restrict__ unc_sum) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int32_t idx = rays_alive[n];
    const size_t base = (size_t)n * n_step;
    float t = rays_t[idx], ws = weights_sum[idx], d = depth[idx];
    float r = image[3 * (size_t)idx], g = image[3 * (size_t)idx + 1], b = image[3 * (size_t)idx + 2];
    float a0 = NA >= 1 ? amb0_sum[idx] : 0.0f, a1 = NA >= 2 ? amb1_sum[idx] : 0.0f, u = UNC ? unc_sum[idx] : 0.0f;
    uint32_t step = 0;
    while (step < n_step) {
        const size_t i = base + step;
        const float delta = deltas[2 * i];
        if (delta == 0.0f) break;
        const float alpha = alpha_of(sigmas[i], delta);
        const float T = __fsub_rn(1.0f, ws);
        const float w = __fmul_rn(alpha, T);
        ws = __fadd_rn(ws, w);
        t = deltas[2 * i + 1];
        d = __fmaf_rn(w, t, d);
        r = __fmaf_rn(w, rgbs[3 * i], r); g = __fmaf_rn(w, rgbs[3 * i + 1], g); b = __fmaf_rn(w, rgbs[3 * i + 2], b);
        if (NA >= 1) a0 = (AMB == 2) ? __fmaf_rn(w, amb0[i], a0) : __fadd_rn(a0, amb0[i]);
        if (NA >= 2) a1 = (AMB == 2) ? __fmaf_rn(w, amb1[i], a1) : __fadd_rn(a1, amb1[i]);
        if (UNC) u = __fmaf_rn(w, unc[i], u);
        if (T < T_thresh) break;
        step++;
    }
    if (step < n_step) rays_alive[n] = -1; else rays_t[idx] = t;
    weights_sum[idx] = ws; depth[idx] = d;
    image[3 * (size_t)idx] = r; image[3 * (size_t)idx + 1] = g; image[3 * (size_t)idx + 2] = b;
    if (NA >= 1) amb0_sum[idx] = a0;
    if (NA >= 2) amb1_sum[idx] = a1;
    if (UNC) unc_sum[idx] = u;
}

// ---------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------
template <int AMB, int NA, bool UNC>
static int launch_train_fwd(const char *what, const float *sigmas, const float *rgbs, const float *amb0, const float *amb1, const float *unc,
                            const float *deltas, const int32_t *rays, uint32_t M, uint32_t N, float T_thresh,
                            float *ws, float *a0s, float *a1s, float *us, float *depth, float *image, void *stream) {
    B2N_REQUIRE(sigmas && rgbs && deltas && rays && ws && depth && image, "%s: null pointer", what);
    B2N_REQUIRE((NA < 1 || (amb0 && a0s)) && (NA < 2 || (amb1 && a1s)) && (!UNC || (unc && us)), "%s: null pointer", what);
    if (N == 0) return 0;
    const uint32_t cap = stage_cap();
    const size_t smem = Stage<NA, UNC>::bytes(cap);
    auto kern = k_comp_train_fwd<AMB, NA, UNC>;
    B2N_SMEM(kern, smem);
    kern<<<ceil_div<uint32_t>(N, CT_THREADS), CT_THREADS, smem, as_stream(stream)>>>(sigmas, rgbs, amb0, amb1, unc, deltas, rays, M, N, T_thresh, cap,
                                                                                     ws, a0s, a1s, us, depth, image);
    return check_launch(what);
}

template <int AMB, int NA, bool UNC>
static int launch_train_bwd(const char *what, const float *g_ws, const float *g_a0, const float *g_a1, const float *g_u, const float *g_img,
                            const float *sigmas, const float *rgbs, const float *amb0, const float *unc, const float *deltas,
                            const int32_t *rays, const float *ws, const float *a0s, const float *us, const float *image,
                            uint32_t M, uint32_t N, float T_thresh, float *gs, float *grgb, float *ga0, float *ga1, float *gu, void *stream) {
    B2N_REQUIRE(g_ws && g_img && sigmas && rgbs && deltas && rays && ws && image && gs && grgb, "%s: null pointer", what);
    B2N_REQUIRE((NA < 1 || (g_a0 && ga0)) && (NA < 2 || (g_a1 && ga1)) && (!UNC || (g_u && unc && us && gu)) && (AMB != 2 || (amb0 && a0s)),
                "%s: null pointer", what);
    if (N == 0) return 0;
    const uint32_t cap = stage_cap();
    const size_t smem = Stage<NA, UNC>::bytes(cap);
    auto kern = k_comp_train_bwd<AMB, NA, UNC>;
    B2N_SMEM(kern, smem);
    kern<<<ceil_div<uint32_t>(N, CT_THREADS), CT_THREADS, smem, as_stream(stream)>>>(g_ws, g_a0, g_a1, g_u, g_img, sigmas, rgbs, amb0, unc, deltas, rays,
                                                                                     ws, a0s, us, image, M, N, T_thresh, cap, gs, grgb, ga0, ga1, gu);
    return check_launch(what);
}

template <int AMB, int NA, bool UNC>
static int launch_infer(const char *what, uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *rays_alive, float *rays_t,
                        const float *sigmas, const float *rgbs, const float *deltas, const float *amb0, const float *amb1, const float *unc,
                        float *ws, float *depth, float *image, float *a0s, float *a1s, float *us, void *stream) {
    B2N_REQUIRE(rays_alive && rays_t && sigmas && rgbs && deltas && ws && depth && image, "%s: null pointer", what);
    B2N_REQUIRE((NA < 1 || (amb0 && a0s)) && (NA < 2 || (amb1 && a1s)) && (!UNC || (unc && us)), "%s: null pointer", what);
    if (n_alive == 0) return 0;
    k_comp_infer<AMB, NA, UNC><<<ceil_div<uint32_t>(n_alive, 128), 128, 0, as_stream(stream)>>>(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas,
                                                                                                 amb0, amb1, unc, ws, depth, image, a0s, a1s, us);
    return check_launch(what);
}

}  // namespace b2n

using namespace b2n;

extern "C" {

int b2n_composite_rays_train_forward(const float *sigmas, const float *rgbs, const float *ambient, const float *deltas, const int32_t *rays,
                                     uint32_t M, uint32_t N, float T_thresh, float *weights_sum, float *ambient_sum, float *depth, float *image, void *stream) {
    return launch_train_fwd<1, 1, false>("composite_rays_train_forward", sigmas, rgbs, ambient, nullptr, nullptr, deltas, rays, M, N, T_thresh,
                                         weights_sum, ambient_sum, nullptr, nullptr, depth, image, stream);
}
int b2n_composite_rays_train_backward(const float *grad_weights_sum, const float *grad_ambient_sum, const float *grad_image, const float *sigmas,
                                      const float *rgbs, const float *ambient, const float *deltas, const int32_t *rays, const float *weights_sum,
                                      const float *ambient_sum, const float *image, uint32_t M, uint32_t N, float T_thresh,
                                      float *grad_sigmas, float *grad_rgbs, float *grad_ambient, void *stream) {
    (void)ambient; (void)ambient_sum;
    return launch_train_bwd<1, 1, false>("composite_rays_train_backward", grad_weights_sum, grad_ambient_sum, nullptr, nullptr, grad_image, sigmas, rgbs,
                                         nullptr, nullptr, deltas, rays, weights_sum, nullptr, nullptr, image, M, N, T_thresh,
                                         grad_sigmas, grad_rgbs, grad_ambient, nullptr, nullptr, stream);
}
int b2n_composite_rays_train_sigma_forward(const float *sigmas, const float *rgbs, const float *ambient, const float *deltas, const int32_t *rays,
                                           uint32_t M, uint32_t N, float T_thresh, float *weights_sum, float *ambient_sum, float *depth, float *image, void *stream) {
    return launch_train_fwd<2, 1, false>("composite_rays_train_sigma_forward", sigmas, rgbs, ambient, nullptr, nullptr, deltas, rays, M, N, T_thresh,
                                         weights_sum, ambient_sum, nullptr, nullptr, depth, image, stream);
}
int b2n_composite_rays_train_sigma_backward(const float *grad_weights_sum, const float *grad_ambient_sum, const float *grad_image, const float *sigmas,
                                            const float *rgbs, const float *ambient, const float *deltas, const int32_t *rays, const float *weights_sum,
                                            const float *ambient_sum, const float *image, uint32_t M, uint32_t N, float T_thresh,
                                            float *grad_sigmas, float *grad_rgbs, float *grad_ambient, void *stream) {
    return launch_train_bwd<2, 1, false>("composite_rays_train_sigma_backward", grad_weights_sum, grad_ambient_sum, nullptr, nullptr, grad_image, sigmas, rgbs,
                                         ambient, nullptr, deltas, rays, weights_sum, ambient_sum, nullptr, image, M, N, T_thresh,
                                         grad_sigmas, grad_rgbs, grad_ambient, nullptr, nullptr, stream);
}
int b2n_composite_rays_train_uncertainty_forward(const float *sigmas, const float *rgbs, const float *ambient, const float *uncertainty, const float *deltas,
                                                 const int32_t *rays, uint32_t M, uint32_t N, float T_thresh, float *weights_sum, float *ambient_sum,
                                                 float *uncertainty_sum, float *depth, float *image, void *stream) {
    return launch_train_fwd<1, 1, true>("composite_rays_train_uncertainty_forward", sigmas, rgbs, ambient, nullptr, uncertainty, deltas, rays, M, N, T_thresh,
                                        weights_sum, ambient_sum, nullptr, uncertainty_sum, depth, image, stream);
}
int b2n_composite_rays_train_uncertainty_backward(const float *grad_weights_sum, const float *grad_ambient_sum, const float *grad_uncertainty_sum,
                                                  const float *grad_image, const float *sigmas, const float *rgbs, const float *ambient,
                                                  const float *uncertainty, const float *deltas, const int32_t *rays, const float *weights_sum,
                                                  const float *ambient_sum, const float *uncertainty_sum, const float *image, uint32_t M, uint32_t N,
                                                  float T_thresh, float *grad_sigmas, float *grad_rgbs, float *grad_ambient, float *grad_uncertainty, void *stream) {
    (void)ambient; (void)ambient_sum;
    return launch_train_bwd<1, 1, true>("composite_rays_train_uncertainty_backward", grad_weights_sum, grad_ambient_sum, nullptr, grad_uncertainty_sum, grad_image,
                                        sigmas, rgbs, nullptr, uncertainty, deltas, rays, weights_sum, nullptr, uncertainty_sum, image, M, N, T_thresh,
                                        grad_sigmas, grad_rgbs, grad_ambient, nullptr, grad_uncertainty, stream);
}
int b2n_composite_rays_train_triplane_forward(const float *sigmas, const float *rgbs, const float *amb_aud, const float *amb_eye, const float *uncertainty,
                                              const float *deltas, const int32_t *rays, uint32_t M, uint32_t N, float T_thresh, float *weights_sum,
                                              float *amb_aud_sum, float *amb_eye_sum, float *uncertainty_sum, float *depth, float *image, void *stream) {
    return launch_train_fwd<1, 2, true>("composite_rays_train_triplane_forward", sigmas, rgbs, amb_aud, amb_eye, uncertainty, deltas, rays, M, N, T_thresh,
                                        weights_sum, amb_aud_sum, amb_eye_sum, uncertainty_sum, depth, image, stream);
}
int b2n_composite_rays_train_triplane_backward(const float *grad_weights_sum, const float *grad_amb_aud_sum, const float *grad_amb_eye_sum,
                                               const float *grad_uncertainty_sum, const float *grad_image, const float *sigmas, const float *rgbs,
                                               const float *amb_aud, const float *amb_eye, const float *uncertainty, const float *deltas,
                                               const int32_t *rays, const float *weights_sum, const float *amb_aud_sum, const float *amb_eye_sum,
                                               const float *uncertainty_sum, const float *image, uint32_t M, uint32_t N, float T_thresh,
                                               float *grad_sigmas, float *grad_rgbs, float *grad_amb_aud, float *grad_amb_eye, float *grad_uncertainty, void *stream) {
    (void)amb_aud; (void)amb_eye; (void)amb_aud_sum; (void)amb_eye_sum;
    return launch_train_bwd<1, 2, true>("composite_rays_train_triplane_backward", grad_weights_sum, grad_amb_aud_sum, grad_amb_eye_sum, grad_uncertainty_sum,
                                        grad_image, sigmas, rgbs, nullptr, uncertainty, deltas, rays, weights_sum, nullptr, uncertainty_sum, image, M, N,
                                        T_thresh, grad_sigmas, grad_rgbs, grad_amb_aud, grad_amb_eye, grad_uncertainty, stream);
}

int b2n_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *rays_alive, float *rays_t, const float *sigmas, const float *rgbs,
                       const float *deltas, float *weights_sum, float *depth, float *image, void *stream) {
    return launch_infer<0, 0, false>("composite_rays", n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, nullptr, nullptr, nullptr,
                                     weights_sum, depth, image, nullptr, nullptr, nullptr, stream);
}
int b2n_composite_rays_ambient(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *rays_alive, float *rays_t, const float *sigmas, const float *rgbs,
                               const float *deltas, const float *ambients, float *weights_sum, float *depth, float *image, float *ambient_sum, void *stream) {
    return launch_infer<1, 1, false>("composite_rays_ambient", n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, ambients, nullptr, nullptr,
                                     weights_sum, depth, image, ambient_sum, nullptr, nullptr, stream);
}
int b2n_composite_rays_ambient_sigma(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *rays_alive, float *rays_t, const float *sigmas,
                                     const float *rgbs, const float *deltas, const float *ambients, float *weights_sum, float *depth, float *image,
                                     float *ambient_sum, void *stream) {
    return launch_infer<2, 1, false>("composite_rays_ambient_sigma", n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, ambients, nullptr, nullptr,
                                     weights_sum, depth, image, ambient_sum, nullptr, nullptr, stream);
}
int b2n_composite_rays_uncertainty(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *rays_alive, float *rays_t, const float *sigmas,
                                   const float *rgbs, const float *deltas, const float *ambients, const float *uncertainties, float *weights_sum,
                                   float *depth, float *image, float *ambient_sum, float *uncertainty_sum, void *stream) {
    return launch_infer<1, 1, true>("composite_rays_uncertainty", n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, ambients, nullptr,
                                    uncertainties, weights_sum, depth, image, ambient_sum, nullptr, uncertainty_sum, stream);
}
int b2n_composite_rays_triplane(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *rays_alive, float *rays_t, const float *sigmas,
                                const float *rgbs, const float *deltas, const float *ambs_aud, const float *ambs_eye, const float *uncertainties,
                                float *weights_sum, float *depth, float *image, float *amb_aud_sum, float *amb_eye_sum, float *uncertainty_sum, void *stream) {
    return launch_infer<1, 2, true>("composite_rays_triplane", n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, ambs_aud, ambs_eye,
                                    uncertainties, weights_sum, depth, image, amb_aud_sum, amb_eye_sum, uncertainty_sum, stream);
}

}  // extern "C"
