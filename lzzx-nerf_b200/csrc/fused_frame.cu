// fused_frame.cu — one inference frame (renderer.py:406-570, run_cuda_for_inference) with no host synchronisation.
//
// The reference drives its march / network / composite loop from the host and reads the alive-ray count back every
// iteration (rays_alive[rays_alive >= 0], renderer.py:542 — 16 D2H syncs per frame).  Here the loop state lives in a
// small device-side control block per iteration:
//     ctrl = { n_alive, n_step, step, n_samples, done, ... }
// rewritten in place by the last CTA of each iteration's composite kernel, so the loop body has FIXED kernel arguments: it is
// either a CUDA-graph WHILE node (b2n_frame_graph_*) or max_steps unrolled copies whose kernels exit at once after the loop
// has ended (b2n_render_frame).  Three kernels per iteration: march, fused head, composite.  Semantics are the reference's:
//     n_step = max(min(N // n_alive, 8), 1)                                   renderer.py:506-513
//     march_rays(n_alive, n_step, ...) -> network -> composite_rays_triplane  renderer.py:518-534
//     rays_alive = rays_alive[rays_alive >= 0]                                renderer.py:542
//         (here: survivors are appended to the other alive buffer with one atomic per warp; the list order differs from the
//          reference's stable compaction, which no per-ray result depends on)
//     step += n_step; loop while step < max_steps and n_alive > 0             renderer.py:503
// Per-ray arithmetic is the same code as the per-op kernels (dda.cuh, composite order), so images match the op-by-op
// path bit for bit given the same network outputs.
#include <stdlib.h>
#include "common.cuh"
#include "dda.cuh"
#include "fused_head.cuh"

struct b2n_model;
namespace b2n {
int head_forward_on_model(const b2n_model *m, const float *xyzs, const float *dirs, uint32_t M, const float *enc_a, const float *ind_code, const float *eye,
                          const int32_t *n_valid, float density_scale, float *sigmas, float *rgbs, float *amb_aud, float *amb_eye, float *unc, cudaStream_t st,
                          const float *live_deltas = nullptr, const b2n_head_saved *saved = nullptr, uint32_t head_ctas = 0);
}

namespace b2n {

struct FrameCtrl { int32_t n_alive, n_step, step, n_samples, done, buf, iter, n_live; };     // 32 B; buf = which alive[] buffer holds this iteration's ids

struct FrameWs {            // carved out of the caller's workspace
    FrameCtrl *ctrl;        // the iteration being executed (fixed address: kernel arguments never change, so the loop can be a graph WHILE node)
    int32_t *alive[2];      // ping-pong alive ray ids [N]
    int32_t *alive_mid;     // rays of this iteration that produced samples (input of head / composite)
    int32_t *counters;      // [0] |alive_mid|, [1] |next alive|, [2],[3] last-block tickets
    float *occ_box;         // [6] world-space box around every occupied cell, grown by 2 cells (empty: min > max)
    float *nears, *fars, *rays_t, *ws, *depth, *aud_sum, *eye_sum, *unc_sum, *image;   // per ray
    float *xyzs, *dirs, *deltas, *sigmas, *rgbs, *amb_aud, *amb_eye, *unc;              // per sample (<= N + 128)
};

constexpr uint32_t FR_THREADS = 128;
// (With the four-warpgroup head kernel nothing co-resides on an SM that runs a head CTA any more — 512 threads x 125 registers — so the march CTA is sized for
// coherent list appends and one-wave launches, see k_frame_march.)
constexpr uint32_t FM_THREADS = 128;
constexpr uint32_t FR_MAX_ITERS = 64;

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t carve(FrameWs *w, uint8_t *base, uint32_t N) {
    size_t off = 0;
    auto take = [&](size_t bytes) { uint8_t *p = base ? base + off : nullptr; off = align_up(off + bytes, 256); return p; };
    const size_t Np = (size_t)N + 128;
    FrameWs t;
    t.ctrl = (FrameCtrl *)take(sizeof(FrameCtrl) * 8);
    t.alive[0] = (int32_t *)take(4 * Np); t.alive[1] = (int32_t *)take(4 * Np);
    t.alive_mid = (int32_t *)take(4 * Np);
    t.counters = (int32_t *)take(4 * 8);
    t.occ_box = (float *)take(4 * OCC_BOX_FLOATS);   // per-CTA partial boxes of k_occ_box (reduced by every CTA of k_frame_init)
    float **per_ray[] = {&t.nears, &t.fars, &t.rays_t, &t.ws, &t.depth, &t.aud_sum, &t.eye_sum, &t.unc_sum};
    for (auto p : per_ray) *p = (float *)take(4 * Np);
    t.image = (float *)take(12 * Np);
    t.xyzs = (float *)take(12 * Np); t.dirs = (float *)take(12 * Np); t.deltas = (float *)take(8 * Np);
    t.sigmas = (float *)take(4 * Np); t.rgbs = (float *)take(12 * Np);
    t.amb_aud = (float *)take(4 * Np); t.amb_eye = (float *)take(4 * Np); t.unc = (float *)take(4 * Np);
    if (w) *w = t;
    return off;
}

// near/far + state reset + ctrl[0].  The marching interval of every ray is clipped to the grown occupied box of the bitfield (k_occ_box, raymarch.cu; exactness
// argument at dda.cuh:clip_to_box): the far end moves to the box exit, a ray that misses the box never starts, and the orbit t <- t + step_of(t) is run without
// probing up to the box entry, so the first march iteration starts where the samples are instead of walking ~35 empty cells per ray.
__global__ void __launch_bounds__(256) k_frame_init(const float *__restrict__ rays_o, const float *__restrict__ rays_d, uint32_t N, float min_near,
                                                     float a0, float a1, float a2, float a3, float a4, float a5, uint32_t max_steps, float bound, float dt_gamma,
                                                     uint32_t C, uint32_t H, int clip, uint32_t tile_w, FrameWs w) {
    __shared__ float bx[6];
    if (threadIdx.x < 6 * 32) {                   // warp a reduces component a of the partial boxes
        const uint32_t a = threadIdx.x >> 5, lane = threadIdx.x & 31;
        float v = a < 3 ? 3.0e38f : -3.0e38f;
        for (uint32_t q = lane; q < OCC_PARTS; q += 32) v = a < 3 ? fminf(v, w.occ_box[q * 6 + a]) : fmaxf(v, w.occ_box[q * 6 + a]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const float u = __shfl_xor_sync(0xffffffffu, v, o); v = a < 3 ? fminf(v, u) : fmaxf(v, u); }
        if (lane == 0) {
            if (a < 3 && v <= -bound) v = -INFINITY;           // the box reaches the cube face: see k_occ_box
            if (a >= 3 && v >= bound) v = INFINITY;
            bx[a] = v;
        }
    }
    __syncthreads();
    DdaRay r;
    r.dx = r.dy = r.dz = 1.0f;
    r.init_common(bound, dt_gamma, max_steps, C, H, 0.0f);
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        r.ox = rays_o[3 * n]; r.oy = rays_o[3 * n + 1]; r.oz = rays_o[3 * n + 2];
        r.rdx = 1.0f / rays_d[3 * n]; r.rdy = 1.0f / rays_d[3 * n + 1]; r.rdz = 1.0f / rays_d[3 * n + 2];
        float tn, tf;
        near_far_one(r.ox, r.oy, r.oz, r.rdx, r.rdy, r.rdz, a0, a1, a2, a3, a4, a5, min_near, tn, tf);
        float t = r.perturb(tn, 0.0f);            // noise = 0 (perturb is off at inference): fma(dt, 0, t) == t, kept for op parity
        if (clip) { r.far = tf; tf = r.clip_to_box(bx, t); }
        w.nears[n] = tn; w.fars[n] = tf; w.rays_t[n] = t;
        w.ws[n] = 0.0f; w.depth[n] = 0.0f; w.aud_sum[n] = 0.0f; w.eye_sum[n] = 0.0f; w.unc_sum[n] = 0.0f;
        w.image[3 * n] = 0.0f; w.image[3 * n + 1] = 0.0f; w.image[3 * n + 2] = 0.0f;
        // first alive list: identity, or — when the caller says the rays are the pixels of an image of width tile_w (row-major) — 8 x 16 pixel tiles, so that a
        // network tile of 128 samples is a compact patch of the image instead of a 128-pixel strip of one row (the tri-plane cells it gathers are shared in both directions)
        uint32_t slot = n;
        if (tile_w) {
            const uint32_t row = n / tile_w, col = n - row * tile_w;
            slot = ((row >> 3) * (tile_w >> 4) + (col >> 4)) * 128u + ((row & 7u) << 4) + (col & 15u);
        }
        w.alive[0][slot] = (int32_t)n;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        FrameCtrl c = {};
        c.n_alive = (int32_t)N; c.n_step = 1; c.step = 0; c.n_samples = 0; c.n_live = 0; c.buf = 0; c.iter = 0;      // N // N = 1
        for (int q = 0; q < 8; q++) w.counters[q] = 0;
        c.done = (N == 0 || max_steps == 0) ? 1 : 0;
        if (c.done) { c.n_alive = 0; c.n_samples = 0; }
        w.ctrl[0] = c;
    }
}

// Warp-aggregated append: lanes with `keep` get consecutive slots of a global list (one atomic per warp; a warp's rays stay adjacent,
// which keeps the table gathers of the head kernel coherent).  The order across warps is arbitrary — per-ray results do not depend on it.
__device__ __forceinline__ uint32_t warp_append(bool keep, int32_t *counter) {
    const uint32_t ballot = __ballot_sync(0xffffffffu, keep), lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0 && ballot) base = (uint32_t)atomicAdd(counter, (int32_t)__popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    return base + __popc(ballot & ((1u << lane) - 1u));
}
// The same with ONE atomic per thread block: the block's kept rays stay adjacent in the list, in thread order.  The lists are scrambled at the granularity of the
// appending unit (the order in which the atomics arrive), and the network kernel's tiles of 128 samples gather from the tables with far better locality when the
// rays of a tile are neighbouring pixels.
template <uint32_t THREADS>
__device__ __forceinline__ uint32_t block_append(bool keep, int32_t *counter) {
    __shared__ uint32_t s_cnt[THREADS / 32], s_base;
    const uint32_t ballot = __ballot_sync(0xffffffffu, keep), lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_cnt[warp] = (uint32_t)__popc(ballot);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
#pragma unroll
        for (uint32_t q = 0; q < THREADS / 32; q++) tot += s_cnt[q];
        s_base = tot ? (uint32_t)atomicAdd(counter, (int32_t)tot) : 0u;
    }
    __syncthreads();
    uint32_t before = 0;
#pragma unroll
    for (uint32_t q = 0; q < THREADS / 32; q++) if (q < warp) before += s_cnt[q];
    return s_base + before + __popc(ballot & ((1u << lane) - 1u));
}
// true in exactly one thread block of the grid: the one that finishes last (all other blocks' writes are visible to it)
__device__ __forceinline__ bool last_block_done(int32_t *ticket) {
    __shared__ int s_last;
    // the block's writes are ordered before the barrier (CTA scope); ONE device-scope fence by the thread that takes the ticket then orders them — cumulatively —
    // before the ticket (the grid-barrier pattern of cooperative groups).  A fence in every thread made each of them wait for its own outstanding stores.
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(ticket, 1) == (int32_t)gridDim.x - 1);
        if (s_last) __threadfence();
    }
    __syncthreads();
    return s_last != 0;
}

// march_rays for the alive rays of this iteration (raymarching.cu:828-929).  Rays that produce no sample at all are finished (the
// reference's composite would see delta == 0 in their first slot, add nothing and drop them, raymarching.cu:2193, 2235) — they are
// dropped HERE, so the head network only evaluates slots of rays that still contribute: the surviving ray ids and their n_step slots are
// written compacted (alive_mid / xyzs / dirs / deltas); slots a ray did not fill are zero (the reference's torch.zeros, raymarching.py:384).
// Held to 64 registers (no spills), 1024 resident threads per SM: a mid-frame launch has ~100 k live rays, which must fit ONE wave of the 148 SMs — at 96 registers
// (640 threads per SM) the last CTAs ran as a second wave and doubled the latency-bound kernel's critical path (3570 -> 3785 frames/s; 768 / 1280 / 1536 threads
// per SM: 3682 / 3752 / 3646).  128-thread CTAs with one list append per CTA (block_append): 3785 -> 3940 (64 / 256 / 512 threads: 3915 / 3900 / 3855).
__global__ void __launch_bounds__(FM_THREADS, 8) k_frame_march(const float *__restrict__ rays_o, const float *__restrict__ rays_d, const uint8_t *__restrict__ grid,
                                                             float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, FrameWs w) {
    const FrameCtrl c = w.ctrl[0];
    if (c.done) return;
    const uint32_t n = blockIdx.x * FM_THREADS + threadIdx.x;
    const uint32_t n_step = (uint32_t)c.n_step;
    const bool valid = n < (uint32_t)c.n_alive;
    float ts[8], dts[8];
    uint32_t step = 0;
    int32_t id = 0;
    DdaRay r;
    if (valid) {
        id = w.alive[c.buf][n];
        r.init(rays_o + 3 * (size_t)id, rays_d + 3 * (size_t)id, bound, dt_gamma, max_steps, C, H, w.fars[id]);
        float t = w.rays_t[id];
        step = r.march_auto<4>(grid, t, n_step, [&](uint32_t ks, float tk, float dt) {
#pragma unroll
            for (uint32_t k = 0; k < 8; k++) if (k == ks) { ts[k] = tk; dts[k] = dt; }
        });
    }
    const bool has = valid && step > 0;
    const uint32_t p = block_append<FM_THREADS>(has, &w.counters[0]);
    // The warp's surviving rays own the contiguous slot range [p0 n_step, (p0 + cnt) n_step): the 8 floats of a slot are staged through a 1 KB per-warp tile
    // (the kernel must fit beside another frame's head CTA, which leaves ~7 KB of shared memory per SM) and leave as full 128-byte lines instead of
    // 4-byte stores at a 12 n_step-byte stride.
    __shared__ float s_stage[FM_THREADS / 32][256];
    float *sw = s_stage[threadIdx.x >> 5];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t ballot = __ballot_sync(0xffffffffu, has), cnt = __popc(ballot), rank = __popc(ballot & ((1u << lane) - 1u));
    if (cnt) {
        const uint32_t p0 = __shfl_sync(0xffffffffu, p - rank, __ffs(ballot) - 1);
        if (has) w.alive_mid[p] = id;
#pragma unroll 1
        for (uint32_t which = 0; which < 3; which++) {                    // 0: xyzs, 1: dirs, 2: deltas
            const uint32_t comps = which == 2 ? 2u : 3u;
            float *dst = which == 0 ? w.xyzs : (which == 1 ? w.dirs : w.deltas);
            const uint32_t rpr = min(32u, 256u / (n_step * comps));       // rays per staging round (n_step <= 8: at least 10)
#pragma unroll 1
            for (uint32_t r0 = 0; r0 < cnt; r0 += rpr) {
                if (has && rank >= r0 && rank < r0 + rpr) {
                    float *q = sw + (rank - r0) * n_step * comps;
#pragma unroll
                    for (uint32_t k = 0; k < 8; k++) {
                        if (k < n_step) {
                            const bool f = k < step;
                            if (which == 0) {      // the sample position is a pure function of t: same expression as DdaRay::probe
                                q[3 * k] = f ? clampf(__fmaf_rn(ts[k], r.dx, r.ox), -bound, bound) : 0.0f;
                                q[3 * k + 1] = f ? clampf(__fmaf_rn(ts[k], r.dy, r.oy), -bound, bound) : 0.0f;
                                q[3 * k + 2] = f ? clampf(__fmaf_rn(ts[k], r.dz, r.oz), -bound, bound) : 0.0f;
                            } else if (which == 1) { q[3 * k] = f ? r.dx : 0.0f; q[3 * k + 1] = f ? r.dy : 0.0f; q[3 * k + 2] = f ? r.dz : 0.0f; }
                            else { q[2 * k] = f ? dts[k] : 0.0f; q[2 * k + 1] = f ? __fadd_rn(ts[k], dts[k]) : 0.0f; }
                        }
                    }
                }
                __syncwarp();
                const uint32_t nfl = min(rpr, cnt - r0) * n_step * comps;
                float *g = dst + (size_t)comps * (p0 + r0) * n_step;
                for (uint32_t e = lane; e < nfl; e += 32) g[e] = sw[e];
                __syncwarp();
            }
        }
    }
    if (last_block_done(&w.counters[2])) {
        if (threadIdx.x == 0) {
            const int32_t live = w.counters[0];
            w.ctrl[0].n_live = live;
            w.ctrl[0].n_samples = live * (int32_t)n_step;
            w.counters[2] = 0;
        }
    }
}

// composite_rays_triplane (raymarching.cu:2142-2249) over the rays that produced samples; survivors are appended to the next alive list;
// the last CTA to finish publishes the next iteration's control block (n_step = max(min(N // n_alive, 8), 1), renderer.py:506-513) and,
// when the loop is a graph WHILE node, its condition.
__global__ void __launch_bounds__(FR_THREADS) k_frame_composite(float T_thresh, uint32_t N, uint32_t max_steps, FrameWs w, cudaGraphConditionalHandle handle,
                                                                 int use_handle) {
    const FrameCtrl c = w.ctrl[0];
    if (c.done) return;
    const uint32_t n = blockIdx.x * FR_THREADS + threadIdx.x;
    bool survive = false;
    int32_t idx = 0;
    if (n < (uint32_t)c.n_live) {
        const uint32_t n_step = (uint32_t)c.n_step;
        idx = w.alive_mid[n];
        const size_t base = (size_t)n * n_step;
        float t = w.rays_t[idx], ws = w.ws[idx], d = w.depth[idx];
        float r = w.image[3 * (size_t)idx], g = w.image[3 * (size_t)idx + 1], b = w.image[3 * (size_t)idx + 2];
        float a0 = w.aud_sum[idx], a1 = w.eye_sum[idx], u = w.unc_sum[idx];
        uint32_t step = 0;
        while (step < n_step) {
            const size_t i = base + step;
            const float delta = w.deltas[2 * i];
            if (delta == 0.0f) break;
            const float alpha = __fsub_rn(1.0f, __expf(-__fmul_rn(w.sigmas[i], delta)));
            const float T = __fsub_rn(1.0f, ws);
            const float wgt = __fmul_rn(alpha, T);
            ws = __fadd_rn(ws, wgt);
            t = w.deltas[2 * i + 1];
            d = __fmaf_rn(wgt, t, d);
            r = __fmaf_rn(wgt, w.rgbs[3 * i], r); g = __fmaf_rn(wgt, w.rgbs[3 * i + 1], g); b = __fmaf_rn(wgt, w.rgbs[3 * i + 2], b);
            a0 = __fadd_rn(a0, w.amb_aud[i]); a1 = __fadd_rn(a1, w.amb_eye[i]);
            u = __fmaf_rn(wgt, w.unc[i], u);
            if (T < T_thresh) break;
            step++;
        }
        survive = !(step < n_step);
        if (survive) w.rays_t[idx] = t;
        w.ws[idx] = ws; w.depth[idx] = d;
        w.image[3 * (size_t)idx] = r; w.image[3 * (size_t)idx + 1] = g; w.image[3 * (size_t)idx + 2] = b;
        w.aud_sum[idx] = a0; w.eye_sum[idx] = a1; w.unc_sum[idx] = u;
    }
    const uint32_t p = block_append<FR_THREADS>(survive, &w.counters[1]);
    if (survive) w.alive[c.buf ^ 1][p] = idx;
    if (last_block_done(&w.counters[3])) {
        if (threadIdx.x == 0) {
            FrameCtrl nx = {};
            nx.n_alive = w.counters[1];
            nx.step = c.step + c.n_step;
            nx.buf = c.buf ^ 1; nx.iter = c.iter + 1;
            nx.done = (nx.n_alive <= 0 || nx.step >= (int32_t)max_steps) ? 1 : 0;
            if (nx.done) { nx.n_alive = 0; nx.n_step = 1; }
            else {
                int32_t ns = (int32_t)N / nx.n_alive;
                ns = ns < 8 ? ns : 8; ns = ns > 1 ? ns : 1;
                nx.n_step = ns;
            }
            w.ctrl[0] = nx;            // every other CTA has finished, so nobody still reads the old block
            w.counters[0] = 0; w.counters[1] = 0; w.counters[3] = 0;
            if (use_handle) cudaGraphSetConditional(handle, nx.done ? 0u : 1u);
        }
    }
}

// get_rays, all-pixel branch (nerf_triplane/utils.py:227-312) for one pose: pixel n = (row j, column i), direction ((i + 0.5 - cx) / fx,
// (j + 0.5 - cy) / fy, 1) / norm rotated by pose[:3,:3]; origin pose[:3,3].  fp32, one rounding per torch op.
__global__ void __launch_bounds__(256) k_frame_rays(const float *__restrict__ pose, float fx, float fy, float cx, float cy, uint32_t W, uint32_t N,
                                                     float *__restrict__ rays_o, float *__restrict__ rays_d) {
    const float r00 = pose[0], r01 = pose[1], r02 = pose[2], ox = pose[3], r10 = pose[4], r11 = pose[5], r12 = pose[6], oy = pose[7],
                r20 = pose[8], r21 = pose[9], r22 = pose[10], oz = pose[11];
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const uint32_t j = n / W, i = n - j * W;
        const float xs = __fdiv_rn(__fsub_rn(__fadd_rn((float)i, 0.5f), cx), fx), ys = __fdiv_rn(__fsub_rn(__fadd_rn((float)j, 0.5f), cy), fy);
        const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(xs, xs), __fmul_rn(ys, ys)), 1.0f));
        const float dx = __fdiv_rn(xs, nrm), dy = __fdiv_rn(ys, nrm), dz = __fdiv_rn(1.0f, nrm);
        // directions @ R^T: d_k = sum_c dir_c * R[k][c]
        __stcs(rays_d + 3 * (size_t)n, __fmaf_rn(dz, r02, __fmaf_rn(dy, r01, __fmul_rn(dx, r00))));
        __stcs(rays_d + 3 * (size_t)n + 1, __fmaf_rn(dz, r12, __fmaf_rn(dy, r11, __fmul_rn(dx, r10))));
        __stcs(rays_d + 3 * (size_t)n + 2, __fmaf_rn(dz, r22, __fmaf_rn(dy, r21, __fmul_rn(dx, r20))));
        __stcs(rays_o + 3 * (size_t)n, ox); __stcs(rays_o + 3 * (size_t)n + 1, oy); __stcs(rays_o + 3 * (size_t)n + 2, oz);
    }
}
// (image * 255).astype(uint8): truncation (TrainerUtil.py:668); image is already clamped to [0, 1]
__global__ void __launch_bounds__(256) k_image_rgb8(const float *__restrict__ image, uint32_t n3, uint8_t *__restrict__ out) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n3; k += gridDim.x * blockDim.x) out[k] = (uint8_t)(int)__fmul_rn(image[k], 255.0f);
}

// image = clamp(image + (1 - weights_sum) * bg, 0, 1)  (renderer.py:559-561)
__global__ void __launch_bounds__(256) k_frame_finish(uint32_t N, const float *__restrict__ bg, FrameWs w, float *__restrict__ image_out,
                                                       float *__restrict__ ws_out, float *__restrict__ depth_out, uint8_t *__restrict__ rgb8_out) {
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const float ws = w.ws[n];
        const float k = __fsub_rn(1.0f, ws);
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            const float b = bg ? bg[3 * (size_t)n + ch] : 1.0f;
            const float v = __fadd_rn(w.image[3 * (size_t)n + ch], __fmul_rn(k, b));
            const float c = fminf(fmaxf(v, 0.0f), 1.0f);
            __stcs(image_out + 3 * (size_t)n + ch, c);
            if (rgb8_out) rgb8_out[3 * (size_t)n + ch] = (uint8_t)(int)__fmul_rn(c, 255.0f);
        }
        if (ws_out) ws_out[n] = ws;
        if (depth_out) depth_out[n] = w.depth[n];
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct FramePlan {
    const b2n_model *m; b2n_render_cfg cfg; const float *rays_o, *rays_d; uint32_t N; const uint8_t *bitfield;
    const float *enc_a, *ind_code, *eye, *bg; FrameWs w; float *image_out, *ws_out, *depth_out;
    b2n_frame_io io;        // pose == NULL / rgb8_out == NULL: stage not used
};

static int enqueue_init(const FramePlan &p, cudaStream_t st) {
    if (p.io.pose) {
        const uint32_t g0 = ceil_div<uint32_t>(p.N, 256);
        k_frame_rays<<<g0 < (uint32_t)sm_count() * 8 ? g0 : (uint32_t)sm_count() * 8, 256, 0, st>>>(p.io.pose, p.io.fx, p.io.fy, p.io.cx, p.io.cy, p.io.W, p.N,
                                                                                                 const_cast<float *>(p.rays_o), const_cast<float *>(p.rays_d));
        if (check_launch("render_frame(rays)")) return 1;
    }
    // exact empty-space clipping needs whole 32-cell words and a power-of-two grid (Morton blocks); otherwise march unclipped
    const uint32_t H = p.cfg.grid_size;
    const int clip = (H & (H - 1)) == 0 && H >= 4 && ((uintptr_t)p.bitfield & 3) == 0;
    if (clip) {
        k_occ_box<<<OCC_PARTS, 256, 0, st>>>(p.bitfield, p.cfg.cascade, H, p.cfg.bound, p.w.occ_box, 0);
        if (check_launch("render_frame(occupied box)")) return 1;
    }
    const uint32_t sms = (uint32_t)sm_count();
    uint32_t g = ceil_div<uint32_t>(p.N, 256); if (g > sms * 8) g = sms * 8;
    // image-shaped ray sets: width from the pose stage, else from the caller's hint; whole 8 x 16 tiles only
    uint32_t tile_w = p.io.pose ? p.io.W : p.cfg.image_width;
    if (tile_w == 0 || (tile_w & 15u) || p.N % tile_w != 0 || ((p.N / tile_w) & 7u)) tile_w = 0;
    if (const char *e = getenv("B2N_FRAME_TILES")) { if (e[0] == '0') tile_w = 0; }      // A/B
    k_frame_init<<<g, 256, 0, st>>>(p.rays_o, p.rays_d, p.N, p.cfg.min_near, p.cfg.aabb[0], p.cfg.aabb[1], p.cfg.aabb[2], p.cfg.aabb[3], p.cfg.aabb[4], p.cfg.aabb[5],
                                    p.cfg.max_steps, p.cfg.bound, p.cfg.dt_gamma, p.cfg.cascade, H, clip, tile_w, p.w);
    return check_launch("render_frame(init)");
}
// one loop iteration: march (+ drop rays without samples) -> fused head -> composite (+ survivor list, next control block)
static int enqueue_iteration(const FramePlan &p, cudaStream_t st, cudaGraphConditionalHandle handle, int use_handle) {
    const uint32_t ctas = ceil_div<uint32_t>(p.N, FR_THREADS);
    k_frame_march<<<ceil_div<uint32_t>(p.N, FM_THREADS), FM_THREADS, 0, st>>>(p.rays_o, p.rays_d, p.bitfield, p.cfg.bound, p.cfg.dt_gamma, p.cfg.max_steps, p.cfg.cascade, p.cfg.grid_size, p.w);
    if (check_launch("render_frame(march)")) return 1;
    if (int rc = head_forward_on_model(p.m, p.w.xyzs, p.w.dirs, p.N, p.enc_a, p.ind_code, p.eye, &p.w.ctrl[0].n_samples, p.cfg.density_scale, p.w.sigmas, p.w.rgbs,
                                       p.w.amb_aud, p.w.amb_eye, p.w.unc, st, p.w.deltas, nullptr, p.cfg.head_ctas)) return rc;
    k_frame_composite<<<ctas, FR_THREADS, 0, st>>>(p.cfg.T_thresh, p.N, p.cfg.max_steps, p.w, handle, use_handle);
    return check_launch("render_frame(composite)");
}
static int enqueue_finish(const FramePlan &p, cudaStream_t st) {
    const uint32_t sms = (uint32_t)sm_count();
    uint32_t g = ceil_div<uint32_t>(p.N, 256); if (g > sms * 8) g = sms * 8;
    k_frame_finish<<<g, 256, 0, st>>>(p.N, p.bg, p.w, p.image_out, p.ws_out, p.depth_out, p.io.rgb8_out);
    return check_launch("render_frame(finish)");
}

static int make_plan(FramePlan &p, const b2n_model *m, const b2n_render_cfg *cfg, const float *rays_o, const float *rays_d, uint32_t N, const uint8_t *bitfield,
                     const float *enc_a, const float *ind_code, const float *eye, const float *bg_color, void *workspace, float *image_out, float *ws_out, float *depth_out) {
    B2N_REQUIRE(m && cfg && rays_o && rays_d && bitfield && enc_a && workspace && image_out, "render_frame: null pointer");
    B2N_REQUIRE(((uintptr_t)workspace & 255) == 0, "render_frame: workspace must be 256-byte aligned");
    B2N_REQUIRE(cfg->max_steps <= FR_MAX_ITERS, "render_frame: max_steps=%u exceeds the %u-iteration launch plan", cfg->max_steps, FR_MAX_ITERS);
    B2N_REQUIRE(cfg->cascade >= 1 && cfg->cascade <= 24 && cfg->grid_size >= 1 && cfg->grid_size <= 1024, "render_frame: bad cascade / grid size");
    p.m = m; p.cfg = *cfg; p.rays_o = rays_o; p.rays_d = rays_d; p.N = N; p.bitfield = bitfield; p.enc_a = enc_a; p.ind_code = ind_code; p.eye = eye; p.bg = bg_color;
    p.image_out = image_out; p.ws_out = ws_out; p.depth_out = depth_out;
    p.io = b2n_frame_io{};
    carve(&p.w, (uint8_t *)workspace, N);
    return 0;
}

}  // namespace b2n

using namespace b2n;

// A frame as ONE CUDA graph whose loop is a WHILE conditional node: init -> while (!done) { march, head, composite } -> finish.
// The condition is written on the device by the last CTA of k_frame_composite (cudaGraphSetConditional), so exactly the iterations the reference's host loop would
// run are executed, with a single graph launch per frame and no host synchronisation.
struct b2n_frame_graph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    uint64_t kernel_nodes_body = 0, kernel_nodes_fixed = 0;
};

extern "C" {

uint64_t b2n_render_frame_workspace_bytes(uint32_t N) { return (uint64_t)carve(nullptr, nullptr, N); }

int b2n_render_frame(const b2n_model *m, const b2n_render_cfg *cfg, const float *rays_o, const float *rays_d, uint32_t N, const uint8_t *bitfield,
                     const float *enc_a, const float *ind_code, const float *eye, const float *bg_color, void *workspace,
                     float *image_out, float *weights_sum_out, float *depth_out, void *stream) {
    FramePlan p;
    if (int rc = make_plan(p, m, cfg, rays_o, rays_d, N, bitfield, enc_a, ind_code, eye, bg_color, workspace, image_out, weights_sum_out, depth_out)) return rc;
    if (N == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (int rc = enqueue_init(p, st)) return rc;
    for (uint32_t it = 0; it < cfg->max_steps; it++)        // worst case: n_step = 1 every iteration; finished iterations exit at once
        if (int rc = enqueue_iteration(p, st, 0, 0)) return rc;
    return enqueue_finish(p, st);
}

int b2n_get_rays(const float *pose, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W, float *rays_o, float *rays_d, void *stream) {
    B2N_REQUIRE(pose && rays_o && rays_d, "get_rays: null pointer");
    B2N_REQUIRE(fx != 0.0f && fy != 0.0f, "get_rays: zero focal length");
    const uint64_t N64 = (uint64_t)H * W;
    B2N_REQUIRE(N64 <= 0xffffffffull, "get_rays: image too large");
    if (N64 == 0) return 0;
    const uint32_t N = (uint32_t)N64, g0 = ceil_div<uint32_t>(N, 256), cap = (uint32_t)sm_count() * 8;
    k_frame_rays<<<g0 < cap ? g0 : cap, 256, 0, as_stream(stream)>>>(pose, fx, fy, cx, cy, W, N, rays_o, rays_d);
    return check_launch("get_rays");
}

int b2n_image_to_rgb8(const float *image, uint32_t n_pixels, uint8_t *rgb8_out, void *stream) {
    B2N_REQUIRE(image && rgb8_out, "image_to_rgb8: null pointer");
    if (n_pixels == 0) return 0;
    const uint32_t n3 = 3u * n_pixels, g0 = ceil_div<uint32_t>(n3, 256), cap = (uint32_t)sm_count() * 8;
    k_image_rgb8<<<g0 < cap ? g0 : cap, 256, 0, as_stream(stream)>>>(image, n3, rgb8_out);
    return check_launch("image_to_rgb8");
}

int b2n_frame_graph_create(b2n_frame_graph **out, const b2n_model *m, const b2n_render_cfg *cfg, const b2n_audio_weights *audio, const float *auds,
                           uint32_t audio_L, float *enc_a, const float *rays_o, const float *rays_d, uint32_t N, const uint8_t *bitfield,
                           const float *ind_code, const float *eye, const float *bg_color, void *workspace, float *image_out, float *weights_sum_out,
                           float *depth_out) {
    return b2n_frame_graph_create_io(out, m, cfg, audio, auds, audio_L, enc_a, const_cast<float *>(rays_o), const_cast<float *>(rays_d), N, bitfield, ind_code, eye,
                                     bg_color, workspace, image_out, weights_sum_out, depth_out, nullptr);
}

int b2n_frame_graph_create_io(b2n_frame_graph **out, const b2n_model *m, const b2n_render_cfg *cfg, const b2n_audio_weights *audio, const float *auds,
                              uint32_t audio_L, float *enc_a, float *rays_o, float *rays_d, uint32_t N, const uint8_t *bitfield,
                              const float *ind_code, const float *eye, const float *bg_color, void *workspace, float *image_out, float *weights_sum_out,
                              float *depth_out, const b2n_frame_io *io) {
    B2N_REQUIRE(out && N > 0, "frame_graph_create: null pointer / empty frame");
    B2N_REQUIRE(!io || !io->pose || ((uint64_t)io->H * io->W == N && io->fx != 0.0f && io->fy != 0.0f), "frame_graph_create: pose given but H*W != N or zero focal length");
    B2N_REQUIRE(!audio || (auds && enc_a), "frame_graph_create: audio weights given without auds / enc_a buffers");
    FramePlan p;
    if (int rc = make_plan(p, m, cfg, rays_o, rays_d, N, bitfield, enc_a, ind_code, eye, bg_color, workspace, image_out, weights_sum_out, depth_out)) return rc;
    if (io) p.io = *io;
    b2n_frame_graph *fg = new b2n_frame_graph();
    cudaStream_t cs = nullptr;
    int rc = 3;
    const uint64_t l0 = g_launches.load();
    do {
        if (cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess) { set_error("frame_graph_create: stream creation failed"); break; }
        if (cudaGraphCreate(&fg->graph, 0) != cudaSuccess) { set_error("frame_graph_create: cudaGraphCreate failed"); break; }
        // prologue: [audio encoder] + init
        if (cudaStreamBeginCaptureToGraph(cs, fg->graph, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed) != cudaSuccess) { set_error("frame_graph_create: capture-to-graph unsupported"); break; }
        int e = 0;
        if (audio) e = b2n_audio_encode(audio, auds, audio_L, enc_a, cs);
        if (!e) e = enqueue_init(p, cs);
        cudaStreamCaptureStatus status; const cudaGraphNode_t *deps = nullptr; size_t ndeps = 0;
        cudaError_t ce = cudaStreamGetCaptureInfo(cs, &status, nullptr, nullptr, &deps, &ndeps);
        cudaGraphNode_t dep_copy[8]; size_t nd = ndeps < 8 ? ndeps : 8;
        for (size_t i = 0; i < nd; i++) dep_copy[i] = deps[i];
        cudaGraph_t tmp = nullptr;
        cudaError_t ce2 = cudaStreamEndCapture(cs, &tmp);
        if (e || ce != cudaSuccess || ce2 != cudaSuccess) { if (!e) set_error("frame_graph_create: prologue capture failed: %s", cudaGetErrorString(ce != cudaSuccess ? ce : ce2)); break; }
        fg->kernel_nodes_fixed = g_launches.load() - l0;
        // WHILE node
        cudaGraphConditionalHandle handle;
        if ((ce = cudaGraphConditionalHandleCreate(&handle, fg->graph, 1, cudaGraphCondAssignDefault)) != cudaSuccess) { set_error("frame_graph_create: conditional handle: %s", cudaGetErrorString(ce)); break; }
        cudaGraphNodeParams np = {};
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = handle; np.conditional.type = cudaGraphCondTypeWhile; np.conditional.size = 1;
        cudaGraphNode_t loop_node;
        if ((ce = cudaGraphAddNode(&loop_node, fg->graph, dep_copy, nd, &np)) != cudaSuccess) { set_error("frame_graph_create: conditional node: %s", cudaGetErrorString(ce)); break; }
        cudaGraph_t body = np.conditional.phGraph_out[0];
        const uint64_t l1 = g_launches.load();
        if ((ce = cudaStreamBeginCaptureToGraph(cs, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed)) != cudaSuccess) { set_error("frame_graph_create: body capture: %s", cudaGetErrorString(ce)); break; }
        e = enqueue_iteration(p, cs, handle, 1);
        ce2 = cudaStreamEndCapture(cs, &tmp);
        if (e || ce2 != cudaSuccess) { if (!e) set_error("frame_graph_create: body capture failed: %s", cudaGetErrorString(ce2)); break; }
        fg->kernel_nodes_body = g_launches.load() - l1;
        // epilogue
        if ((ce = cudaStreamBeginCaptureToGraph(cs, fg->graph, &loop_node, nullptr, 1, cudaStreamCaptureModeRelaxed)) != cudaSuccess) { set_error("frame_graph_create: epilogue capture: %s", cudaGetErrorString(ce)); break; }
        e = enqueue_finish(p, cs);
        ce2 = cudaStreamEndCapture(cs, &tmp);
        if (e || ce2 != cudaSuccess) { if (!e) set_error("frame_graph_create: epilogue capture failed: %s", cudaGetErrorString(ce2)); break; }
        fg->kernel_nodes_fixed += 1;
        if ((ce = cudaGraphInstantiate(&fg->exec, fg->graph, 0)) != cudaSuccess) { set_error("frame_graph_create: instantiate: %s", cudaGetErrorString(ce)); break; }
        rc = 0;
    } while (0);
    if (cs) cudaStreamDestroy(cs);
    if (rc) {
        (void)cudaGetLastError();
        if (fg->exec) cudaGraphExecDestroy(fg->exec);
        if (fg->graph) cudaGraphDestroy(fg->graph);
        delete fg;
        return rc;
    }
    *out = fg;
    return 0;
}

int b2n_frame_graph_launch(b2n_frame_graph *fg, void *stream) {
    B2N_REQUIRE(fg && fg->exec, "frame_graph_launch: null graph");
    B2N_CUDA(cudaGraphLaunch(fg->exec, as_stream(stream)));
    return 0;
}

/* kernels per loop iteration / outside the loop (for launch accounting: launches per frame = fixed + body * iterations) */
int b2n_frame_graph_info(const b2n_frame_graph *fg, uint64_t *kernels_fixed, uint64_t *kernels_per_iteration) {
    B2N_REQUIRE(fg && kernels_fixed && kernels_per_iteration, "frame_graph_info: null pointer");
    *kernels_fixed = fg->kernel_nodes_fixed; *kernels_per_iteration = fg->kernel_nodes_body;
    return 0;
}

/* number of loop iterations the last frame rendered into `workspace` executed (reads 4 bytes back; synchronises the stream) */
int b2n_frame_iterations(const void *workspace, uint32_t N, int32_t *iterations, void *stream) {
    B2N_REQUIRE(workspace && iterations, "frame_iterations: null pointer");
    FrameWs w; carve(&w, (uint8_t *)const_cast<void *>(workspace), N);
    FrameCtrl c;
    B2N_CUDA(cudaMemcpyAsync(&c, w.ctrl, sizeof(c), cudaMemcpyDeviceToHost, as_stream(stream)));
    B2N_CUDA(cudaStreamSynchronize(as_stream(stream)));
    *iterations = c.iter;
    return 0;
}

void b2n_frame_graph_destroy(b2n_frame_graph *fg) {
    if (!fg) return;
    if (fg->exec) cudaGraphExecDestroy(fg->exec);
    if (fg->graph) cudaGraphDestroy(fg->graph);
    delete fg;
}

}  // extern "C"
