// dda.cuh — the occupancy-bitfield DDA shared by the training marcher, the inference marcher and the fused
// frame kernel.  One probe = one iteration of the reference's while-loop body (raymarching.cu:400-441 /
// :466-517 / :875-928).  Rounding is explicit so the result does not depend on compiler contraction:
// the FMA placements are those of the reference's sm_100a SASS.
#pragma once
#include "common.cuh"

namespace b2n {

struct DdaSample { float x, y, z, dt; };

// raymarching.cu:108-144
__device__ __forceinline__ void near_far_one(float ox, float oy, float oz, float rdx, float rdy, float rdz,
                                             float a0, float a1, float a2, float a3, float a4, float a5,
                                             float min_near, float &tn, float &tf) {
    const float FMAX = 3.402823466e+38f;
    float n0 = __fmul_rn(__fsub_rn(a0, ox), rdx), f0 = __fmul_rn(__fsub_rn(a3, ox), rdx), s;
    if (n0 > f0) { s = n0; n0 = f0; f0 = s; }
    float n1 = __fmul_rn(__fsub_rn(a1, oy), rdy), f1 = __fmul_rn(__fsub_rn(a4, oy), rdy);
    if (n1 > f1) { s = n1; n1 = f1; f1 = s; }
    if (n0 > f1 || n1 > f0) { tn = tf = FMAX; return; }
    if (n1 > n0) n0 = n1;
    if (f1 < f0) f0 = f1;
    float n2 = __fmul_rn(__fsub_rn(a2, oz), rdz), f2 = __fmul_rn(__fsub_rn(a5, oz), rdz);
    if (n2 > f2) { s = n2; n2 = f2; f2 = s; }
    if (n0 > f2 || n2 > f0) { tn = tf = FMAX; return; }
    if (n2 > n0) n0 = n2;
    if (f2 < f0) f0 = f2;
    if (n0 < min_near) n0 = min_near;
    tn = n0; tf = f0;
}

struct DdaRay {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
    float sx, sy, sz;                 // 0.5 * sign(d)
    float bound, rbound, dt_gamma, dt_min, dt_max, rH, Hf, Hm1f, H3f, ncas_m1, far;
    uint32_t C;

    __device__ __forceinline__ void init(const float *o, const float *d, float bound_, float dt_gamma_, uint32_t max_steps,
                                         uint32_t C_, uint32_t H, float far_) {
        ox = o[0]; oy = o[1]; oz = o[2];
        dx = d[0]; dy = d[1]; dz = d[2];
        init_common(bound_, dt_gamma_, max_steps, C_, H, far_);
    }
    __device__ __forceinline__ void init_common(float bound_, float dt_gamma_, uint32_t max_steps, uint32_t C_, uint32_t H, float far_) {
        rdx = 1.0f / dx; rdy = 1.0f / dy; rdz = 1.0f / dz;            // IEEE division (raymarching.cu:378)
        sx = copysignf(0.5f, dx); sy = copysignf(0.5f, dy); sz = copysignf(0.5f, dz);
        bound = bound_; rbound = 1.0f / bound_; dt_gamma = dt_gamma_; C = C_; far = far_;
        Hf = (float)H; Hm1f = (float)(H - 1); rH = 1.0f / Hf;
        H3f = (float)(H * H * H);                                      // `const float H3 = H*H*H` (:380)
        ncas_m1 = __fsub_rn((float)C_, 1.0f);
        dt_max = __fmul_rn(3.4641015529632568359f, (float)(1 << (C_ - 1))) / Hf;   // 2*SQRT3*(1<<(C-1))/H (:386)
        dt_min = fminf(dt_max, 3.4641015529632568359f / (float)max_steps);         // (:387)
    }
    __device__ __forceinline__ float step_of(float t) const { return clampf(__fmul_rn(t, dt_gamma), dt_min, dt_max); }
    // t0 += clamp(t0*dt_gamma, dt_min, dt_max) * noise   (:392 / :873)
    __device__ __forceinline__ float perturb(float t0, float noise) const { return __fmaf_rn(step_of(t0), noise, t0); }

    // mip level of a point / of a step size (raymarching.cu:42-54).  With one cascade the level is 0.
    __device__ __forceinline__ int level_of(float x, float y, float z, float dt) const {
        if (C == 1) return 0;
        int e1, e2;
        (void)frexpf(fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z))), &e1);
        (void)frexpf(__fmul_rn(0.5f, __fmul_rn(dt, Hf)), &e2);
        const int l1 = (int)fminf(ncas_m1, fmaxf(0.0f, (float)e1));
        const int l2 = (int)fminf(ncas_m1, fmaxf(0.0f, (float)e2));
        return l1 > l2 ? l1 : l2;
    }

    // Exact empty-space clipping (see k_occ_box, raymarch.cu).  `box` = the occupied cells of the bitfield grown by two cells, sides that reach the scene cube
    // opened to infinity.  Every t the reference visits is a point of the orbit t <- t + step_of(t) started at t0 — occupied or empty cell alike (a sample
    // advances by dt = step_of(t), the empty-cell do-while repeats the same update) — and outside the grown box every probed cell is empty.  So (a) marching
    // may stop at the box exit, (b) a ray that misses the box has no sample, and (c) the orbit may be run WITHOUT probing (4 dependent flops a step instead of a
    // ~60-instruction probe) up to the first orbit point inside the box: the probes that follow are made at the reference's own t values, bit for bit.
    // Returns the clipped far (0 on a miss) and advances t to the first orbit point >= the box entry.
    __device__ __forceinline__ float clip_to_box(const float *__restrict__ box, float &t) const {
        const float ax0 = (box[0] - ox) * rdx, ax1 = (box[3] - ox) * rdx, ay0 = (box[1] - oy) * rdy, ay1 = (box[4] - oy) * rdy;
        const float az0 = (box[2] - oz) * rdz, az1 = (box[5] - oz) * rdz;
        const float t_in = fmaxf(fmaxf(fminf(ax0, ax1), fminf(ay0, ay1)), fminf(az0, az1));
        const float t_out = fminf(fminf(fmaxf(ax0, ax1), fmaxf(ay0, ay1)), fmaxf(az0, az1));
        if (!(t_in <= t_out)) return 0.0f;                   // miss or empty box (an axis-parallel ray lying IN a box plane, 0 * inf = NaN, also lands here: it is two cells from anything occupied)
        const float f = fminf(far, t_out);
        // entry: t_in is rounded, so stop one step early (the margin of two cells absorbs it: the skipped points are provably outside the occupied cells + 1 cell)
        const float t_stop = t_in - dt_max;
        while (t < t_stop && t < f) t = __fadd_rn(t, step_of(t));
        return f;
    }

    // mip_bound = fminf(scalbnf(1, level), bound) and mip_rbound = 1 / mip_bound (raymarching.cu:412-413) without the division inside the march loop: the
    // reciprocal of a power of two is that power negated, exactly; the other case is 1 / bound, divided once per ray.
    __device__ __forceinline__ void mip_of(int level, float &mb, float &rmb) const {
        const float p2 = __int_as_float((127 + level) << 23);
        if (p2 <= bound) { mb = p2; rmb = __int_as_float((127 - level) << 23); } else { mb = bound; rmb = rbound; }
    }

    // The probe split in two, each a pure function of (ray, t) with the arithmetic of probe() below: the bitfield index of the cell at parameter t, and the
    // empty-cell branch (advance t past the cell's exit).  march<G>() uses them to keep G probes in flight per thread.
    __device__ __forceinline__ uint32_t cell_of(float t) const {
        const float x = clampf(__fmaf_rn(t, dx, ox), -bound, bound);
        const float y = clampf(__fmaf_rn(t, dy, oy), -bound, bound);
        const float z = clampf(__fmaf_rn(t, dz, oz), -bound, bound);
        const int level = level_of(x, y, z, step_of(t));
        float mip_bound, mip_rbound;
        mip_of(level, mip_bound, mip_rbound);
        const float fx = clampf(__fmul_rn(__fmul_rn(0.5f, __fmaf_rn(x, mip_rbound, 1.0f)), Hf), 0.0f, Hm1f);
        const float fy = clampf(__fmul_rn(__fmul_rn(0.5f, __fmaf_rn(y, mip_rbound, 1.0f)), Hf), 0.0f, Hm1f);
        const float fz = clampf(__fmul_rn(__fmul_rn(0.5f, __fmaf_rn(z, mip_rbound, 1.0f)), Hf), 0.0f, Hm1f);
        return (uint32_t)__fmaf_rn((float)level, H3f, (float)morton_enc((uint32_t)(int)fx, (uint32_t)(int)fy, (uint32_t)(int)fz));
    }
    __device__ __forceinline__ void skip_empty(float &t) const {
        const float x = clampf(__fmaf_rn(t, dx, ox), -bound, bound);
        const float y = clampf(__fmaf_rn(t, dy, oy), -bound, bound);
        const float z = clampf(__fmaf_rn(t, dz, oz), -bound, bound);
        const int level = level_of(x, y, z, step_of(t));
        float mip_bound, mip_rbound;
        mip_of(level, mip_bound, mip_rbound);
        const int nx = (int)clampf(__fmul_rn(__fmul_rn(0.5f, __fmaf_rn(x, mip_rbound, 1.0f)), Hf), 0.0f, Hm1f);
        const int ny = (int)clampf(__fmul_rn(__fmul_rn(0.5f, __fmaf_rn(y, mip_rbound, 1.0f)), Hf), 0.0f, Hm1f);
        const int nz = (int)clampf(__fmul_rn(__fmul_rn(0.5f, __fmaf_rn(z, mip_rbound, 1.0f)), Hf), 0.0f, Hm1f);
        const float tx = __fmul_rn(__fmaf_rn(__fmaf_rn(__fmul_rn(__fadd_rn(__fadd_rn((float)nx, 0.5f), sx), rH), 2.0f, -1.0f), mip_bound, -x), rdx);
        const float ty = __fmul_rn(__fmaf_rn(__fmaf_rn(__fmul_rn(__fadd_rn(__fadd_rn((float)ny, 0.5f), sy), rH), 2.0f, -1.0f), mip_bound, -y), rdy);
        const float tz = __fmul_rn(__fmaf_rn(__fmaf_rn(__fmul_rn(__fadd_rn(__fadd_rn((float)nz, 0.5f), sz), rH), 2.0f, -1.0f), mip_bound, -z), rdz);
        const float tt = __fadd_rn(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
        do { t = __fadd_rn(t, step_of(t)); } while (t < tt);
    }

    // ONE-STEP-PER-CELL rays.  In an empty cell the reference advances `do t += step_of(t) while (t < tt)` with tt = t + min(tx, ty, tz) the exit of the cell.
    // Along the axis a with the largest |d_a| the exit is at most one cell away: t_a <= cell |1 / d_a|, so min(tx, ty, tz) <= cell * min_a |1 / d_a|.  If that
    // bound (with 1e-4 of slack for the rounding of the exit arithmetic, which is a few ulps) is <= dt_min <= step_of(t), then tt <= t + step_of(t) — float
    // addition is monotone — and the do-while runs exactly once: an empty cell advances ONE orbit step, like an occupied one.  With the reference's own step
    // rule dt_max = cell diagonal this holds for every ray whose direction is not within a fraction of a degree of a cell diagonal (max |d_a| >= 0.5775),
    // as long as dt_min == dt_max (max_steps <= H / 2^(C-1), the run configuration).  For such a ray the march is a plain walk over the orbit: no exit
    // arithmetic at all, and nothing data-dependent in the control flow but the sample counter.  `cell` is the coarsest cascade's (the largest) cell.
    __device__ __forceinline__ bool one_step_per_cell() const {
        const float mb = fminf(__int_as_float((127 + (int)C - 1) << 23), bound);
        const float cell = __fmul_rn(__fmul_rn(mb, 2.0f), rH);
        return __fmul_rn(__fmul_rn(cell, 1.0001f), fminf(fabsf(rdx), fminf(fabsf(rdy), fabsf(rdz)))) <= dt_min;
    }
    // march<G> for a one-step-per-cell ray: same t sequence and samples as the reference loop
    template <int G, typename F>
    __device__ __forceinline__ uint32_t march_orbit(const uint8_t *__restrict__ grid, float &t, uint32_t max_n, F &&on_sample) const {
        uint32_t num = 0;
        while (t < far && num < max_n) {
            float ts[G];
            uint32_t occ[G];
            ts[0] = t;
#pragma unroll
            for (int j = 1; j < G; j++) ts[j] = __fadd_rn(ts[j - 1], step_of(ts[j - 1]));
#pragma unroll
            for (int j = 0; j < G; j++) { const uint32_t idx = cell_of(ts[j]); occ[j] = (__ldg(grid + (idx >> 3)) >> (idx & 7)) & 1u; }
#pragma unroll
            for (int j = 0; j < G; j++) {
                if (ts[j] < far && num < max_n) {
                    if (occ[j]) { on_sample(num, ts[j], step_of(ts[j])); num++; }
                    t = __fadd_rn(ts[j], step_of(ts[j]));
                }
            }
        }
        return num;
    }
    // dispatch: the cheap walk when the ray qualifies, the general loop otherwise
    template <int G, typename F>
    __device__ __forceinline__ uint32_t march_auto(const uint8_t *__restrict__ grid, float &t, uint32_t max_n, F &&on_sample) const {
        return one_step_per_cell() ? march_orbit<G>(grid, t, max_n, on_sample) : march<G>(grid, t, max_n, on_sample);
    }

    // The reference's marching loop (`while (t < far && step < max_n) { probe }`, raymarching.cu:400-441) with G probes in flight per thread.  Every t the loop
    // visits lies on the orbit t <- t + step_of(t), so the next G - 1 orbit points are computed ahead and their cells fetched together (G independent index
    // computations and bitfield loads instead of G dependent round trips); the loop is then replayed serially over the fetched bits, and a fetched bit is used
    // only if the loop really arrives at that t (an empty cell may jump over orbit points: the group is then abandoned and refetched from the true t).  Same
    // t sequence, same samples, bit for bit; on_sample(k, t, dt) is called for the k-th sample.  Returns the number of samples; t = where the loop stopped.
    template <int G, typename F>
    __device__ __forceinline__ uint32_t march(const uint8_t *__restrict__ grid, float &t, uint32_t max_n, F &&on_sample) const {
        uint32_t num = 0;
        while (t < far && num < max_n) {
            float ts[G];
            uint32_t occ[G];
            ts[0] = t;
#pragma unroll
            for (int j = 1; j < G; j++) ts[j] = __fadd_rn(ts[j - 1], step_of(ts[j - 1]));
#pragma unroll
            for (int j = 0; j < G; j++) { const uint32_t idx = cell_of(ts[j]); occ[j] = (__ldg(grid + (idx >> 3)) >> (idx & 7)) & 1u; }
#pragma unroll
            for (int j = 0; j < G; j++) {
                if (ts[j] != t || !(t < far) || num >= max_n) break;
                if (occ[j]) { const float dt = step_of(t); on_sample(num, t, dt); num++; t = __fadd_rn(t, dt); }
                else skip_empty(t);
            }
        }
        return num;
    }

    // One loop iteration at parameter t.  Occupied: fills s, returns true (caller advances t by s.dt).
    // Empty: advances t past the voxel exit with the reference's do-while and returns false.
    __device__ __forceinline__ bool probe(const uint8_t *__restrict__ grid, float &t, DdaSample &s) const {
        const float x = clampf(__fmaf_rn(t, dx, ox), -bound, bound);
        const float y = clampf(__fmaf_rn(t, dy, oy), -bound, bound);
        const float z = clampf(__fmaf_rn(t, dz, oz), -bound, bound);
        const float dt = step_of(t);
        const int level = level_of(x, y, z, dt);
        float mip_bound, mip_rbound;
        mip_of(level, mip_bound, mip_rbound);
        // 0.5 * (x * mip_rbound + 1) * H — exact in the reference's double, so one fp32 rounding here
        const float fx = clampf(__fmul_rn(__fmul_rn(0.5f, __fmaf_rn(x, mip_rbound, 1.0f)), Hf), 0.0f, Hm1f);
        const float fy = clampf(__fmul_rn(__fmul_rn(0.5f, __fmaf_rn(y, mip_rbound, 1.0f)), Hf), 0.0f, Hm1f);
        const float fz = clampf(__fmul_rn(__fmul_rn(0.5f, __fmaf_rn(z, mip_rbound, 1.0f)), Hf), 0.0f, Hm1f);
        const int nx = (int)fx, ny = (int)fy, nz = (int)fz;
        const uint32_t index = (uint32_t)__fmaf_rn((float)level, H3f, (float)morton_enc((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
        const uint32_t occ = (__ldg(grid + (index >> 3)) >> (index & 7)) & 1u;
        if (occ) { s.x = x; s.y = y; s.z = z; s.dt = dt; return true; }
        // distance to the voxel exit (:431-435)
        const float tx = __fmul_rn(__fmaf_rn(__fmaf_rn(__fmul_rn(__fadd_rn(__fadd_rn((float)nx, 0.5f), sx), rH), 2.0f, -1.0f), mip_bound, -x), rdx);
        const float ty = __fmul_rn(__fmaf_rn(__fmaf_rn(__fmul_rn(__fadd_rn(__fadd_rn((float)ny, 0.5f), sy), rH), 2.0f, -1.0f), mip_bound, -y), rdy);
        const float tz = __fmul_rn(__fmaf_rn(__fmaf_rn(__fmul_rn(__fadd_rn(__fadd_rn((float)nz, 0.5f), sz), rH), 2.0f, -1.0f), mip_bound, -z), rdz);
        const float tt = __fadd_rn(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
        do { t = __fadd_rn(t, step_of(t)); } while (t < tt);
        return false;
    }
};

// Grown occupied box of a bitfield (k_occ_box, raymarch.cu): OCC_PARTS per-CTA partial boxes, then the final [lo xyz | hi xyz] at parts + 6 * OCC_PARTS
// and a ticket counter behind it.
constexpr uint32_t OCC_PARTS = 128;
constexpr uint32_t OCC_BOX_FLOATS = 6 * OCC_PARTS + 6 + 2;
__global__ void k_occ_box(const uint8_t *__restrict__ grid, uint32_t C, uint32_t H, float bound, float *__restrict__ parts, int finalize);

}  // namespace b2n
