// raymarch.cu — occupancy-grid ray marching and density-grid utilities for sm_100a.
//
// Semantics follow the reference kernels (raymarching/src/raymarching.cu, cited per function); the
// arithmetic is spelled with explicit rounding intrinsics (__fmaf_rn / __fmul_rn / __fadd_rn) in exactly
// the places where nvcc contracted the reference (checked against its sm_100a SASS), so per-ray sample
// counts, voxel indices and sample positions are bit-identical.  What differs is the organisation:
//   * march_rays_train allocates sample segments with a deterministic two-level scan (per-CTA totals +
//     in-CTA scan) instead of two global atomics per ray, so rays[] comes out in ray order and segments
//     of consecutive rays are contiguous — the composite kernels stage them through shared memory.
//   * no FP64: the reference's `0.5 * (x*rb+1) * H` widens to double only by C++ promotion rules; the
//     product is exact in double, so one fp32 multiply with a single rounding is bit-identical.
#include "common.cuh"
#include "dda.cuh"

namespace b2n {

// ---------------------------------------------------------------------------------------------------
// utils
// ---------------------------------------------------------------------------------------------------

// raymarching.cu:92-145 — slab test; miss => both FLT_MAX; near clamped to min_near
__global__ void __launch_bounds__(256) k_near_far(const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                                                   const float *__restrict__ aabb, uint32_t N, float min_near,
                                                   float *__restrict__ nears, float *__restrict__ fars) {
    const float a0 = aabb[0], a1 = aabb[1], a2 = aabb[2], a3 = aabb[3], a4 = aabb[4], a5 = aabb[5];
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const float ox = rays_o[3 * n], oy = rays_o[3 * n + 1], oz = rays_o[3 * n + 2];
        const float rdx = 1.0f / rays_d[3 * n], rdy = 1.0f / rays_d[3 * n + 1], rdz = 1.0f / rays_d[3 * n + 2];
        float tn, tf;
        near_far_one(ox, oy, oz, rdx, rdy, rdz, a0, a1, a2, a3, a4, a5, min_near, tn, tf);
        nears[n] = tn;
        fars[n] = tf;
    }
}

// raymarching.cu:163-198 — background-sphere coordinates
__global__ void __launch_bounds__(256) k_sph_from_ray(const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                                                       float radius, uint32_t N, float *__restrict__ coords) {
    const float RPI = 0.3183098861837907f;
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const float ox = rays_o[3 * n], oy = rays_o[3 * n + 1], oz = rays_o[3 * n + 2];
        const float dx = rays_d[3 * n], dy = rays_d[3 * n + 1], dz = rays_d[3 * n + 2];
        const float A = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
        const float B = __fmaf_rn(oz, dz, __fmaf_rn(oy, dy, __fmul_rn(ox, dx)));
        const float C = __fmaf_rn(-radius, radius, __fmaf_rn(oz, oz, __fmaf_rn(oy, oy, __fmul_rn(ox, ox))));
        const float t = (-B + sqrtf(__fmaf_rn(B, B, -__fmul_rn(A, C)))) / A;
        const float x = __fmaf_rn(t, dx, ox), y = __fmaf_rn(t, dy, oy), z = __fmaf_rn(t, dz, oz);
        const float theta = atan2f(sqrtf(__fmaf_rn(z, z, __fmul_rn(x, x))), y);
        const float phi = atan2f(z, x);
        coords[2 * n] = __fmaf_rn(__fmul_rn(2.0f, theta), RPI, -1.0f);
        coords[2 * n + 1] = __fmul_rn(phi, RPI);
    }
}

// raymarching.cu:214-226 / 237-254
__global__ void __launch_bounds__(256) k_morton3D(const int32_t *__restrict__ coords, uint32_t N, int32_t *__restrict__ indices) {
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x)
        indices[n] = (int32_t)morton_enc((uint32_t)coords[3 * n], (uint32_t)coords[3 * n + 1], (uint32_t)coords[3 * n + 2]);
}
__global__ void __launch_bounds__(256) k_morton3D_invert(const int32_t *__restrict__ indices, uint32_t N, int32_t *__restrict__ coords) {
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const int32_t m = indices[n];
        coords[3 * n] = (int32_t)compact3((uint32_t)(m >> 0));
        coords[3 * n + 1] = (int32_t)compact3((uint32_t)(m >> 1));
        coords[3 * n + 2] = (int32_t)compact3((uint32_t)(m >> 2));
    }
}

// raymarching.cu:268-289 — one thread packs 32 cells (one 128 B line in, one 32-bit word out).
// N is the number of output BYTES; the tail (N % 4) is handled bytewise.
__global__ void __launch_bounds__(256) k_packbits(const float *__restrict__ grid, uint32_t N, float thresh, uint8_t *__restrict__ bitfield) {
    const uint32_t words = N / 4;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < words; w += gridDim.x * blockDim.x) {
        const float4 *g = reinterpret_cast<const float4 *>(grid) + (size_t)w * 8;
        uint32_t bits = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const float4 v = __ldcs(g + q);
            bits |= (uint32_t)(v.x > thresh) << (4 * q);
            bits |= (uint32_t)(v.y > thresh) << (4 * q + 1);
            bits |= (uint32_t)(v.z > thresh) << (4 * q + 2);
            bits |= (uint32_t)(v.w > thresh) << (4 * q + 3);
        }
        reinterpret_cast<uint32_t *>(bitfield)[w] = bits;   // little-endian: byte k of the word = cells 8k..8k+7
    }
    if (blockIdx.x == 0) {
        for (uint32_t n = words * 4 + threadIdx.x; n < N; n += blockDim.x) {
            uint32_t bits = 0;
            for (int i = 0; i < 8; i++) bits |= (uint32_t)(grid[(size_t)n * 8 + i] > thresh) << i;
            bitfield[n] = (uint8_t)bits;
        }
    }
}

// raymarching.cu:304-335 — 6-neighbour max in Morton order.  Neighbours are found IN Morton space: with X = 0x49249249 the x bits of an index,
// x + 1 is ((m | ~X) + 1) & X and x - 1 is ((m & X) - 1) & X (carry / borrow ripple through the foreign bits), so no decode / re-encode per
// neighbour; "x + 1 < H" is a comparison of dilated values (dilation preserves order) against dilate(H).
__global__ void __launch_bounds__(256) k_morton3D_dilation(const float *__restrict__ grid, uint32_t C, uint32_t H, uint32_t dilH, float *__restrict__ out) {
    const uint32_t H3 = H * H * H, total = C * H3;
    constexpr uint32_t X = 0x49249249u;
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < total; n += gridDim.x * blockDim.x) {
        const uint32_t c = n / H3, m = n - c * H3;
        const float *g = grid + (size_t)c * H3;
        float r = __ldg(grid + n);
#pragma unroll
        for (uint32_t a = 0; a < 3; a++) {
            const uint32_t A = X << a;                       // bits of this axis
            const uint32_t ma = m & A, rest = m & ~A;
            const uint32_t up = ((m | ~A) + (1u << a)) & A;  // axis coordinate + 1 (wraps to 0 past the last bit)
            if (up != 0 && (uint64_t)up < ((uint64_t)dilH << a)) r = fmaxf(r, __ldg(g + (rest | up)));
            if (ma != 0) r = fmaxf(r, __ldg(g + (rest | ((ma - (1u << a)) & A))));
        }
        __stcs(out + n, r);
    }
}

// ---------------------------------------------------------------------------------------------------
// training march — raymarching.cu:353-518
// ---------------------------------------------------------------------------------------------------
constexpr int MT_THREADS = 128;

// pass 1: count occupied steps per ray; rays[n] = (n, -, count); per-CTA totals to scratch
__global__ void __launch_bounds__(MT_THREADS) k_march_train_count(
        const float *__restrict__ rays_o, const float *__restrict__ rays_d, const uint8_t *__restrict__ grid,
        float bound, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
        const float *__restrict__ nears, const float *__restrict__ fars, const float *__restrict__ noises,
        int32_t *__restrict__ rays, const int32_t *__restrict__ counter, int32_t *__restrict__ cta_totals, float *__restrict__ t_cache) {
    const uint32_t n = blockIdx.x * MT_THREADS + threadIdx.x;
    uint32_t num = 0;
    if (n < N) {
        DdaRay r;
        r.init(rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n, bound, dt_gamma, max_steps, C, H, fars[n]);
        float t = r.perturb(nears[n], noises[n]);
        DdaSample s;
        // t_cache [max_steps][N] (coalesced over rays): the parameter t of every sample, so that the write pass does not walk the bitfield again —
        // a sample's position and step are functions of (ray, t) alone
        while (t < r.far && num < max_steps) {
            if (r.probe(grid, t, s)) { if (t_cache) t_cache[(size_t)num * N + n] = t; num++; t = __fadd_rn(t, s.dt); }
        }
        rays[3 * (size_t)n] = (int32_t)n;
        rays[3 * (size_t)n + 2] = (int32_t)num;
    }
    // CTA total (warp shuffle + smem)
    __shared__ uint32_t wsum[MT_THREADS / 32];
    uint32_t v = num;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < MT_THREADS / 32; w++) tot += wsum[w];
        cta_totals[1 + blockIdx.x] = (int32_t)tot;
        if (blockIdx.x == 0) cta_totals[0] = counter[0];   // snapshot: offsets start at the counter's current value
    }
}

// pass 2: offsets = snapshot + prefix(CTA totals) + in-CTA exclusive scan; re-march and write samples
__global__ void __launch_bounds__(MT_THREADS) k_march_train_write(
        const float *__restrict__ rays_o, const float *__restrict__ rays_d, const uint8_t *__restrict__ grid,
        float bound, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
        const float *__restrict__ nears, const float *__restrict__ fars, const float *__restrict__ noises,
        float *__restrict__ xyzs, float *__restrict__ dirs, float *__restrict__ deltas,
        int32_t *__restrict__ rays, int32_t *__restrict__ counter, const int32_t *__restrict__ cta_totals, const float *__restrict__ t_cache) {
    __shared__ uint32_t red[MT_THREADS / 32];
    __shared__ uint32_t s_base;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // base = snapshot + sum of totals of CTAs before this one
    uint32_t part = 0;
    for (uint32_t b = threadIdx.x; b < blockIdx.x; b += MT_THREADS) part += (uint32_t)cta_totals[1 + b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t b = (uint32_t)cta_totals[0];
#pragma unroll
        for (int w = 0; w < MT_THREADS / 32; w++) b += red[w];
        s_base = b;
    }
    __syncthreads();
    const uint32_t n = blockIdx.x * MT_THREADS + threadIdx.x;
    const uint32_t num = (n < N) ? (uint32_t)rays[3 * (size_t)n + 2] : 0u;
    // in-CTA exclusive scan of num
    uint32_t inc = num;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    __syncthreads();
    if (lane == 31) red[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < MT_THREADS / 32; w++) if (w < (int)warp) woff += red[w];
    const uint32_t off = s_base + woff + inc - num;
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == MT_THREADS - 1) {
        // this thread's inclusive prefix is the grand total (threads past N contribute 0)
        atomicAdd(counter, (int32_t)(off + num - (uint32_t)cta_totals[0]));
        atomicAdd(counter + 1, (int32_t)N);
    }
    if (n >= N) return;
    rays[3 * (size_t)n + 1] = (int32_t)off;
    if (num == 0 || off + num > M) return;       // raymarching.cu:456-457
    DdaRay r;
    r.init(rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n, bound, dt_gamma, max_steps, C, H, fars[n]);
    float *px = xyzs + 3 * (size_t)off, *pd = dirs + 3 * (size_t)off, *pl = deltas + 2 * (size_t)off;
    if (t_cache) {                               // replay the recorded parameters: same expressions as DdaRay::probe for an occupied cell
        for (uint32_t k = 0; k < num; k++) {
            const float tk = __ldcs(t_cache + (size_t)k * N + n);
            const float dt = r.step_of(tk);
            px[0] = clampf(__fmaf_rn(tk, r.dx, r.ox), -bound, bound);
            px[1] = clampf(__fmaf_rn(tk, r.dy, r.oy), -bound, bound);
            px[2] = clampf(__fmaf_rn(tk, r.dz, r.oz), -bound, bound);
            pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
            pl[0] = dt; pl[1] = __fadd_rn(tk, dt);
            px += 3; pd += 3; pl += 2;
        }
        return;
    }
    float t = r.perturb(nears[n], noises[n]);
    uint32_t step = 0;
    DdaSample s;
    while (t < r.far && step < num) {
        if (r.probe(grid, t, s)) {
            t = __fadd_rn(t, s.dt);
            px[0] = s.x; px[1] = s.y; px[2] = s.z;
            pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
            pl[0] = s.dt; pl[1] = t;
            px += 3; pd += 3; pl += 2; step++;
        }
    }
}

// raymarching.cu:536-583 — gradients to ray origins / directions (camera optimisation)
__global__ void __launch_bounds__(128) k_march_train_backward(
        const float *__restrict__ grad_xyzs, const float *__restrict__ grad_dirs, const int32_t *__restrict__ rays,
        const float *__restrict__ deltas, uint32_t N, uint32_t M, float *__restrict__ grad_rays_o, float *__restrict__ grad_rays_d) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t off = (uint32_t)rays[3 * (size_t)n + 1], num = (uint32_t)rays[3 * (size_t)n + 2];
    if (num == 0 || off + num > M) return;
    float go[3], gd[3];
#pragma unroll
    for (int c = 0; c < 3; c++) { go[c] = grad_rays_o[3 * (size_t)n + c]; gd[c] = grad_rays_d[3 * (size_t)n + c]; }
    for (uint32_t k = 0; k < num; k++) {
        const size_t i = (size_t)off + k;
        const float tk = deltas[2 * i + 1];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float gx = grad_xyzs[3 * i + c];
            go[c] = __fadd_rn(go[c], gx);
            gd[c] = __fadd_rn(gd[c], __fmaf_rn(gx, tk, grad_dirs[3 * i + c]));
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) { grad_rays_o[3 * (size_t)n + c] = go[c]; grad_rays_d[3 * (size_t)n + c] = gd[c]; }
}

// ---------------------------------------------------------------------------------------------------
// inference march — raymarching.cu:828-929
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_march_rays(
        uint32_t n_alive, uint32_t n_step, const int32_t *__restrict__ rays_alive, const float *__restrict__ rays_t,
        const float *__restrict__ rays_o, const float *__restrict__ rays_d, float bound, float dt_gamma,
        uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t *__restrict__ grid,
        const float *__restrict__ fars, float *__restrict__ xyzs, float *__restrict__ dirs, float *__restrict__ deltas,
        const float *__restrict__ noises) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int32_t id = rays_alive[n];
    DdaRay r;
    r.init(rays_o + 3 * (size_t)id, rays_d + 3 * (size_t)id, bound, dt_gamma, max_steps, C, H, fars[id]);
    float t = r.perturb(rays_t[id], noises[n]);
    float *px = xyzs + 3 * (size_t)n * n_step, *pd = dirs + 3 * (size_t)n * n_step, *pl = deltas + 2 * (size_t)n * n_step;
    uint32_t step = 0;
    DdaSample s;
    while (t < r.far && step < n_step) {
        if (r.probe(grid, t, s)) {
            t = __fadd_rn(t, s.dt);
            px[0] = s.x; px[1] = s.y; px[2] = s.z;
            pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
            pl[0] = s.dt; pl[1] = t;
            px += 3; pd += 3; pl += 2; step++;
        }
    }
}

static inline int grid_for(uint32_t n, int threads) {
    const uint32_t want = ceil_div<uint32_t>(n, (uint32_t)threads);
    const uint32_t cap = (uint32_t)sm_count() * 16u;      // grid-stride kernels: a few waves of full SMs
    return (int)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace b2n

using namespace b2n;

extern "C" {

int b2n_near_far_from_aabb(const float *rays_o, const float *rays_d, const float *aabb, uint32_t N, float min_near,
                           float *nears, float *fars, void *stream) {
    B2N_REQUIRE(rays_o && rays_d && aabb && nears && fars, "near_far_from_aabb: null pointer");
    if (N == 0) return 0;
    k_near_far<<<grid_for(N, 256), 256, 0, as_stream(stream)>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
    return check_launch("near_far_from_aabb");
}

int b2n_sph_from_ray(const float *rays_o, const float *rays_d, float radius, uint32_t N, float *coords, void *stream) {
    B2N_REQUIRE(rays_o && rays_d && coords, "sph_from_ray: null pointer");
    if (N == 0) return 0;
    k_sph_from_ray<<<grid_for(N, 256), 256, 0, as_stream(stream)>>>(rays_o, rays_d, radius, N, coords);
    return check_launch("sph_from_ray");
}

int b2n_morton3D(const int32_t *coords, uint32_t N, int32_t *indices, void *stream) {
    B2N_REQUIRE(coords && indices, "morton3D: null pointer");
    if (N == 0) return 0;
    k_morton3D<<<grid_for(N, 256), 256, 0, as_stream(stream)>>>(coords, N, indices);
    return check_launch("morton3D");
}

int b2n_morton3D_invert(const int32_t *indices, uint32_t N, int32_t *coords, void *stream) {
    B2N_REQUIRE(coords && indices, "morton3D_invert: null pointer");
    if (N == 0) return 0;
    k_morton3D_invert<<<grid_for(N, 256), 256, 0, as_stream(stream)>>>(indices, N, coords);
    return check_launch("morton3D_invert");
}

int b2n_packbits(const float *grid, uint32_t N, float density_thresh, uint8_t *bitfield, void *stream) {
    B2N_REQUIRE(grid && bitfield, "packbits: null pointer");
    B2N_REQUIRE(((uintptr_t)grid & 15) == 0 && ((uintptr_t)bitfield & 3) == 0, "packbits: grid must be 16-byte and bitfield 4-byte aligned");
    if (N == 0) return 0;
    k_packbits<<<grid_for(ceil_div<uint32_t>(N, 4u), 256), 256, 0, as_stream(stream)>>>(grid, N, density_thresh, bitfield);
    return check_launch("packbits");
}

int b2n_morton3D_dilation(const float *grid, uint32_t C, uint32_t H, float *grid_dilation, void *stream) {
    B2N_REQUIRE(grid && grid_dilation, "morton3D_dilation: null pointer");
    B2N_REQUIRE(H >= 1 && H <= 1024, "morton3D_dilation: H=%u out of the 10-bit Morton range", H);
    if (C == 0) return 0;
    k_morton3D_dilation<<<grid_for(C * H * H * H, 256), 256, 0, as_stream(stream)>>>(grid, C, H, morton_enc(H, 0, 0), grid_dilation);
    return check_launch("morton3D_dilation");
}

int b2n_march_rays_train(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound, float dt_gamma,
                         uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float *nears,
                         const float *fars, float *xyzs, float *dirs, float *deltas, int32_t *rays, int32_t *counter,
                         const float *noises, void *stream) {
    B2N_REQUIRE(rays_o && rays_d && grid && nears && fars && rays && counter && noises, "march_rays_train: null pointer");
    B2N_REQUIRE(M == 0 || (xyzs && dirs && deltas), "march_rays_train: null output with M=%u", M);
    B2N_REQUIRE(C >= 1 && C <= 24 && H >= 1 && H <= 1024, "march_rays_train: cascade=%u / grid=%u unsupported", C, H);
    if (N == 0) return 0;
    const uint32_t ctas = ceil_div<uint32_t>(N, MT_THREADS);
    int32_t *totals = (int32_t *)scratch(sizeof(int32_t) * (size_t)(ctas + 1), 0);
    B2N_REQUIRE(totals, "march_rays_train: scratch allocation failed");
    // per-sample t cache between the two passes (max_steps x N floats: 4 MB for the 65 536-ray step); large max_steps fall back to re-marching
    float *t_cache = nullptr;
    // (at least 16 MB is requested: the grow-only scratch block then keeps its address for every batch up to 262 144 rays x 16 steps, so a CUDA graph that
    // captured this call is not left pointing at a freed block when a later, larger batch comes along)
    if (max_steps <= 64 && (uint64_t)max_steps * N <= (64ull << 20)) {
        const size_t need = sizeof(float) * (size_t)max_steps * N;
        t_cache = (float *)scratch(need > (16u << 20) ? need : (size_t)(16u << 20), 2);
    }
    k_march_train_count<<<ctas, MT_THREADS, 0, as_stream(stream)>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H,
                                                                      nears, fars, noises, rays, counter, totals, t_cache);
    if (check_launch("march_rays_train(count)")) return 1;
    k_march_train_write<<<ctas, MT_THREADS, 0, as_stream(stream)>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M,
                                                                      nears, fars, noises, xyzs, dirs, deltas, rays, counter, totals, t_cache);
    return check_launch("march_rays_train(write)");
}

int b2n_march_rays_train_backward(const float *grad_xyzs, const float *grad_dirs, const int32_t *rays, const float *deltas,
                                  uint32_t N, uint32_t M, float *grad_rays_o, float *grad_rays_d, void *stream) {
    B2N_REQUIRE(grad_xyzs && grad_dirs && rays && deltas && grad_rays_o && grad_rays_d, "march_rays_train_backward: null pointer");
    if (N == 0) return 0;
    k_march_train_backward<<<ceil_div<uint32_t>(N, 128), 128, 0, as_stream(stream)>>>(grad_xyzs, grad_dirs, rays, deltas, N, M, grad_rays_o, grad_rays_d);
    return check_launch("march_rays_train_backward");
}

int b2n_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t *rays_alive, const float *rays_t, const float *rays_o,
                   const float *rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                   const uint8_t *grid, const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas,
                   const float *noises, void *stream) {
    (void)nears;
    B2N_REQUIRE(rays_alive && rays_t && rays_o && rays_d && grid && fars && xyzs && dirs && deltas && noises, "march_rays: null pointer");
    B2N_REQUIRE(C >= 1 && C <= 24 && H >= 1 && H <= 1024, "march_rays: cascade=%u / grid=%u unsupported", C, H);
    if (n_alive == 0 || n_step == 0) return 0;
    k_march_rays<<<ceil_div<uint32_t>(n_alive, 128), 128, 0, as_stream(stream)>>>(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound,
                                                                                   dt_gamma, max_steps, C, H, grid, fars, xyzs, dirs, deltas, noises);
    return check_launch("march_rays");
}

}  // extern "C"
