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
#include <stdlib.h>
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

// The same dilation for H >= 16 (H a power of two), tiled: 4096 consecutive Morton indices are one aligned 16 x 16 x 16 cube, so a CTA reads its cube with
// contiguous 16-byte loads into an 18^3 shared-memory brick, fetches only the six 16 x 16 halo faces from the neighbouring cubes (0.375 scattered reads per cell
// instead of 6), takes the 6-neighbour max out of shared memory and writes the cube back contiguously.  Cells outside the grid are NaN in the brick:
// fmaxf(v, NaN) = v, exactly the reference's "skip the missing neighbour" (raymarching.cu:326-331), NaN inputs included.
constexpr uint32_t DT = 16, DP = DT + 2;
// shared-memory strides of the brick: the lanes of a warp differ in (z0, x1, y1, z1, x2) of their quad — with the natural 18 / 324 strides four of those
// five bits fall on the same banks.  Row stride 24 (two rows = 16 banks apart) and plane stride 433 (= 1 mod 32) leave at most two lanes per bank.
constexpr uint32_t DSY = 24, DSZ = DP * DSY + 1;
// (Tried: the 2 x 2 x 2 bricks of an aligned 32^3 cube as a thread-block cluster, sibling faces read through distributed shared memory instead of scattered global
// loads — 485 us against 339 us on 8 x 256^3: two cluster barriers and the co-scheduling of eight 31 KB CTAs cost more than the 768 L2 sector reads they save.)
__global__ void __launch_bounds__(256) k_morton3D_dilation_tiled(const float *__restrict__ grid, uint32_t H, float *__restrict__ out) {
    __shared__ float s[DP * DSZ];
    const uint32_t H3 = H * H * H, per_cas = H3 / (DT * DT * DT);
    const uint32_t c = blockIdx.x / per_cas, mb = blockIdx.x - c * per_cas;
    const float *g = grid + (size_t)c * H3;
    const uint32_t bx = compact3(mb) * DT, by = compact3(mb >> 1) * DT, bz = compact3(mb >> 2) * DT;
    const size_t base = (size_t)c * H3 + (size_t)mb * (DT * DT * DT);
    auto at = [&](uint32_t x, uint32_t y, uint32_t z) -> float & { return s[z * DSZ + y * DSY + x]; };      // brick coordinates 0..17 (cell + 1)
    // interior: 1024 float4 = 4 consecutive Morton indices each = cells (x..x+1, y..y+1, z) of one 2 x 2 quad
    for (uint32_t q = threadIdx.x; q < DT * DT * DT / 4; q += 256) {
        const float4 v = __ldcs(reinterpret_cast<const float4 *>(grid + base) + q);
        const uint32_t i = 4 * q, x = compact3(i), y = compact3(i >> 1), z = compact3(i >> 2);
        at(x + 1, y + 1, z + 1) = v.x; at(x + 2, y + 1, z + 1) = v.y; at(x + 1, y + 2, z + 1) = v.z; at(x + 2, y + 2, z + 1) = v.w;
    }
    // halo: 6 faces x 256 cells
    const float QNAN = __int_as_float(0x7fc00000);
    for (uint32_t h = threadIdx.x; h < 6 * DT * DT; h += 256) {
        const uint32_t f = h / (DT * DT), r = h - f * (DT * DT), u = r % DT, v = r / DT;
        const uint32_t axis = f >> 1, hi = f & 1u;
        uint32_t l[3];                                   // brick coordinates
        l[axis] = hi ? DT + 1 : 0; l[(axis + 1) % 3] = u + 1; l[(axis + 2) % 3] = v + 1;
        const int gx = (int)bx + (int)l[0] - 1, gy = (int)by + (int)l[1] - 1, gz = (int)bz + (int)l[2] - 1;
        const bool in = gx >= 0 && gy >= 0 && gz >= 0 && gx < (int)H && gy < (int)H && gz < (int)H;
        at(l[0], l[1], l[2]) = in ? __ldg(g + morton_enc((uint32_t)gx, (uint32_t)gy, (uint32_t)gz)) : QNAN;
    }
    __syncthreads();
    for (uint32_t q = threadIdx.x; q < DT * DT * DT / 4; q += 256) {
        const uint32_t i = 4 * q, x0 = compact3(i) + 1, y0 = compact3(i >> 1) + 1, z = compact3(i >> 2) + 1;
        float r[4];
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            const uint32_t x = x0 + (k & 1u), y = y0 + (k >> 1);
            float m = at(x, y, z);                       // same fmaxf order as the reference: x+1, x-1, y+1, y-1, z+1, z-1
            m = fmaxf(m, at(x + 1, y, z)); m = fmaxf(m, at(x - 1, y, z)); m = fmaxf(m, at(x, y + 1, z)); m = fmaxf(m, at(x, y - 1, z));
            m = fmaxf(m, at(x, y, z + 1)); m = fmaxf(m, at(x, y, z - 1));
            r[k] = m;
        }
        __stcs(reinterpret_cast<float4 *>(out + base) + q, make_float4(r[0], r[1], r[2], r[3]));
    }
}

// ---------------------------------------------------------------------------------------------------
// grown occupied box (exact empty-space clipping, dda.cuh:clip_to_box)
// ---------------------------------------------------------------------------------------------------
// Box around all occupied cells of the bitfield, grown by two cells of the cell's own cascade, in world units (union over cascades).  A 32-bit word of the
// Morton-ordered bitfield is a 4 x 4 x 2 block of cells (Morton bits 0,3 -> x, 1,4 -> y, 2 -> z); a non-empty word is taken whole (a superset is all the
// exactness argument needs).  EXACTNESS: a probe at parameter t reads the cell containing clamp(o + t d, -bound, bound) up to one float ulp, i.e. at most the
// neighbouring cell; a point more than two cells from every occupied cell therefore probes an empty cell.  The clamp maps points outside the scene cube onto
// its faces, so a side of the box that comes within the cube face is opened to infinity (points beyond that face are not clipped); a side that stays inside
// the cube is kept by the clamp (clamp(p)_a stays beyond it whenever p_a is).  The last CTA reduces the partial boxes.
__global__ void __launch_bounds__(256) k_occ_box(const uint8_t *__restrict__ grid, uint32_t C, uint32_t H, float bound, float *__restrict__ parts, int finalize) {
    __shared__ float red[6][8];
    __shared__ int s_last;
    const uint32_t H3 = H * H * H, words = C * H3 / 32;
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (uint32_t wd = blockIdx.x * blockDim.x + threadIdx.x; wd < words; wd += gridDim.x * blockDim.x) {
        const uint32_t bits = __ldg(reinterpret_cast<const uint32_t *>(grid) + wd);
        if (bits) {
            const uint32_t idx = wd * 32, level = idx / H3, m = idx - level * H3;
            const float mb = fminf(scalbnf(1.0f, (int)level), bound);
            const float cell = mb * 2.0f / (float)H;                // cell size of this cascade in world units
            const uint32_t c[3] = {compact3(m), compact3(m >> 1), compact3(m >> 2)};
            const float ext[3] = {3.0f, 3.0f, 1.0f};
#pragma unroll
            for (int a = 0; a < 3; a++) {
                lo[a] = fminf(lo[a], -mb + ((float)c[a] - 2.0f) * cell);
                hi[a] = fmaxf(hi[a], -mb + ((float)c[a] + ext[a] + 3.0f) * cell);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o)); hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o)); }
    if ((threadIdx.x & 31) == 0) for (int a = 0; a < 3; a++) { red[a][threadIdx.x >> 5] = lo[a]; red[3 + a][threadIdx.x >> 5] = hi[a]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = red[threadIdx.x][0];
        for (uint32_t w = 1; w < blockDim.x / 32; w++) v = threadIdx.x < 3 ? fminf(v, red[threadIdx.x][w]) : fmaxf(v, red[threadIdx.x][w]);
        parts[blockIdx.x * 6 + threadIdx.x] = v;
    }
    if (!finalize) return;                        // the consumer reduces the partial boxes itself (k_frame_init)
    int32_t *ticket = reinterpret_cast<int32_t *>(parts + 6 * OCC_PARTS + 6);
    __syncthreads();                              // (one cumulative device-scope fence by the ticket taker, see last_block_done in fused_frame.cu)
    if (threadIdx.x == 0) { __threadfence(); s_last = (atomicAdd(ticket, 1) == (int32_t)gridDim.x - 1); }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < 6 * 32) {                   // warp a reduces component a of the partial boxes
        const uint32_t a = threadIdx.x >> 5, lane = threadIdx.x & 31;
        float v = a < 3 ? 3.0e38f : -3.0e38f;
        for (uint32_t q = lane; q < gridDim.x; q += 32) { const float u = __ldcg(parts + q * 6 + a); v = a < 3 ? fminf(v, u) : fmaxf(v, u); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const float u = __shfl_xor_sync(0xffffffffu, v, o); v = a < 3 ? fminf(v, u) : fmaxf(v, u); }
        if (lane == 0) {
            if (a < 3 && v <= -bound) v = -INFINITY;           // the box reaches the cube face: beyond it clamp() lands on cells that may be occupied
            if (a >= 3 && v >= bound) v = INFINITY;
            parts[6 * OCC_PARTS + a] = v;
        }
    }
    if (threadIdx.x == 0) *ticket = 0;             // re-armed for the next launch on this buffer
}

// ---------------------------------------------------------------------------------------------------
// training march — raymarching.cu:353-518
// ---------------------------------------------------------------------------------------------------
constexpr int MT_THREADS = 128;
__host__ __device__ __forceinline__ uint32_t ceil_div_u(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// pass 1: count occupied steps per ray; rays[n] = (n, -, count); totals of every group of 128 rays to scratch.
// t_cache [N][max_steps]: the parameter t of every sample, so that the write pass does not walk the bitfield again — a sample's position and step are
// functions of (ray, t) alone.
// Training rays are RANDOM pixels (provider.py:669): a third of them reach the occupied box, scattered over every warp, so a thread-per-ray march runs every
// warp for the longest ray at a third of its lanes.  A CTA therefore first clips a tile of 512 rays (cheap, all lanes), compacts the ones that have something to
// march into a dense list in shared memory, and marches those with full warps and four probes in flight per thread (DdaRay::march); counts go back to ray
// order through shared memory.
// MC_GROUPS = groups of MT_THREADS rays per CTA: 4 (a 512-ray tile packs the live rays into full warps) when the batch still fills the machine with CTAs, else 1
template <int MC_GROUPS>
__global__ void __launch_bounds__(MT_THREADS) k_march_train_count(
        const float *__restrict__ rays_o, const float *__restrict__ rays_d, const uint8_t *__restrict__ grid,
        float bound, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
        const float *__restrict__ nears, const float *__restrict__ fars, const float *__restrict__ noises,
        int32_t *__restrict__ rays, const int32_t *__restrict__ counter, int32_t *__restrict__ cta_totals, float *__restrict__ t_cache,
        const float *__restrict__ box, int32_t *__restrict__ ticket) {
    constexpr uint32_t TILE = MC_GROUPS * MT_THREADS;
    __shared__ float s_t0[TILE], s_far[TILE];
    __shared__ uint16_t s_list[TILE];
    __shared__ uint16_t s_num[TILE];
    __shared__ uint32_t s_nlive, s_wcnt[MT_THREADS / 32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tile0 = blockIdx.x * TILE;
    if (threadIdx.x == 0) s_nlive = 0;
    __syncthreads();
#pragma unroll 1
    for (uint32_t g = 0; g < MC_GROUPS; g++) {
        const uint32_t q = g * MT_THREADS + threadIdx.x, n = tile0 + q;
        bool live = false;
        if (n < N) {
            DdaRay r;
            r.init(rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n, bound, dt_gamma, max_steps, C, H, fars[n]);
            float t = r.perturb(nears[n], noises[n]);
            if (box) r.far = r.clip_to_box(box, t);
            live = t < r.far && max_steps > 0;
            s_t0[q] = t; s_far[q] = r.far;
        }
        s_num[q] = 0;
        const uint32_t ballot = __ballot_sync(0xffffffffu, live);
        uint32_t base = 0;
        if (lane == 0 && ballot) base = atomicAdd(&s_nlive, (uint32_t)__popc(ballot));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (live) s_list[base + __popc(ballot & ((1u << lane) - 1u))] = (uint16_t)q;
    }
    __syncthreads();
    const uint32_t n_live = s_nlive;
#pragma unroll 1
    for (uint32_t j = threadIdx.x; j < n_live; j += MT_THREADS) {
        const uint32_t q = s_list[j], m = tile0 + q;
        DdaRay r;
        r.init(rays_o + 3 * (size_t)m, rays_d + 3 * (size_t)m, bound, dt_gamma, max_steps, C, H, s_far[q]);
        float t = s_t0[q];
        float *tc = t_cache ? t_cache + (size_t)m * max_steps : nullptr;
        s_num[q] = (uint16_t)r.march_auto<4>(grid, t, max_steps < 65535u ? max_steps : 65535u, [&](uint32_t k, float tk, float) { if (tc) tc[k] = tk; });
    }
    __syncthreads();
    const uint32_t groups = ceil_div_u(N, MT_THREADS);
#pragma unroll 1
    for (uint32_t g = 0; g < MC_GROUPS; g++) {
        const uint32_t q = g * MT_THREADS + threadIdx.x, n = tile0 + q, gi = blockIdx.x * MC_GROUPS + g;
        if (gi >= groups) break;
        const uint32_t num = s_num[q];
        if (n < N) {
            rays[3 * (size_t)n] = (int32_t)n;
            rays[3 * (size_t)n + 2] = (int32_t)num;
        }
        uint32_t v = num;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if (lane == 0) s_wcnt[warp] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < MT_THREADS / 32; w++) tot += s_wcnt[w];
            cta_totals[1 + gi] = (int32_t)tot;
        }
    }
    // the last CTA to finish turns the group totals into slot offsets: prefix[g] = counter snapshot + samples of all groups before g  (prefix = cta_totals + groups + 1)
    __shared__ int s_last;
    __syncthreads();                              // (one cumulative device-scope fence by the ticket taker, see last_block_done in fused_frame.cu)
    if (threadIdx.x == 0) { __threadfence(); s_last = (atomicAdd(ticket, 1) == (int32_t)gridDim.x - 1); }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    int32_t *prefix = cta_totals + groups + 1;
    // tiles of 4 x 128 consecutive totals (coalesced), scanned in registers / by shuffles, with a running carry
    uint32_t carry = (uint32_t)counter[0];                 // snapshot: offsets start at the counter's current value
    for (uint32_t t0 = 0; t0 < groups; t0 += 4 * MT_THREADS) {
        uint32_t v[4];
#pragma unroll
        for (uint32_t c = 0; c < 4; c++) { const uint32_t b = t0 + 4 * threadIdx.x + c; v[c] = b < groups ? (uint32_t)__ldcg(cta_totals + 1 + b) : 0u; }
        const uint32_t mine = v[0] + v[1] + v[2] + v[3];
        uint32_t inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        __syncthreads();
        if (lane == 31) s_wcnt[warp] = inc;
        __syncthreads();
        uint32_t run = carry + inc - mine, tile_total = 0;
#pragma unroll
        for (int w = 0; w < MT_THREADS / 32; w++) { if (w < (int)warp) run += s_wcnt[w]; tile_total += s_wcnt[w]; }
#pragma unroll
        for (uint32_t c = 0; c < 4; c++) { const uint32_t b = t0 + 4 * threadIdx.x + c; if (b < groups) prefix[b] = (int32_t)run; run += v[c]; }
        carry += tile_total;
    }
    if (threadIdx.x == 0) prefix[groups] = (int32_t)carry;
    if (threadIdx.x == 0) { cta_totals[0] = counter[0]; *ticket = 0; }
}

// Shared prologue of both write kernels: offsets = snapshot + prefix(CTA totals) + in-CTA exclusive scan.  Returns this thread's (offset, count); the last
// thread of the grid publishes the counters.  s_base = the CTA's first slot.
__device__ __forceinline__ void march_train_offsets(uint32_t N, const int32_t *__restrict__ rays, int32_t *__restrict__ counter, const int32_t *__restrict__ cta_totals,
                                                    uint32_t *red, uint32_t &s_base, uint32_t &off, uint32_t &num) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = (uint32_t)cta_totals[gridDim.x + 1 + blockIdx.x];       // prefix[] behind the totals (k_march_train_count's last CTA)
    __syncthreads();
    const uint32_t n = blockIdx.x * MT_THREADS + threadIdx.x;
    num = (n < N) ? (uint32_t)rays[3 * (size_t)n + 2] : 0u;
    uint32_t inc = num;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    __syncthreads();
    if (lane == 31) red[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < MT_THREADS / 32; w++) if (w < (int)warp) woff += red[w];
    off = s_base + woff + inc - num;
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == MT_THREADS - 1) {
        // this thread's inclusive prefix is the grand total (threads past N contribute 0)
        atomicAdd(counter, (int32_t)(off + num - (uint32_t)cta_totals[0]));
        atomicAdd(counter + 1, (int32_t)N);
    }
}

// pass 2 (t_cache present; max_steps <= 64).  The samples of a CTA's 128 consecutive rays form one contiguous slot range [s_base, s_base + total).  The pass is
// SAMPLE-parallel: every ray first marks its slots in a byte map (slot -> ray of the CTA), then thread = slot: it looks up its ray, reads the cached parameter t
// of the sample (neighbouring slots of a ray are neighbouring floats of one t_cache row), evaluates position / step and writes them into a shared-memory tile
// laid out like the output (stride-3 / stride-2 floats over the threads: conflict-free), and the tile leaves as aligned 16-byte stores.  All lanes work and all
// loads of a round are independent — the earlier thread-per-ray version walked rays of 0..16 samples side by side (half the lanes idle, scattered st.shared).
// Same expressions as DdaRay::probe for an occupied cell: bit-identical samples.
constexpr uint32_t MW_TILE = 512;                 // slots staged per round (16 KB of shared memory)
constexpr uint32_t MW_MAX_STEPS = 64;             // = mt_cached()'s bound: the slot map of a CTA is 128 x 64 bytes
__global__ void __launch_bounds__(MT_THREADS) k_march_train_emit(
        const float *__restrict__ rays_o, const float *__restrict__ rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
        float *__restrict__ xyzs, float *__restrict__ dirs, float *__restrict__ deltas,
        int32_t *__restrict__ rays, int32_t *__restrict__ counter, const int32_t *__restrict__ cta_totals, const float *__restrict__ t_cache) {
    __shared__ uint32_t red[MT_THREADS / 32];
    __shared__ uint32_t s_base, s_keep, s_total;
    __shared__ __align__(16) float s_xyz[MW_TILE * 3 + 4], s_dir[MW_TILE * 3 + 4], s_del[MW_TILE * 2 + 4];
    __shared__ uint8_t s_owner[MT_THREADS * MW_MAX_STEPS];
    __shared__ uint32_t s_first[MT_THREADS];                // first slot of the ray (relative to s_base)
    __shared__ float s_od[MT_THREADS][6];
    uint32_t off, num;
    march_train_offsets(N, rays, counter, cta_totals, red, s_base, off, num);
    const uint32_t n = blockIdx.x * MT_THREADS + threadIdx.x;
    if (n < N) {
        rays[3 * (size_t)n + 1] = (int32_t)off;
        if (num) {
#pragma unroll
            for (int c = 0; c < 3; c++) { s_od[threadIdx.x][c] = rays_o[3 * (size_t)n + c]; s_od[threadIdx.x][3 + c] = rays_d[3 * (size_t)n + c]; }
        }
    }
    if (threadIdx.x == 0) s_keep = 0xffffffffu;
    __syncthreads();
    // rays past the buffer are dropped whole (raymarching.cu:456-457); offsets ascend, so the kept slots are a prefix of the CTA's range
    const uint32_t a = off - s_base;
    if (num > 0 && off + num > M) atomicMin(&s_keep, a);
    if (threadIdx.x == MT_THREADS - 1) s_total = a + num;
    s_first[threadIdx.x] = a;
    __syncthreads();
    const uint32_t total = min(s_total, s_keep);
    for (uint32_t k = 0; k < num && a + k < total; k++) s_owner[a + k] = (uint8_t)threadIdx.x;
    DdaRay q;                                              // only the step rule is used
    q.dx = q.dy = q.dz = 1.0f;
    q.init_common(bound, dt_gamma, max_steps, C, H, 0.0f);
    const float *tc = t_cache + (size_t)blockIdx.x * MT_THREADS * max_steps;
    // the tile keeps the 16-byte phase of its destination, so it leaves as aligned 16-byte stores (a round advances every array by a multiple of 16 bytes)
    const uint32_t ph_x = (uint32_t)((reinterpret_cast<uintptr_t>(xyzs + 3 * (size_t)s_base) >> 2) & 3), ph_d = (uint32_t)((reinterpret_cast<uintptr_t>(dirs + 3 * (size_t)s_base) >> 2) & 3),
                   ph_l = (uint32_t)((reinterpret_cast<uintptr_t>(deltas + 2 * (size_t)s_base) >> 2) & 3);
    auto flush = [&](float *dst, const float *src, uint32_t nfl, uint32_t phase) {      // dst[e] = src[phase + e], e < nfl
        const uint32_t head = min((4u - phase) & 3u, nfl), nv = (nfl - head) >> 2, tail = nfl - head - 4 * nv;
        if (threadIdx.x < head) dst[threadIdx.x] = src[phase + threadIdx.x];
        float4 *d4 = reinterpret_cast<float4 *>(dst + head);
        const float4 *s4 = reinterpret_cast<const float4 *>(src + phase + head);
        for (uint32_t v = threadIdx.x; v < nv; v += MT_THREADS) __stcs(d4 + v, s4[v]);
        if (threadIdx.x < tail) dst[head + 4 * nv + threadIdx.x] = src[phase + head + 4 * nv + threadIdx.x];
    };
    __syncthreads();                                       // slot map complete
    for (uint32_t j0 = 0; j0 < total; j0 += MW_TILE) {
        const uint32_t j1 = min(j0 + MW_TILE, total);
#pragma unroll 4
        for (uint32_t sl = j0 + threadIdx.x; sl < j1; sl += MT_THREADS) {
            const uint32_t r = s_owner[sl];
            const float tk = __ldcs(tc + r * max_steps + (sl - s_first[r]));
            const float dt = q.step_of(tk);
            const uint32_t j = sl - j0;
            float *px = s_xyz + ph_x + 3 * j, *pd = s_dir + ph_d + 3 * j, *pl = s_del + ph_l + 2 * j;
            const float d0 = s_od[r][3], d1 = s_od[r][4], d2 = s_od[r][5];
            px[0] = clampf(__fmaf_rn(tk, d0, s_od[r][0]), -bound, bound);
            px[1] = clampf(__fmaf_rn(tk, d1, s_od[r][1]), -bound, bound);
            px[2] = clampf(__fmaf_rn(tk, d2, s_od[r][2]), -bound, bound);
            pd[0] = d0; pd[1] = d1; pd[2] = d2;
            pl[0] = dt; pl[1] = __fadd_rn(tk, dt);
        }
        __syncthreads();
        const uint32_t cnt = j1 - j0;
        const size_t g0 = (size_t)s_base + j0;
        flush(xyzs + 3 * g0, s_xyz, 3 * cnt, ph_x);
        flush(dirs + 3 * g0, s_dir, 3 * cnt, ph_d);
        flush(deltas + 2 * g0, s_del, 2 * cnt, ph_l);
        __syncthreads();
    }
}

// pass 2 without a t cache (very large max_steps): re-march and write samples, one thread per ray
__global__ void __launch_bounds__(MT_THREADS) k_march_train_write(
        const float *__restrict__ rays_o, const float *__restrict__ rays_d, const uint8_t *__restrict__ grid,
        float bound, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
        const float *__restrict__ nears, const float *__restrict__ fars, const float *__restrict__ noises,
        float *__restrict__ xyzs, float *__restrict__ dirs, float *__restrict__ deltas,
        int32_t *__restrict__ rays, int32_t *__restrict__ counter, const int32_t *__restrict__ cta_totals, const float *__restrict__ box) {
    __shared__ uint32_t red[MT_THREADS / 32];
    __shared__ uint32_t s_base;
    uint32_t off, num;
    march_train_offsets(N, rays, counter, cta_totals, red, s_base, off, num);
    const uint32_t n = blockIdx.x * MT_THREADS + threadIdx.x;
    if (n >= N) return;
    rays[3 * (size_t)n + 1] = (int32_t)off;
    if (num == 0 || off + num > M) return;       // raymarching.cu:456-457
    DdaRay r;
    r.init(rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n, bound, dt_gamma, max_steps, C, H, fars[n]);
    float *px = xyzs + 3 * (size_t)off, *pd = dirs + 3 * (size_t)off, *pl = deltas + 2 * (size_t)off;
    float t = r.perturb(nears[n], noises[n]);
    if (box) r.far = r.clip_to_box(box, t);
    r.march_auto<4>(grid, t, num, [&](uint32_t k, float tk, float dt) {
        px[3 * k] = clampf(__fmaf_rn(tk, r.dx, r.ox), -bound, bound); px[3 * k + 1] = clampf(__fmaf_rn(tk, r.dy, r.oy), -bound, bound);
        px[3 * k + 2] = clampf(__fmaf_rn(tk, r.dz, r.oz), -bound, bound);
        pd[3 * k] = r.dx; pd[3 * k + 1] = r.dy; pd[3 * k + 2] = r.dz;
        pl[2 * k] = dt; pl[2 * k + 1] = __fadd_rn(tk, dt);
    });
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
// One thread marches one alive ray (the t recurrence is serial), keeping the <= n_step sample parameters in shared memory; the CTA's 128 rays own the
// contiguous slot range [128 b n_step, 128 (b + 1) n_step), which is then produced slot by slot and written as full lines (unfilled slots as zeros — the
// "ray ended" sentinel of the composite, raymarching.cu:982; the reference relies on a torch.zeros fill for them).
constexpr int MI_THREADS = 128;
constexpr uint32_t MI_MAX_STEP = 8;               // the staged path covers n_step <= 8 (renderer.py:506-513 never asks for more)
__global__ void __launch_bounds__(MI_THREADS) k_march_rays(
        uint32_t n_alive, uint32_t n_step, const int32_t *__restrict__ rays_alive, const float *__restrict__ rays_t,
        const float *__restrict__ rays_o, const float *__restrict__ rays_d, float bound, float dt_gamma,
        uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t *__restrict__ grid,
        const float *__restrict__ fars, float *__restrict__ xyzs, float *__restrict__ dirs, float *__restrict__ deltas,
        const float *__restrict__ noises, const float *__restrict__ box) {
    __shared__ float s_t[MI_THREADS][MI_MAX_STEP + 1];       // +1: conflict-free rows
    __shared__ float s_ray[MI_THREADS][6];
    __shared__ uint32_t s_cnt[MI_THREADS];
    __shared__ float s_xyz[MI_THREADS * 3], s_dir[MI_THREADS * 3], s_del[MI_THREADS * 2];
    const uint32_t n = blockIdx.x * MI_THREADS + threadIdx.x;
    DdaRay r;
    uint32_t step = 0;
    if (n < n_alive) {
        const int32_t id = rays_alive[n];
        r.init(rays_o + 3 * (size_t)id, rays_d + 3 * (size_t)id, bound, dt_gamma, max_steps, C, H, fars[id]);
        float t = r.perturb(rays_t[id], noises[n]);
        if (box) r.far = r.clip_to_box(box, t);
        step = r.march_auto<4>(grid, t, n_step, [&](uint32_t k, float tk, float) { s_t[threadIdx.x][k] = tk; });
        s_ray[threadIdx.x][0] = r.ox; s_ray[threadIdx.x][1] = r.oy; s_ray[threadIdx.x][2] = r.oz;
        s_ray[threadIdx.x][3] = r.dx; s_ray[threadIdx.x][4] = r.dy; s_ray[threadIdx.x][5] = r.dz;
    } else {
        r.dx = r.dy = r.dz = 1.0f;
        r.init_common(bound, dt_gamma, max_steps, C, H, 0.0f);
    }
    s_cnt[threadIdx.x] = step;
    __syncthreads();
    const uint32_t rays_here = min((uint32_t)MI_THREADS, n_alive - blockIdx.x * MI_THREADS);
    const uint32_t slots = rays_here * n_step;
    const size_t g0 = (size_t)blockIdx.x * MI_THREADS * n_step;
    for (uint32_t j0 = 0; j0 < slots; j0 += MI_THREADS) {
        const uint32_t j = j0 + threadIdx.x;
        if (j < slots) {
            const uint32_t ry = j / n_step, k = j - ry * n_step;
            const bool have = k < s_cnt[ry];
            const float tk = have ? s_t[ry][k] : 0.0f;
            const float dt = r.step_of(tk);
            const float *q = s_ray[ry];
            s_xyz[3 * threadIdx.x] = have ? clampf(__fmaf_rn(tk, q[3], q[0]), -bound, bound) : 0.0f;
            s_xyz[3 * threadIdx.x + 1] = have ? clampf(__fmaf_rn(tk, q[4], q[1]), -bound, bound) : 0.0f;
            s_xyz[3 * threadIdx.x + 2] = have ? clampf(__fmaf_rn(tk, q[5], q[2]), -bound, bound) : 0.0f;
            s_dir[3 * threadIdx.x] = have ? q[3] : 0.0f; s_dir[3 * threadIdx.x + 1] = have ? q[4] : 0.0f; s_dir[3 * threadIdx.x + 2] = have ? q[5] : 0.0f;
            s_del[2 * threadIdx.x] = have ? dt : 0.0f; s_del[2 * threadIdx.x + 1] = have ? __fadd_rn(tk, dt) : 0.0f;
        }
        __syncthreads();
        const uint32_t cnt = min((uint32_t)MI_THREADS, slots - j0);
        for (uint32_t e = threadIdx.x; e < 3 * cnt; e += MI_THREADS) { xyzs[3 * (g0 + j0) + e] = s_xyz[e]; dirs[3 * (g0 + j0) + e] = s_dir[e]; }
        for (uint32_t e = threadIdx.x; e < 2 * cnt; e += MI_THREADS) deltas[2 * (g0 + j0) + e] = s_del[e];
        __syncthreads();
    }
}

// n_step > 8 (not reachable from renderer.py): the plain thread-per-ray writer
__global__ void __launch_bounds__(128) k_march_rays_long(
        uint32_t n_alive, uint32_t n_step, const int32_t *__restrict__ rays_alive, const float *__restrict__ rays_t,
        const float *__restrict__ rays_o, const float *__restrict__ rays_d, float bound, float dt_gamma,
        uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t *__restrict__ grid,
        const float *__restrict__ fars, float *__restrict__ xyzs, float *__restrict__ dirs, float *__restrict__ deltas,
        const float *__restrict__ noises, const float *__restrict__ box) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int32_t id = rays_alive[n];
    DdaRay r;
    r.init(rays_o + 3 * (size_t)id, rays_d + 3 * (size_t)id, bound, dt_gamma, max_steps, C, H, fars[id]);
    float t = r.perturb(rays_t[id], noises[n]);
    if (box) r.far = r.clip_to_box(box, t);
    float *px = xyzs + 3 * (size_t)n * n_step, *pd = dirs + 3 * (size_t)n * n_step, *pl = deltas + 2 * (size_t)n * n_step;
    r.march_auto<4>(grid, t, n_step, [&](uint32_t k, float tk, float dt) {
        px[3 * k] = clampf(__fmaf_rn(tk, r.dx, r.ox), -bound, bound); px[3 * k + 1] = clampf(__fmaf_rn(tk, r.dy, r.oy), -bound, bound);
        px[3 * k + 2] = clampf(__fmaf_rn(tk, r.dz, r.oz), -bound, bound);
        pd[3 * k] = r.dx; pd[3 * k + 1] = r.dy; pd[3 * k + 2] = r.dz;
        pl[2 * k] = dt; pl[2 * k + 1] = __fadd_rn(tk, dt);
    });
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
    if (H >= DT && (H & (H - 1)) == 0 && ((uintptr_t)grid & 15) == 0 && ((uintptr_t)grid_dilation & 15) == 0)
        k_morton3D_dilation_tiled<<<C * (H / DT) * (H / DT) * (H / DT), 256, 0, as_stream(stream)>>>(grid, H, grid_dilation);
    else
        k_morton3D_dilation<<<grid_for(C * H * H * H, 256), 256, 0, as_stream(stream)>>>(grid, C, H, morton_enc(H, 0, 0), grid_dilation);
    return check_launch("morton3D_dilation");
}

// workspace layout: [box OCC_BOX_FLOATS floats | cta totals (ctas + 1) int32 | t cache N * max_steps floats (only when max_steps <= 64)]
static inline size_t mt_align(size_t v) { return (v + 255) & ~(size_t)255; }
static inline bool mt_cached(uint32_t N, uint32_t max_steps) { return max_steps <= MW_MAX_STEPS && (uint64_t)max_steps * N <= (64ull << 20); }

uint64_t b2n_march_rays_train_workspace_bytes(uint32_t N, uint32_t max_steps) {
    const uint32_t ctas = ceil_div<uint32_t>(N ? N : 1, MT_THREADS);
    return mt_align(sizeof(float) * OCC_BOX_FLOATS) + mt_align(sizeof(int32_t) * 2 * (size_t)(ctas + 1)) +
           (mt_cached(N, max_steps) ? mt_align(sizeof(float) * (size_t)max_steps * N) : 0);
}

int b2n_march_rays_train_ws(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound, float dt_gamma,
                            uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float *nears,
                            const float *fars, float *xyzs, float *dirs, float *deltas, int32_t *rays, int32_t *counter,
                            const float *noises, void *workspace, void *stream) {
    B2N_REQUIRE(rays_o && rays_d && grid && nears && fars && rays && counter && noises, "march_rays_train: null pointer");
    B2N_REQUIRE(M == 0 || (xyzs && dirs && deltas), "march_rays_train: null output with M=%u", M);
    B2N_REQUIRE(C >= 1 && C <= 24 && H >= 1 && H <= 1024, "march_rays_train: cascade=%u / grid=%u unsupported", C, H);
    B2N_REQUIRE(workspace && ((uintptr_t)workspace & 15) == 0, "march_rays_train: workspace must be a 16-byte aligned device buffer of b2n_march_rays_train_workspace_bytes()");
    if (N == 0) return 0;
    cudaStream_t st = as_stream(stream);
    const uint32_t ctas = ceil_div<uint32_t>(N, MT_THREADS);
    char *ws = (char *)workspace;
    float *box = (float *)ws;
    int32_t *totals = (int32_t *)(ws + mt_align(sizeof(float) * OCC_BOX_FLOATS));
    // per-sample t cache between the two passes ([N][max_steps] floats: 4 MB for the 65 536-ray step); very large max_steps fall back to re-marching
    float *t_cache = mt_cached(N, max_steps) ? (float *)((char *)totals + mt_align(sizeof(int32_t) * 2 * (size_t)(ctas + 1))) : nullptr;
    // exact empty-space clipping needs whole 32-cell words and a power-of-two grid (Morton blocks); otherwise march unclipped
    const bool clip = (H & (H - 1)) == 0 && H >= 4 && ((uintptr_t)grid & 3) == 0;
    B2N_CUDA(cudaMemsetAsync(box + 6 * OCC_PARTS + 6, 0, 2 * sizeof(float), st));       // the two last-CTA tickets (box reduction, offset scan)
    if (clip) {
        k_occ_box<<<OCC_PARTS, 256, 0, st>>>(grid, C, H, bound, box, 1);
        if (check_launch("march_rays_train(box)")) return 1;
    }
    const float *boxp = clip ? box + 6 * OCC_PARTS : nullptr;
    int32_t *ticket = reinterpret_cast<int32_t *>(box + 6 * OCC_PARTS + 7);
    if (ctas >= 8u * (uint32_t)sm_count())
        k_march_train_count<4><<<ceil_div<uint32_t>(ctas, 4u), MT_THREADS, 0, st>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears, fars, noises, rays, counter, totals, t_cache, boxp, ticket);
    else
        k_march_train_count<1><<<ctas, MT_THREADS, 0, st>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears, fars, noises, rays, counter, totals, t_cache, boxp, ticket);
    if (check_launch("march_rays_train(count)")) return 1;
    if (t_cache)
        k_march_train_emit<<<ctas, MT_THREADS, 0, st>>>(rays_o, rays_d, bound, dt_gamma, max_steps, N, C, H, M, xyzs, dirs, deltas, rays, counter, totals, t_cache);
    else
        k_march_train_write<<<ctas, MT_THREADS, 0, st>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, noises, xyzs, dirs, deltas, rays,
                                                         counter, totals, boxp);
    return check_launch("march_rays_train(write)");
}

// The reference's argument list (raymarching.h:18): scratch comes from a library-internal grow-only block per device, so calls on ONE device must be
// stream-ordered with each other (the reference's own kernels all run on the legacy default stream).  Callers that march concurrently on several streams
// (or capture into CUDA graphs that outlive later, larger calls) use b2n_march_rays_train_ws with their own workspace — the Python drop-in does.
int b2n_march_rays_train(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound, float dt_gamma,
                         uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float *nears,
                         const float *fars, float *xyzs, float *dirs, float *deltas, int32_t *rays, int32_t *counter,
                         const float *noises, void *stream) {
    if (N == 0) return 0;
    void *ws = scratch((size_t)b2n_march_rays_train_workspace_bytes(N, max_steps), 0);
    B2N_REQUIRE(ws, "march_rays_train: scratch allocation failed");
    return b2n_march_rays_train_ws(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas, rays, counter, noises, ws, stream);
}

int b2n_march_rays_train_backward(const float *grad_xyzs, const float *grad_dirs, const int32_t *rays, const float *deltas,
                                  uint32_t N, uint32_t M, float *grad_rays_o, float *grad_rays_d, void *stream) {
    B2N_REQUIRE(grad_xyzs && grad_dirs && rays && deltas && grad_rays_o && grad_rays_d, "march_rays_train_backward: null pointer");
    if (N == 0) return 0;
    k_march_train_backward<<<ceil_div<uint32_t>(N, 128), 128, 0, as_stream(stream)>>>(grad_xyzs, grad_dirs, rays, deltas, N, M, grad_rays_o, grad_rays_d);
    return check_launch("march_rays_train_backward");
}

uint64_t b2n_march_rays_workspace_bytes(void) { return mt_align(sizeof(float) * OCC_BOX_FLOATS); }

int b2n_march_rays_ws(uint32_t n_alive, uint32_t n_step, const int32_t *rays_alive, const float *rays_t, const float *rays_o,
                      const float *rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                      const uint8_t *grid, const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas,
                      const float *noises, void *workspace, void *stream) {
    (void)nears;
    B2N_REQUIRE(rays_alive && rays_t && rays_o && rays_d && grid && fars && xyzs && dirs && deltas && noises, "march_rays: null pointer");
    B2N_REQUIRE(C >= 1 && C <= 24 && H >= 1 && H <= 1024, "march_rays: cascade=%u / grid=%u unsupported", C, H);
    if (n_alive == 0 || n_step == 0) return 0;
    cudaStream_t st = as_stream(stream);
    const bool clip = workspace != nullptr && (H & (H - 1)) == 0 && H >= 4 && ((uintptr_t)grid & 3) == 0;
    float *box = (float *)workspace;
    if (clip) {
        B2N_CUDA(cudaMemsetAsync(box + 6 * OCC_PARTS + 6, 0, 2 * sizeof(float), st));
        k_occ_box<<<OCC_PARTS, 256, 0, st>>>(grid, C, H, bound, box, 1);
        if (check_launch("march_rays(box)")) return 1;
    }
    const float *boxp = clip ? box + 6 * OCC_PARTS : nullptr;
    if (n_step <= MI_MAX_STEP)
        k_march_rays<<<ceil_div<uint32_t>(n_alive, MI_THREADS), MI_THREADS, 0, st>>>(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H,
                                                                                     grid, fars, xyzs, dirs, deltas, noises, boxp);
    else
        k_march_rays_long<<<ceil_div<uint32_t>(n_alive, 128), 128, 0, st>>>(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid,
                                                                            fars, xyzs, dirs, deltas, noises, boxp);
    return check_launch("march_rays");
}

// the reference's argument list (raymarching.h:20): library-internal scratch for the occupied box — see b2n_march_rays_train
int b2n_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t *rays_alive, const float *rays_t, const float *rays_o,
                   const float *rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                   const uint8_t *grid, const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas,
                   const float *noises, void *stream) {
    void *ws = scratch((size_t)b2n_march_rays_workspace_bytes(), 1);
    B2N_REQUIRE(ws, "march_rays: scratch allocation failed");
    return b2n_march_rays_ws(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, nears, fars, xyzs, dirs, deltas, noises, ws,
                             stream);
}

}  // extern "C"
