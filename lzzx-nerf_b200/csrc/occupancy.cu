// occupancy.cu — density-grid maintenance of the head model beyond the per-op utilities:
//   * mark_untrained_grid (renderer.py:633-697): cells no training camera sees get density -1 — one kernel over the Morton-ordered grid
//     instead of the reference's 8 x cascade x ceil(B/64) batched matmul / mask / index_put rounds;
//   * update_extra_state's point generation and tail (renderer.py:729-766): the jittered 128^3 lattice is produced directly in MORTON order
//     (so the fused head kernel's sigma output IS tmp_grid — no `tmp_grid[cas, indices] = sigmas` scatter), then ONE kernel does
//     morton3D_dilation + the EMA-max into density_grid + the sum for mean_density, and a second one packs the bitfield against
//     min(mean_density, density_thresh) read on the device (no host read-back between the stages).
#include "common.cuh"

namespace b2n {

// renderer.py:653-691.  One thread = one cell (cascade, Morton index); loops over the B poses until one camera covers the cell.
// Arithmetic in the reference's fp32 op order: world = 2 c / (G - 1) - 1, scaled by (bound_c - half), minus the camera position, times R
// (c2w[:3,:3]; row-vector times matrix), then z > 0, |x| < cx/fx * z + 2 half, |y| < cy/fy * z + 2 half.
__global__ void __launch_bounds__(256) k_mark_untrained(const float *__restrict__ poses, uint32_t B, float kx, float ky, uint32_t cascade, uint32_t G, float bound,
                                                         float *__restrict__ density_grid) {
    extern __shared__ float s_pose[];            // [B][12]: R row-major (9) + t (3)
    for (uint32_t i = threadIdx.x; i < B * 12; i += blockDim.x) {
        const uint32_t b = i / 12, k = i - b * 12;
        s_pose[i] = k < 9 ? poses[b * 16 + (k / 3) * 4 + (k % 3)] : poses[b * 16 + (k - 9) * 4 + 3];
    }
    __syncthreads();
    const uint32_t G3 = G * G * G, total = cascade * G3;
    const float gm1 = (float)(G - 1);
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < total; n += gridDim.x * blockDim.x) {
        const uint32_t cas = n / G3, m = n - cas * G3;
        const float cb = fminf((float)(1u << cas), bound);
        const float half = cb / (float)G;
        const float sc = cb - half, margin = half * 2.0f;
        const float wx = __fmul_rn(__fsub_rn(__fdiv_rn(2.0f * (float)compact3(m), gm1), 1.0f), sc);
        const float wy = __fmul_rn(__fsub_rn(__fdiv_rn(2.0f * (float)compact3(m >> 1), gm1), 1.0f), sc);
        const float wz = __fmul_rn(__fsub_rn(__fdiv_rn(2.0f * (float)compact3(m >> 2), gm1), 1.0f), sc);
        bool seen = false;
        for (uint32_t b = 0; b < B && !seen; b++) {
            const float *p = s_pose + b * 12;
            const float dx = __fsub_rn(wx, p[9]), dy = __fsub_rn(wy, p[10]), dz = __fsub_rn(wz, p[11]);
            const float cx_ = __fmaf_rn(dz, p[6], __fmaf_rn(dy, p[3], __fmul_rn(dx, p[0])));
            const float cy_ = __fmaf_rn(dz, p[7], __fmaf_rn(dy, p[4], __fmul_rn(dx, p[1])));
            const float cz_ = __fmaf_rn(dz, p[8], __fmaf_rn(dy, p[5], __fmul_rn(dx, p[2])));
            seen = cz_ > 0.0f && fabsf(cx_) < __fadd_rn(__fmul_rn(kx, cz_), margin) && fabsf(cy_) < __fadd_rn(__fmul_rn(ky, cz_), margin);
        }
        if (!seen) density_grid[n] = -1.0f;
    }
}

// renderer.py:729-747 for one cascade: Morton index m -> lattice coords -> jittered position, written at row m.
// rand01 [G^3, 3] is indexed in the REFERENCE's point order ((x * G + y) * G + z, the meshgrid order of custom_meshgrid(xs, ys, zs)),
// so a caller that draws torch.rand_like(cas_xyzs) under the same seed reproduces the reference's jitter.
__global__ void __launch_bounds__(256) k_grid_points(const float *__restrict__ rand01, uint32_t G, float sc, float half, float *__restrict__ xyzs) {
    const uint32_t G3 = G * G * G;
    const float gm1 = (float)(G - 1);
    for (uint32_t m = blockIdx.x * blockDim.x + threadIdx.x; m < G3; m += gridDim.x * blockDim.x) {
        const uint32_t c[3] = {compact3(m), compact3(m >> 1), compact3(m >> 2)};
        const size_t r = ((size_t)c[0] * G + c[1]) * G + c[2];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            float v = __fmul_rn(__fsub_rn(__fdiv_rn(2.0f * (float)c[a], gm1), 1.0f), sc);
            if (rand01 != nullptr) v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(__ldcs(rand01 + 3 * r + a), 2.0f), 1.0f), half));
            xyzs[3 * (size_t)m + a] = v;
        }
    }
}

// renderer.py:752-758: tmp = dilate(sigma * density_scale); where (grid >= 0 & tmp >= 0): grid = max(grid * decay, tmp); sum += max(grid, 0).
// The 6-neighbour walk in Morton space is k_morton3D_dilation's (raymarch.cu).
__global__ void __launch_bounds__(256) k_grid_ema(const float *__restrict__ sigma, float density_scale, uint32_t C, uint32_t H, uint32_t dilH, float decay,
                                                   float *__restrict__ density_grid, double *__restrict__ sum) {
    const uint32_t H3 = H * H * H, total = C * H3;
    constexpr uint32_t X = 0x49249249u;
    double acc = 0.0;
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < total; n += gridDim.x * blockDim.x) {
        const uint32_t c = n / H3, m = n - c * H3;
        const float *g = sigma + (size_t)c * H3;
        float r = __ldg(sigma + n);
#pragma unroll
        for (uint32_t a = 0; a < 3; a++) {
            const uint32_t A = X << a;
            const uint32_t ma = m & A, rest = m & ~A;
            const uint32_t up = ((m | ~A) + (1u << a)) & A;
            if (up != 0 && (uint64_t)up < ((uint64_t)dilH << a)) r = fmaxf(r, __ldg(g + (rest | up)));
            if (ma != 0) r = fmaxf(r, __ldg(g + (rest | ((ma - (1u << a)) & A))));
        }
        r = __fmul_rn(r, density_scale);          // max commutes with a positive scale; scale <= 0 is rejected by the entry point
        float d = density_grid[n];
        if (d >= 0.0f && r >= 0.0f) { d = fmaxf(__fmul_rn(d, decay), r); density_grid[n] = d; }
        acc += (double)fmaxf(d, 0.0f);
    }
    __shared__ double s[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; i++) t += s[i];
        atomicAdd(sum, t);
    }
}

// renderer.py:757-766: mean_density = mean(clamp(grid, 0)); bitfield = packbits(grid, min(mean_density, density_thresh)) — threshold formed on the device.
__global__ void __launch_bounds__(256) k_grid_pack(const float *__restrict__ grid, uint32_t n_bytes, const double *__restrict__ sum, double inv_total, float density_thresh,
                                                    uint8_t *__restrict__ bitfield, float *__restrict__ mean_out) {
    const float mean = (float)(sum[0] * inv_total);
    const float thresh = fminf(mean, density_thresh);
    if (blockIdx.x == 0 && threadIdx.x == 0) mean_out[0] = mean;
    const uint32_t words = n_bytes / 4;
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
        reinterpret_cast<uint32_t *>(bitfield)[w] = bits;
    }
}

}  // namespace b2n

using namespace b2n;

static uint32_t stride_grid(uint32_t n, uint32_t threads, uint32_t per_sm) {
    uint32_t g = ceil_div<uint32_t>(n, threads);
    const uint32_t cap = (uint32_t)sm_count() * per_sm;
    return g > cap ? cap : (g == 0 ? 1 : g);
}

static uint32_t dilate_bits(uint32_t H) { return spread3(H); }

extern "C" int b2n_mark_untrained_grid(const float *poses, uint32_t B, float fx, float fy, float cx, float cy, uint32_t cascade, uint32_t grid_size, float bound,
                                       float *density_grid, void *stream) {
    B2N_REQUIRE(poses && density_grid, "mark_untrained_grid: null pointer");
    B2N_REQUIRE(B >= 1 && B <= 1024, "mark_untrained_grid: %u poses per call (1..1024; call again for more — a cell marked -1 by one call and seen by a later one "
                "must be handled by the caller)", B);
    B2N_REQUIRE(cascade >= 1 && cascade <= 8 && grid_size >= 2 && grid_size <= 1024 && (grid_size & (grid_size - 1)) == 0, "mark_untrained_grid: cascade=%u grid_size=%u unsupported",
                cascade, grid_size);
    const uint32_t total = cascade * grid_size * grid_size * grid_size;
    k_mark_untrained<<<stride_grid(total, 256, 8), 256, B * 12 * sizeof(float), as_stream(stream)>>>(poses, B, (float)((double)cx / (double)fx),
                                                                                                    (float)((double)cy / (double)fy), cascade, grid_size, bound, density_grid);
    return check_launch("mark_untrained_grid");
}

extern "C" int b2n_density_grid_points(const float *rand01, uint32_t grid_size, uint32_t cas, float bound, float *xyzs, void *stream) {
    B2N_REQUIRE(xyzs, "density_grid_points: null pointer");
    B2N_REQUIRE(grid_size >= 2 && grid_size <= 1024 && (grid_size & (grid_size - 1)) == 0 && cas < 8, "density_grid_points: grid_size=%u cas=%u unsupported", grid_size, cas);
    const float cb = fminf((float)(1u << cas), bound);
    const float half = cb / (float)grid_size;
    k_grid_points<<<stride_grid(grid_size * grid_size * grid_size, 256, 8), 256, 0, as_stream(stream)>>>(rand01, grid_size, cb - half, half, xyzs);
    return check_launch("density_grid_points");
}

extern "C" int b2n_density_grid_update(const float *sigma, float density_scale, float *density_grid, uint32_t cascade, uint32_t grid_size, float decay, float density_thresh,
                                       uint8_t *bitfield, void *stats, void *stream) {
    B2N_REQUIRE(sigma && density_grid && bitfield && stats, "density_grid_update: null pointer");
    B2N_REQUIRE(cascade >= 1 && cascade <= 8 && grid_size >= 8 && grid_size <= 1024 && (grid_size & (grid_size - 1)) == 0, "density_grid_update: cascade=%u grid_size=%u unsupported",
                cascade, grid_size);
    B2N_REQUIRE(density_scale > 0.0f, "density_grid_update: density_scale must be positive");
    B2N_REQUIRE(((uintptr_t)stats & 7) == 0, "density_grid_update: stats must be 8-byte aligned");
    cudaStream_t st = as_stream(stream);
    const uint32_t total = cascade * grid_size * grid_size * grid_size;
    B2N_CUDA(cudaMemsetAsync(stats, 0, 16, st));
    k_grid_ema<<<stride_grid(total, 256, 8), 256, 0, st>>>(sigma, density_scale, cascade, grid_size, dilate_bits(grid_size), decay, density_grid, (double *)stats);
    if (check_launch("density_grid_update(ema)")) return 1;
    k_grid_pack<<<stride_grid(total / 32, 256, 8), 256, 0, st>>>(density_grid, total / 8, (const double *)stats, 1.0 / (double)total, density_thresh, bitfield,
                                                                 (float *)stats + 2);
    return check_launch("density_grid_update(pack)");
}
