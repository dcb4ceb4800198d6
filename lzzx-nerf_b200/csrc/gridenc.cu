// gridenc.cu — multiresolution hash / tiled grid encoder (forward, table-gradient scatter, input gradient).
//
// Semantics: gridencoder/src/gridencoder.cu:35-342 (index function :54-72, interpolation :124-175, input
// derivative :179-222, table backward :264-312, input backward :317-342).  The per-level scale uses the same
// ex2.approx + fma as the reference (`exp2f(level*S)*H - 1`), corners are accumulated in the reference's
// order with one fma each, so fp32 outputs are bit-identical on the same GPU.
//
// B200 organisation:
//   * D == 2 (the tri-plane and torso grids): the two x-neighbours of a cell are adjacent in memory for dense
//     levels, and for hashed levels too because the first hash prime is 1 and table sizes are powers of two
//     (idx(x+1,y) == idx(x,y) ^ 1 when x is even).  The kernel detects `i1 == (i0 ^ 1)` and fetches the
//     aligned pair with ONE vector load (8 B fp32 / 4 B fp16 at C=1), i.e. <=3 loads instead of 4 per level.
//   * tables (<= 2 MB for the head model) stay L2/L1 resident: table loads use the read-only path with default
//     caching, the streamed inputs/outputs use evict-first (__ldcs/__stcs) so they do not displace the tables.
//   * backward: the same pair detection turns two scalar REDs into one `red.global.add.v2.f32` (sm_90+ vector
//     reduction) — the scatter is bound by the SM's RED issue rate, so halving the instruction count is the lever.
#include "common.cuh"
#include <mutex>
#include <cstdlib>

namespace b2n {

__constant__ uint32_t c_primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};

template <uint32_t D>
__device__ __forceinline__ uint32_t grid_slot(uint32_t gridtype, bool align_corners, uint32_t hashmap_size, uint32_t resolution, const uint32_t (&pg)[D]) {
    uint32_t stride = 1, index = 0;
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        if (stride <= hashmap_size) {
            index += pg[d] * stride;
            stride *= align_corners ? resolution : (resolution + 1);
        }
    }
    if (gridtype == 0 && stride > hashmap_size) {
        index = 0;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) index ^= pg[d] * c_primes[d];
    }
    return index % hashmap_size;
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// acc += w * val with the reference's rounding: fp32 = one fma; fp16 (scalar_t = at::Half) = the product is
// narrowed to half, added in float, narrowed again (c10::Half operator+= semantics)
template <typename T> __device__ __forceinline__ void acc_corner(float &acc, float w, float val);
template <> __device__ __forceinline__ void acc_corner<float>(float &acc, float w, float val) { acc = __fmaf_rn(w, val, acc); }
template <> __device__ __forceinline__ void acc_corner<__half>(float &acc, float w, float val) {
    const float p = __half2float(__float2half_rn(__fmul_rn(w, val)));
    acc = __half2float(__float2half_rn(__fadd_rn(acc, p)));
}

struct LevelGeom { float scale; uint32_t resolution, hashmap_size, table_off; };
__device__ __forceinline__ LevelGeom level_geom(const int32_t *__restrict__ offsets, uint32_t level, float S, uint32_t H) {
    LevelGeom g;
    g.table_off = (uint32_t)offsets[level];
    g.hashmap_size = (uint32_t)offsets[level + 1] - g.table_off;
    g.scale = __fmaf_rn(exp2f(__fmul_rn((float)level, S)), (float)H, -1.0f);     // gridencoder.cu:125
    g.resolution = (uint32_t)ceilf(g.scale) + 1u;                               // :126
    return g;
}

// One level of one sample: cell, fractional position and the interpolated C features (gridencoder.cu:124-175); shared by the level-major kernel of the
// reference layout and the row-major kernel below.
template <typename T, uint32_t D, uint32_t C>
__device__ __forceinline__ void grid_level_eval(const float (&in)[D], const LevelGeom &g, const T *__restrict__ tab, uint32_t gridtype, bool align_corners,
                                                float (&res)[C], float (&pos)[D], uint32_t (&pg)[D]) {
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        pos[d] = __fmaf_rn(in[d], g.scale, align_corners ? 0.0f : 0.5f);
        const float fl = floorf(pos[d]);
        pg[d] = (uint32_t)fl;
        pos[d] = __fsub_rn(pos[d], (float)pg[d]);
    }
#pragma unroll
    for (uint32_t c = 0; c < C; c++) res[c] = 0.0f;

    if constexpr (D == 2 && C == 1) {
        // paired-corner fast path: corners (0,1) and (2,3) differ only in x
        const float wx0 = __fsub_rn(1.0f, pos[0]), wx1 = pos[0];
#pragma unroll
        for (uint32_t j = 0; j < 2; j++) {
            const float wy = j ? pos[1] : __fsub_rn(1.0f, pos[1]);
            const uint32_t p0[2] = {pg[0], pg[1] + j}, p1[2] = {pg[0] + 1, pg[1] + j};
            const uint32_t i0 = grid_slot<D>(gridtype, align_corners, g.hashmap_size, g.resolution, p0);
            const uint32_t i1 = grid_slot<D>(gridtype, align_corners, g.hashmap_size, g.resolution, p1);
            float v0, v1;
            if (i1 == (i0 ^ 1u) && !(g.table_off & 1u)) {
                // aligned pair (table offsets are multiples of 8 entries, grid.py:117)
                if (sizeof(T) == 4) {
                    const float2 pr = __ldg(reinterpret_cast<const float2 *>(tab) + (i0 >> 1));
                    v0 = (i0 & 1u) ? pr.y : pr.x; v1 = (i0 & 1u) ? pr.x : pr.y;
                } else {
                    const __half2 pr = __ldg(reinterpret_cast<const __half2 *>(tab) + (i0 >> 1));
                    v0 = __half2float((i0 & 1u) ? pr.y : pr.x); v1 = __half2float((i0 & 1u) ? pr.x : pr.y);
                }
            } else {
                v0 = to_f<T>(__ldg(tab + i0)); v1 = to_f<T>(__ldg(tab + i1));
            }
            acc_corner<T>(res[0], __fmul_rn(wx0, wy), v0);
            acc_corner<T>(res[0], __fmul_rn(wx1, wy), v1);
        }
    } else {
#pragma unroll
        for (uint32_t idx = 0; idx < (1u << D); idx++) {
            float w = 1.0f;
            uint32_t pl[D];
#pragma unroll
            for (uint32_t d = 0; d < D; d++) {
                if ((idx & (1u << d)) == 0) { w = __fmul_rn(w, __fsub_rn(1.0f, pos[d])); pl[d] = pg[d]; }
                else                        { w = __fmul_rn(w, pos[d]);                 pl[d] = pg[d] + 1; }
            }
            const size_t e = (size_t)grid_slot<D>(gridtype, align_corners, g.hashmap_size, g.resolution, pl) * C;
#pragma unroll
            for (uint32_t c = 0; c < C; c++) acc_corner<T>(res[c], w, to_f<T>(__ldg(tab + e + c)));
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// forward — one thread per (sample, level); outputs [L,B,C]
// ---------------------------------------------------------------------------------------------------
template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) k_grid_fwd(const float *__restrict__ inputs, const T *__restrict__ table, const int32_t *__restrict__ offsets,
                                                   T *__restrict__ outputs, uint32_t B, uint32_t L, float S, uint32_t H, T *__restrict__ dy_dx,
                                                   uint32_t gridtype, bool align_corners) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    const LevelGeom g = level_geom(offsets, level, S, H);
    const T *tab = table + (size_t)g.table_off * C;
    T *out = outputs + ((size_t)level * B + b) * C;

    float in[D];
    bool oob = false;
#pragma unroll
    for (uint32_t d = 0; d < D; d++) { in[d] = __ldcs(inputs + (size_t)b * D + d); oob |= (in[d] < 0.0f || in[d] > 1.0f); }
    if (oob) {                                                                   // gridencoder.cu:98-122
#pragma unroll
        for (uint32_t c = 0; c < C; c++) out[c] = from_f<T>(0.0f);
        if (dy_dx) {
            T *dd = dy_dx + (size_t)b * D * L * C + (size_t)level * D * C;
#pragma unroll
            for (uint32_t k = 0; k < D * C; k++) dd[k] = from_f<T>(0.0f);
        }
        return;
    }
    float pos[D], res[C];
    uint32_t pg[D];
    grid_level_eval<T, D, C>(in, g, tab, gridtype, align_corners, res, pos, pg);
#pragma unroll
    for (uint32_t c = 0; c < C; c++) out[c] = from_f<T>(res[c]);

    if (!dy_dx) return;
    T *dd = dy_dx + (size_t)b * D * L * C + (size_t)level * D * C;               // [B, L, D, C]  (:181)
#pragma unroll
    for (uint32_t gd = 0; gd < D; gd++) {
        float rg[C];
#pragma unroll
        for (uint32_t c = 0; c < C; c++) rg[c] = 0.0f;
#pragma unroll
        for (uint32_t idx = 0; idx < (1u << (D - 1)); idx++) {
            float w = g.scale;
            uint32_t pl[D];
#pragma unroll
            for (uint32_t nd = 0; nd + 1 < D; nd++) {
                const uint32_t d = (nd >= gd) ? nd + 1 : nd;
                if ((idx & (1u << nd)) == 0) { w = __fmul_rn(w, __fsub_rn(1.0f, pos[d])); pl[d] = pg[d]; }
                else                         { w = __fmul_rn(w, pos[d]);                 pl[d] = pg[d] + 1; }
            }
            pl[gd] = pg[gd];
            const size_t el = (size_t)grid_slot<D>(gridtype, align_corners, g.hashmap_size, g.resolution, pl) * C;
            pl[gd] = pg[gd] + 1;
            const size_t er = (size_t)grid_slot<D>(gridtype, align_corners, g.hashmap_size, g.resolution, pl) * C;
#pragma unroll
            for (uint32_t c = 0; c < C; c++) {
                // Half - Half narrows to half in the reference (scalar_t arithmetic)
                const float diff = to_f<T>(from_f<T>(__fsub_rn(to_f<T>(__ldg(tab + er + c)), to_f<T>(__ldg(tab + el + c)))));
                acc_corner<T>(rg[c], w, diff);
            }
        }
#pragma unroll
        for (uint32_t c = 0; c < C; c++) dd[gd * C + c] = from_f<T>(rg[c]);
    }
}

// ---------------------------------------------------------------------------------------------------
// forward, row-major output [B, L*C] — what GridEncoder.forward hands to its caller (grid.py:52 permutes the reference's [L,B,C] with a transposing copy of the
// whole output).  One thread evaluates ALL levels of its sample (the tables of the head / torso models sit in L1/L2, so the reference's level-major launch that
// keeps one level's table hot buys nothing on a 126 MB L2); the CTA's 128 rows are staged in shared memory and leave as one contiguous, fully coalesced block.
// Same per-level arithmetic (grid_level_eval) => same bits as the level-major kernel.
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t GR_THREADS = 128;
template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(GR_THREADS) k_grid_fwd_rows(const float *__restrict__ inputs, const T *__restrict__ table, const int32_t *__restrict__ offsets,
                                                               T *__restrict__ outputs, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align_corners) {
    extern __shared__ __align__(16) uint8_t gr_smem[];
    LevelGeom *s_lvl = reinterpret_cast<LevelGeom *>(gr_smem);
    T *s_out = reinterpret_cast<T *>(gr_smem + sizeof(LevelGeom) * ((L + 3) & ~3u));
    const uint32_t LC = L * C, pitch = LC | 1u;                      // odd pitch: conflict-free column writes
    for (uint32_t l = threadIdx.x; l < L; l += GR_THREADS) s_lvl[l] = level_geom(offsets, l, S, H);
    __syncthreads();
    const uint32_t b = blockIdx.x * GR_THREADS + threadIdx.x;
    if (b < B) {
        float in[D];
        bool oob = false;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) { in[d] = __ldcs(inputs + (size_t)b * D + d); oob |= (in[d] < 0.0f || in[d] > 1.0f); }
        T *row = s_out + threadIdx.x * pitch;
        if (oob) {                                                   // gridencoder.cu:98-122
            for (uint32_t k = 0; k < LC; k++) row[k] = from_f<T>(0.0f);
        } else {
#pragma unroll 2
            for (uint32_t l = 0; l < L; l++) {
                const LevelGeom g = s_lvl[l];
                float pos[D], res[C];
                uint32_t pg[D];
                grid_level_eval<T, D, C>(in, g, table + (size_t)g.table_off * C, gridtype, align_corners, res, pos, pg);
#pragma unroll
                for (uint32_t c = 0; c < C; c++) row[l * C + c] = from_f<T>(res[c]);
            }
        }
    }
    __syncthreads();
    const uint32_t rows_here = min(GR_THREADS, B - blockIdx.x * GR_THREADS), total = rows_here * LC;
    T *dst = outputs + (size_t)blockIdx.x * GR_THREADS * LC;
    for (uint32_t e = threadIdx.x; e < total; e += GR_THREADS) { const uint32_t r = e / LC; dst[e] = s_out[r * pitch + (e - r * LC)]; }
}

// ---------------------------------------------------------------------------------------------------
// backward: table gradient scatter — one thread per (sample, level); grad [L,B,C]
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_f32(float *addr, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void red_add_v2_f32(float *addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_h2(__half *addr, __half2 v) {
    asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(addr), "r"(*reinterpret_cast<uint32_t *>(&v)) : "memory");
}

template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) k_grid_bwd(const T *__restrict__ grad, const float *__restrict__ inputs, const int32_t *__restrict__ offsets,
                                                   T *__restrict__ grad_table, uint32_t B, uint32_t L, float S, uint32_t H,
                                                   uint32_t gridtype, bool align_corners) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    const LevelGeom g = level_geom(offsets, level, S, H);
    T *gtab = grad_table + (size_t)g.table_off * C;

    float in[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) { in[d] = __ldcs(inputs + (size_t)b * D + d); if (in[d] < 0.0f || in[d] > 1.0f) return; }
    float pos[D];
    uint32_t pg[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        pos[d] = __fmaf_rn(in[d], g.scale, align_corners ? 0.0f : 0.5f);
        pg[d] = (uint32_t)floorf(pos[d]);
        pos[d] = __fsub_rn(pos[d], (float)pg[d]);
    }
    float gc[C];
#pragma unroll
    for (uint32_t c = 0; c < C; c++) gc[c] = to_f<T>(grad[((size_t)level * B + b) * C + c]);

    if constexpr (D == 2 && C == 1 && sizeof(T) == 4) {
        const float wx0 = __fsub_rn(1.0f, pos[0]), wx1 = pos[0];
#pragma unroll
        for (uint32_t j = 0; j < 2; j++) {
            const float wy = j ? pos[1] : __fsub_rn(1.0f, pos[1]);
            const uint32_t p0[2] = {pg[0], pg[1] + j}, p1[2] = {pg[0] + 1, pg[1] + j};
            const uint32_t i0 = grid_slot<D>(gridtype, align_corners, g.hashmap_size, g.resolution, p0);
            const uint32_t i1 = grid_slot<D>(gridtype, align_corners, g.hashmap_size, g.resolution, p1);
            const float g0 = __fmul_rn(__fmul_rn(wx0, wy), gc[0]), g1 = __fmul_rn(__fmul_rn(wx1, wy), gc[0]);
            float *base = reinterpret_cast<float *>(gtab);
            if (i1 == (i0 ^ 1u) && !(g.table_off & 1u)) {
                if (i0 & 1u) red_add_v2_f32(base + i1, g1, g0); else red_add_v2_f32(base + i0, g0, g1);
            } else {
                red_add_f32(base + i0, g0); red_add_f32(base + i1, g1);
            }
        }
    } else {
#pragma unroll
    for (uint32_t idx = 0; idx < (1u << D); idx++) {
        float w = 1.0f;
        uint32_t pl[D];
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            if ((idx & (1u << d)) == 0) { w = __fmul_rn(w, __fsub_rn(1.0f, pos[d])); pl[d] = pg[d]; }
            else                        { w = __fmul_rn(w, pos[d]);                 pl[d] = pg[d] + 1; }
        }
        const size_t e = (size_t)grid_slot<D>(gridtype, align_corners, g.hashmap_size, g.resolution, pl) * C;
        if (sizeof(T) == 4) {
            float *base = reinterpret_cast<float *>(gtab) + e;
            if (C >= 2) {
#pragma unroll
                for (uint32_t c = 0; c < C; c += 2) red_add_v2_f32(base + c, __fmul_rn(w, gc[c]), __fmul_rn(w, gc[c + 1 < C ? c + 1 : c]));
            } else {
                red_add_f32(base, __fmul_rn(w, gc[0]));
            }
        } else {
            __half *base = reinterpret_cast<__half *>(gtab) + e;     // C even (checked on the host): packed half2 reductions (:298-304)
#pragma unroll
            for (uint32_t c = 0; c + 1 < C; c += 2)
                red_add_h2(base + c, __halves2half2(__float2half_rn(__fmul_rn(w, gc[c])), __float2half_rn(__fmul_rn(w, gc[c + 1]))));
        }
    }
    }
}

// Table backward with the level's gradient table PRIVATISED IN SHARED MEMORY (fp32 tables whose largest level fits, e.g. the 64 KB levels of
// the tri-plane grids).  The direct scatter above is bound by the L2's reduction rate (~50 fp32 reductions per clock for the whole chip, less
// when a training batch concentrates on the few thousand cells of the coarse levels: measured 46 G red/s); here CTA (slice, level) accumulates
// its slice of the batch into a zeroed shared-memory copy of ONE level with shared-memory atomics (hundreds of lane-ops per clock chip-wide),
// then flushes the non-zero quads with `red.global.add.v4.f32`, so the L2 sees (#slices x level size / 4) vector reductions instead of
// 2^D x B scalar ones.  Sums are reassociated (as with any atomic scatter); same products as the direct kernel.
constexpr uint32_t GP_THREADS = 512;

__device__ __forceinline__ void red_add_v4_f32(float *addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <uint32_t D, uint32_t C>
__global__ void __launch_bounds__(GP_THREADS) k_grid_bwd_priv(const float *__restrict__ grad, const float *__restrict__ inputs, const int32_t *__restrict__ offsets,
                                                               float *__restrict__ grad_table, uint32_t B, uint32_t S_slices, float S, uint32_t H,
                                                               uint32_t gridtype, bool align_corners, uint32_t smem_floats) {
    extern __shared__ __align__(16) float s_tab[];
    const uint32_t level = blockIdx.y;
    const LevelGeom g = level_geom(offsets, level, S, H);
    float *gtab = grad_table + (size_t)g.table_off * C;
    const uint32_t n = g.hashmap_size * C;
    const bool priv = n <= smem_floats;                 // a level larger than the launch's shared memory scatters straight to global memory
    float *acc = priv ? s_tab : gtab;
    if (priv) {
        for (uint32_t i = threadIdx.x; i < n; i += GP_THREADS) s_tab[i] = 0.0f;
        __syncthreads();
    }
    const uint32_t per = (B + S_slices - 1) / S_slices;
    const uint32_t b0 = blockIdx.x * per, b1 = min(B, b0 + per);
    constexpr uint32_t U = 4;                        // samples in flight per thread: all their loads are issued before the first atomic
    for (uint32_t bb = b0 + threadIdx.x; bb < b1; bb += U * GP_THREADS) {
        float in[U][D], gc[U][C];
#pragma unroll
        for (uint32_t u = 0; u < U; u++) {
            const uint32_t b = bb + u * GP_THREADS;
            if (b < b1) {
#pragma unroll
                for (uint32_t d = 0; d < D; d++) in[u][d] = __ldcs(inputs + (size_t)b * D + d);
#pragma unroll
                for (uint32_t c = 0; c < C; c++) gc[u][c] = __ldcs(grad + ((size_t)level * B + b) * C + c);
            } else {
#pragma unroll
                for (uint32_t d = 0; d < D; d++) in[u][d] = -1.0f;       // out of range -> skipped below
            }
        }
#pragma unroll
        for (uint32_t u = 0; u < U; u++) {
            bool oob = false;
#pragma unroll
            for (uint32_t d = 0; d < D; d++) oob |= (in[u][d] < 0.0f || in[u][d] > 1.0f);
            if (oob) continue;
            float pos[D];
            uint32_t pg[D];
#pragma unroll
            for (uint32_t d = 0; d < D; d++) {
                pos[d] = __fmaf_rn(in[u][d], g.scale, align_corners ? 0.0f : 0.5f);
                pg[d] = (uint32_t)floorf(pos[d]);
                pos[d] = __fsub_rn(pos[d], (float)pg[d]);
            }
#pragma unroll
            for (uint32_t idx = 0; idx < (1u << D); idx++) {
                float w = 1.0f;
                uint32_t pl[D];
#pragma unroll
                for (uint32_t d = 0; d < D; d++) {
                    if ((idx & (1u << d)) == 0) { w = __fmul_rn(w, __fsub_rn(1.0f, pos[d])); pl[d] = pg[d]; }
                    else                        { w = __fmul_rn(w, pos[d]);                 pl[d] = pg[d] + 1; }
                }
                const uint32_t e = grid_slot<D>(gridtype, align_corners, g.hashmap_size, g.resolution, pl) * C;
#pragma unroll
                for (uint32_t c = 0; c < C; c++) atomicAdd(acc + e + c, __fmul_rn(w, gc[u][c]));
            }
        }
    }
    if (!priv) return;
    __syncthreads();
    if (((g.table_off * C) & 3u) == 0 && (n & 3u) == 0) {        // 16-byte aligned level (offsets are multiples of 8 entries, grid.py:117)
        for (uint32_t i = threadIdx.x; i < n / 4; i += GP_THREADS) {
            const float4 v = reinterpret_cast<const float4 *>(s_tab)[i];
            if (v.x != 0.0f || v.y != 0.0f || v.z != 0.0f || v.w != 0.0f) red_add_v4_f32(gtab + 4 * (size_t)i, v);
        }
    } else {
        for (uint32_t i = threadIdx.x; i < n; i += GP_THREADS) { const float v = s_tab[i]; if (v != 0.0f) red_add_f32(gtab + i, v); }
    }
}

// Tri-plane table backward for the fused training path: the three planes (xy, yz, xz of network.py:208-212) in ONE launch, plane coordinates
// taken straight from xyz (no slicing / normalisation kernels), gradients read from the backward-data kernel's [3][L][M] slabs.  Same
// shared-memory privatisation as k_grid_bwd_priv<2, 1>: CTA (slice, level, plane).
__global__ void __launch_bounds__(GP_THREADS) k_triplane_bwd_priv(const float *__restrict__ grad, const float *__restrict__ xyz, const int32_t *__restrict__ offsets,
                                                                   float *__restrict__ gt_xy, float *__restrict__ gt_yz, float *__restrict__ gt_xz, uint32_t M,
                                                                   uint32_t L, uint32_t S_slices, float S, uint32_t H, float bound, float inv_two_bound,
                                                                   uint32_t smem_floats) {
    extern __shared__ __align__(16) float s_tab[];
    const uint32_t level = blockIdx.y, plane = blockIdx.z;
    const LevelGeom g = level_geom(offsets, level, S, H);
    float *gtab = (plane == 0 ? gt_xy : (plane == 1 ? gt_yz : gt_xz)) + g.table_off;
    const uint32_t ca = plane == 1 ? 1u : 0u, cb = plane == 0 ? 1u : 2u;          // xy = (0,1), yz = (1,2), xz = (0,2)
    const uint32_t n = g.hashmap_size;
    const bool priv = n <= smem_floats;
    float *acc = priv ? s_tab : gtab;
    if (priv) {
        for (uint32_t i = threadIdx.x; i < n; i += GP_THREADS) s_tab[i] = 0.0f;
        __syncthreads();
    }
    const float *gl = grad + ((size_t)plane * L + level) * M;
    const uint32_t per = (M + S_slices - 1) / S_slices;
    const uint32_t b0 = blockIdx.x * per, b1 = min(M, b0 + per);
    constexpr uint32_t U = 4;
    for (uint32_t bb = b0 + threadIdx.x; bb < b1; bb += U * GP_THREADS) {
        float in[U][2], gc[U];
#pragma unroll
        for (uint32_t u = 0; u < U; u++) {
            const uint32_t b = bb + u * GP_THREADS;
            if (b < b1) {
                const float pa = __ldg(xyz + 3 * (size_t)b + ca), pb = __ldg(xyz + 3 * (size_t)b + cb);
                // (x + bound) / (2 bound) like GridEncoder.forward (grid.py:143); exact scaling when 2 bound is a power of two
                in[u][0] = inv_two_bound != 0.0f ? __fmul_rn(__fadd_rn(pa, bound), inv_two_bound) : __fdiv_rn(__fadd_rn(pa, bound), __fmul_rn(2.0f, bound));
                in[u][1] = inv_two_bound != 0.0f ? __fmul_rn(__fadd_rn(pb, bound), inv_two_bound) : __fdiv_rn(__fadd_rn(pb, bound), __fmul_rn(2.0f, bound));
                gc[u] = __ldcs(gl + b);
            } else { in[u][0] = -1.0f; in[u][1] = -1.0f; gc[u] = 0.0f; }
        }
#pragma unroll
        for (uint32_t u = 0; u < U; u++) {
            if (in[u][0] < 0.0f || in[u][0] > 1.0f || in[u][1] < 0.0f || in[u][1] > 1.0f) continue;
            float pos[2];
            uint32_t pg[2];
#pragma unroll
            for (uint32_t d = 0; d < 2; d++) {
                pos[d] = __fmaf_rn(in[u][d], g.scale, 0.5f);
                pg[d] = (uint32_t)floorf(pos[d]);
                pos[d] = __fsub_rn(pos[d], (float)pg[d]);
            }
#pragma unroll
            for (uint32_t idx = 0; idx < 4; idx++) {
                const float wx = (idx & 1u) ? pos[0] : __fsub_rn(1.0f, pos[0]), wy = (idx & 2u) ? pos[1] : __fsub_rn(1.0f, pos[1]);
                const uint32_t pl[2] = {pg[0] + (idx & 1u), pg[1] + ((idx >> 1) & 1u)};
                const uint32_t e = grid_slot<2>(0, false, g.hashmap_size, g.resolution, pl);
                atomicAdd(acc + e, __fmul_rn(__fmul_rn(wx, wy), gc[u]));
            }
        }
    }
    if (!priv) return;
    __syncthreads();
    if ((g.table_off & 3u) == 0 && (n & 3u) == 0 && ((uintptr_t)gtab & 15u) == 0) {
        for (uint32_t i = threadIdx.x; i < n / 4; i += GP_THREADS) {
            const float4 v = reinterpret_cast<const float4 *>(s_tab)[i];
            if (v.x != 0.0f || v.y != 0.0f || v.z != 0.0f || v.w != 0.0f) red_add_v4_f32(gtab + 4 * (size_t)i, v);
        }
    } else {
        for (uint32_t i = threadIdx.x; i < n; i += GP_THREADS) { const float v = s_tab[i]; if (v != 0.0f) red_add_f32(gtab + i, v); }
    }
}

// Same job with FIXED-POINT privatisation.  Shared memory has a native atomic add only for 32-bit integers (ATOMS.ADD); a float atomicAdd there is a
// compare-and-swap loop (LDS + FADD + ATOMS.CAST.SPIN, retried under contention — short_scoreboard 4.6 stalls per issue in the float kernel above).
// So every contribution t = w * g is converted to a fixed-point integer V = rint(t * 2^(K - e)), 2^e > the largest |g| of the CTA's slice (found by
// a first pass over the slice), and added as two fire-and-forget integer atomics, V = hi * 2^lo_bits + lo with |lo| <= 2^(lo_bits - 1) (one atomic when hi == 0).  With at
// most n_max = 4 * slice contributions per slot, lo_bits = 32 - ceil(log2 n_max) and K = min(62 - 2 ceil(log2 n_max), 30) keep both sums inside 32 bits
// (K = 26 for the 65 536-ray step: every contribution is kept to 2^-26 of the slice's largest gradient, the sums are exact — and so independent of
// the order of the atomics — where fp32 accumulation rounds each add to 2^-24 of the running sum).  A non-finite gradient poisons the level (NaN),
// so the GradScaler's overflow check still sees it.
constexpr uint32_t GF_THREADS = 1024;
__global__ void __launch_bounds__(GF_THREADS, 1) k_triplane_bwd_fix(const float *__restrict__ grad, const float *__restrict__ xyz, const int32_t *__restrict__ offsets,
                                                                     float *__restrict__ gt_xy, float *__restrict__ gt_yz, float *__restrict__ gt_xz, uint32_t M,
                                                                     uint32_t L, uint32_t S_slices, float S, uint32_t H, float bound, float inv_two_bound,
                                                                     uint32_t smem_entries, int K, uint32_t lo_bits) {
    extern __shared__ __align__(16) int2 s_acc[];          // .x = hi sum, .y = lo sum (both signed)
    __shared__ float s_red[GF_THREADS / 32];
    __shared__ int s_bad;
    const uint32_t level = blockIdx.y, plane = blockIdx.z;
    const LevelGeom g = level_geom(offsets, level, S, H);
    float *gtab = (plane == 0 ? gt_xy : (plane == 1 ? gt_yz : gt_xz)) + g.table_off;
    const uint32_t ca = plane == 1 ? 1u : 0u, cb = plane == 0 ? 1u : 2u;          // xy = (0,1), yz = (1,2), xz = (0,2)
    const uint32_t n = g.hashmap_size;
    // slot of corner (i, j) = grid_slot<2>() with the level's kind resolved once (it is uniform over the CTA): dense levels index i + j * (res + 1)
    // (always < n: no modulo), hashed levels (i ^ j * 2654435761) mod n — a mask when n is a power of two (grid.py:117 rounds hashed levels to 2^k);
    // the generic `% n` costs ~25 instructions per corner
    const uint32_t row_stride = g.resolution + 1u;
    const bool dense = row_stride <= n && (uint64_t)row_stride * row_stride <= n, pow2 = (n & (n - 1u)) == 0u;
    const bool generic = !dense && !pow2;
    const uint32_t jmul_h = dense ? 0u : 2654435761u, mask_h = dense ? 0xffffffffu : n - 1u, jmul_d = dense ? row_stride : 0u;
    auto slot_of = [&](uint32_t i, uint32_t j) -> uint32_t {
        if (generic) return (i ^ (j * 2654435761u)) % n;
        return ((i ^ (j * jmul_h)) & mask_h) + j * jmul_d;              // branch-free for the two kinds the reference geometry has
    };
    const float *gl = grad + ((size_t)plane * L + level) * M;
    const uint32_t per = (M + S_slices - 1) / S_slices;
    const uint32_t b0 = blockIdx.x * per, b1 = min(M, b0 + per);
    if (n > smem_entries) {                                 // level too large to privatise: scatter straight to global memory
        for (uint32_t b = b0 + threadIdx.x; b < b1; b += GF_THREADS) {
            const float pa = __ldg(xyz + 3 * (size_t)b + ca), pb = __ldg(xyz + 3 * (size_t)b + cb);
            const float u0 = inv_two_bound != 0.0f ? __fmul_rn(__fadd_rn(pa, bound), inv_two_bound) : __fdiv_rn(__fadd_rn(pa, bound), __fmul_rn(2.0f, bound));
            const float u1 = inv_two_bound != 0.0f ? __fmul_rn(__fadd_rn(pb, bound), inv_two_bound) : __fdiv_rn(__fadd_rn(pb, bound), __fmul_rn(2.0f, bound));
            if (u0 < 0.0f || u0 > 1.0f || u1 < 0.0f || u1 > 1.0f) continue;
            const float gc = __ldcs(gl + b);
            float p0 = __fmaf_rn(u0, g.scale, 0.5f), p1 = __fmaf_rn(u1, g.scale, 0.5f);
            const uint32_t i0 = (uint32_t)floorf(p0), i1 = (uint32_t)floorf(p1);
            p0 = __fsub_rn(p0, (float)i0); p1 = __fsub_rn(p1, (float)i1);
#pragma unroll
            for (uint32_t idx = 0; idx < 4; idx++) {
                const float wx = (idx & 1u) ? p0 : __fsub_rn(1.0f, p0), wy = (idx & 2u) ? p1 : __fsub_rn(1.0f, p1);
                red_add_f32(gtab + slot_of(i0 + (idx & 1u), i1 + ((idx >> 1) & 1u)), __fmul_rn(__fmul_rn(wx, wy), gc));
            }
        }
        return;
    }
    for (uint32_t i = threadIdx.x; i < n; i += GF_THREADS) s_acc[i] = make_int2(0, 0);
    if (threadIdx.x == 0) s_bad = 0;
    // ---- pass 1: the largest |g| of the slice (ordinary loads: the second pass finds the values in L1 / L2)
    float gmax = 0.0f;
    bool bad = false;
    for (uint32_t b = b0 + threadIdx.x; b < b1; b += GF_THREADS) {
        const float v = fabsf(__ldg(gl + b));
        bad |= !(v <= 3.0e38f);                             // inf or nan
        gmax = fmaxf(gmax, v);
    }
#pragma unroll
    for (uint32_t o = 16; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    if ((threadIdx.x & 31u) == 0) s_red[threadIdx.x >> 5] = gmax;
    __syncthreads();
    if (bad) s_bad = 1;
    gmax = s_red[0];
#pragma unroll
    for (uint32_t w = 1; w < GF_THREADS / 32; w++) gmax = fmaxf(gmax, s_red[w]);
    __syncthreads();
    if (s_bad) {                                           // uniform
        if (threadIdx.x == 0) red_add_f32(gtab, __int_as_float(0x7fc00000));
        return;
    }
    if (gmax == 0.0f) return;                               // uniform: nothing to add
    int e;
    (void)frexpf(gmax, &e);                                 // gmax < 2^e
    e = max(e, -100);
    const float to_fix = ldexpf(1.0f, K - e);
    const int lo_half = 1 << (lo_bits - 1u);
    // ---- pass 2: fixed-point scatter
    constexpr uint32_t U = 4;
    for (uint32_t bb = b0 + threadIdx.x; bb < b1; bb += U * GF_THREADS) {
        float in[U][2], gc[U];
#pragma unroll
        for (uint32_t u = 0; u < U; u++) {
            const uint32_t b = bb + u * GF_THREADS;
            if (b < b1) {
                const float pa = __ldg(xyz + 3 * (size_t)b + ca), pb = __ldg(xyz + 3 * (size_t)b + cb);
                in[u][0] = inv_two_bound != 0.0f ? __fmul_rn(__fadd_rn(pa, bound), inv_two_bound) : __fdiv_rn(__fadd_rn(pa, bound), __fmul_rn(2.0f, bound));
                in[u][1] = inv_two_bound != 0.0f ? __fmul_rn(__fadd_rn(pb, bound), inv_two_bound) : __fdiv_rn(__fadd_rn(pb, bound), __fmul_rn(2.0f, bound));
                gc[u] = __ldg(gl + b);
            } else { in[u][0] = -1.0f; in[u][1] = -1.0f; gc[u] = 0.0f; }
        }
#pragma unroll
        for (uint32_t u = 0; u < U; u++) {
            if (in[u][0] < 0.0f || in[u][0] > 1.0f || in[u][1] < 0.0f || in[u][1] > 1.0f || gc[u] == 0.0f) continue;
            float pos[2];
            uint32_t pg[2];
#pragma unroll
            for (uint32_t d = 0; d < 2; d++) {
                pos[d] = __fmaf_rn(in[u][d], g.scale, 0.5f);
                pg[d] = (uint32_t)floorf(pos[d]);
                pos[d] = __fsub_rn(pos[d], (float)pg[d]);
            }
#pragma unroll
            for (uint32_t idx = 0; idx < 4; idx++) {
                const float wx = (idx & 1u) ? pos[0] : __fsub_rn(1.0f, pos[0]), wy = (idx & 2u) ? pos[1] : __fsub_rn(1.0f, pos[1]);
                const uint32_t slot = slot_of(pg[0] + (idx & 1u), pg[1] + ((idx >> 1) & 1u));
                const int V = __float2int_rn(__fmul_rn(__fmul_rn(__fmul_rn(wx, wy), gc[u]), to_fix));
                if (V != 0) {
                    // V = hi * 2^lo_bits + lo with a SIGNED lo in [-2^(lo_bits-1), 2^(lo_bits-1)): small contributions of either sign have hi == 0
                    // and cost one atomic
                    const int hi = (V + lo_half) >> lo_bits, lo = V - (hi << lo_bits);
                    if (hi != 0) atomicAdd(&s_acc[slot].x, hi);
                    atomicAdd(&s_acc[slot].y, lo);
                }
            }
        }
    }
    __syncthreads();
    const float from_fix = ldexpf(1.0f, e - K);
    auto value = [&](int2 a) { return __fmul_rn(__ll2float_rn(((long long)a.x << lo_bits) + (long long)a.y), from_fix); };
    if ((g.table_off & 3u) == 0 && (n & 3u) == 0 && ((uintptr_t)gtab & 15u) == 0) {
        for (uint32_t i = threadIdx.x; i < n / 4; i += GF_THREADS) {
            const int4 q0 = reinterpret_cast<const int4 *>(s_acc)[2 * i], q1 = reinterpret_cast<const int4 *>(s_acc)[2 * i + 1];
            const float4 v = make_float4(value(make_int2(q0.x, q0.y)), value(make_int2(q0.z, q0.w)), value(make_int2(q1.x, q1.y)), value(make_int2(q1.z, q1.w)));
            if (v.x != 0.0f || v.y != 0.0f || v.z != 0.0f || v.w != 0.0f) red_add_v4_f32(gtab + 4 * (size_t)i, v);
        }
    } else {
        for (uint32_t i = threadIdx.x; i < n; i += GF_THREADS) { const float v = value(s_acc[i]); if (v != 0.0f) red_add_f32(gtab + i, v); }
    }
}

// largest level (entries) of the grid described by a device `offsets` array — read back ONCE per (pointer, L) and cached; 0 = unknown (the
// stream is being captured and the geometry has not been seen yet).  The value only sizes the privatised kernel's shared memory: the kernel
// re-checks every level against it, so a stale entry costs speed, never correctness.
static uint32_t largest_level_entries(const int32_t *offsets, uint32_t L, cudaStream_t st) {
    struct Entry { const int32_t *p; uint32_t L, dev, max_entries; };
    static Entry cache[16];
    static uint32_t n_cache = 0;
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    for (uint32_t i = 0; i < n_cache; i++) if (cache[i].p == offsets && cache[i].L == L && cache[i].dev == (uint32_t)dev) return cache[i].max_entries;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone || L > 64) { (void)cudaGetLastError(); return 0; }
    int32_t h[65];
    if (cudaMemcpyAsync(h, offsets, sizeof(int32_t) * (L + 1), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    uint32_t mx = 0;
    for (uint32_t l = 0; l < L; l++) { const uint32_t sz = (uint32_t)(h[l + 1] - h[l]); if (h[l + 1] > h[l] && sz > mx) mx = sz; }
    Entry &e = cache[n_cache < 16 ? n_cache++ : 15];
    e.p = offsets; e.L = L; e.dev = (uint32_t)dev; e.max_entries = mx;
    return mx;
}

template <uint32_t D, uint32_t C>
static int launch_bwd_priv(const float *grad, const float *inputs, const int32_t *offsets, float *gtab, uint32_t B, uint32_t L, float S, uint32_t H,
                           uint32_t gridtype, bool ac, uint32_t smem_floats, cudaStream_t st) {
    const size_t smem = sizeof(float) * (size_t)smem_floats;
    auto kern = k_grid_bwd_priv<D, C>;
    B2N_SMEM(kern, smem);
    uint32_t slices = 3u * (uint32_t)sm_count() / L;          // three 512-thread CTAs per SM (64 KB of shared memory each for the tri-plane levels)
    if (const char *e = getenv("B2N_GRID_BWD_SLICES")) slices = (uint32_t)atoi(e);
    const uint32_t cap = ceil_div<uint32_t>(B, 2048);           // at least two rounds of the CTA per slice
    if (slices > cap) slices = cap;
    if (slices < 1) slices = 1;
    kern<<<dim3(slices, L, 1), GP_THREADS, smem, st>>>(grad, inputs, offsets, gtab, B, slices, S, H, gridtype, ac, smem_floats);
    return check_launch("grid_encode_backward");
}

// input gradient: grad_inputs[b,d] = sum_l sum_c grad[l,b,c] * dy_dx[b,l,d,c]   (:317-342)
template <typename T>
__global__ void __launch_bounds__(256) k_grid_input_bwd(const T *__restrict__ grad, const T *__restrict__ dy_dx, T *__restrict__ grad_inputs,
                                                         uint32_t B, uint32_t D, uint32_t C, uint32_t L) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const T *dd = dy_dx + (size_t)b * L * D * C;
    float r = 0.0f;
    for (uint32_t l = 0; l < L; l++)
        for (uint32_t c = 0; c < C; c++)
            acc_corner<T>(r, to_f<T>(grad[((size_t)l * B + b) * C + c]), to_f<T>(dd[(size_t)l * D * C + d * C + c]));
    grad_inputs[t] = from_f<T>(r);
}

// per-level scales exactly as the kernels compute them (ex2.approx + fma) — lets a CPU checker use the GPU's bits
__global__ void k_level_scales(float S, uint32_t H, uint32_t L, float *__restrict__ out) {   // also used by fused_head.cu
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < L) out[l] = __fmaf_rn(exp2f(__fmul_rn((float)l, S)), (float)H, -1.0f);
}

template <typename T, uint32_t D>
static int fwd_dispatch_c(const float *inputs, const T *table, const int32_t *offsets, T *outputs, uint32_t B, uint32_t C, uint32_t L, float S, uint32_t H,
                          T *dy_dx, uint32_t gridtype, bool ac, cudaStream_t st) {
    const dim3 grid(ceil_div<uint32_t>(B, 256), L, 1);
    switch (C) {
        case 1: k_grid_fwd<T, D, 1><<<grid, 256, 0, st>>>(inputs, table, offsets, outputs, B, L, S, H, dy_dx, gridtype, ac); break;
        case 2: k_grid_fwd<T, D, 2><<<grid, 256, 0, st>>>(inputs, table, offsets, outputs, B, L, S, H, dy_dx, gridtype, ac); break;
        case 4: k_grid_fwd<T, D, 4><<<grid, 256, 0, st>>>(inputs, table, offsets, outputs, B, L, S, H, dy_dx, gridtype, ac); break;
        case 8: k_grid_fwd<T, D, 8><<<grid, 256, 0, st>>>(inputs, table, offsets, outputs, B, L, S, H, dy_dx, gridtype, ac); break;
        default: set_error("GridEncoding: C must be 1, 2, 4, or 8."); return 2;
    }
    return check_launch("grid_encode_forward");
}
template <typename T>
static int fwd_dispatch(const float *inputs, const T *table, const int32_t *offsets, T *outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                        uint32_t H, T *dy_dx, uint32_t gridtype, bool ac, cudaStream_t st) {
    switch (D) {
        case 1: return fwd_dispatch_c<T, 1>(inputs, table, offsets, outputs, B, C, L, S, H, dy_dx, gridtype, ac, st);
        case 2: return fwd_dispatch_c<T, 2>(inputs, table, offsets, outputs, B, C, L, S, H, dy_dx, gridtype, ac, st);
        case 3: return fwd_dispatch_c<T, 3>(inputs, table, offsets, outputs, B, C, L, S, H, dy_dx, gridtype, ac, st);
        case 4: return fwd_dispatch_c<T, 4>(inputs, table, offsets, outputs, B, C, L, S, H, dy_dx, gridtype, ac, st);
        case 5: return fwd_dispatch_c<T, 5>(inputs, table, offsets, outputs, B, C, L, S, H, dy_dx, gridtype, ac, st);
        default: set_error("GridEncoding: D must be 1, 2, 3, 4, or 5"); return 2;
    }
}
template <typename T>
static int fwd_rows_dispatch(const float *inputs, const T *table, const int32_t *offsets, T *outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                             uint32_t gridtype, bool ac, cudaStream_t st) {
    const size_t smem = sizeof(LevelGeom) * ((L + 3) & ~3u) + sizeof(T) * GR_THREADS * ((size_t)(L * C) | 1u);
    if (smem > 48 * 1024) { set_error("grid_encode_forward_rows: L * C = %u too wide for the staged kernel", L * C); return 2; }
    const uint32_t grid = ceil_div<uint32_t>(B, GR_THREADS);
#define B2N_ROWS(DD, CC) if (D == DD && C == CC) { k_grid_fwd_rows<T, DD, CC><<<grid, GR_THREADS, smem, st>>>(inputs, table, offsets, outputs, B, L, S, H, gridtype, ac); return check_launch("grid_encode_forward_rows"); }
    B2N_ROWS(2, 1) B2N_ROWS(2, 2) B2N_ROWS(2, 4) B2N_ROWS(2, 8) B2N_ROWS(3, 1) B2N_ROWS(3, 2) B2N_ROWS(3, 4) B2N_ROWS(3, 8) B2N_ROWS(1, 1) B2N_ROWS(1, 2) B2N_ROWS(1, 4) B2N_ROWS(1, 8)
#undef B2N_ROWS
    set_error("grid_encode_forward_rows: D = %u, C = %u not instantiated (D 1..3, C 1, 2, 4, 8)", D, C);
    return 2;
}

template <typename T, uint32_t D>
static int bwd_dispatch_c(const T *grad, const float *inputs, const int32_t *offsets, T *gtab, uint32_t B, uint32_t C, uint32_t L, float S, uint32_t H,
                          uint32_t gridtype, bool ac, cudaStream_t st) {
    const dim3 grid(ceil_div<uint32_t>(B, 256), L, 1);
    switch (C) {
        case 1: k_grid_bwd<T, D, 1><<<grid, 256, 0, st>>>(grad, inputs, offsets, gtab, B, L, S, H, gridtype, ac); break;
        case 2: k_grid_bwd<T, D, 2><<<grid, 256, 0, st>>>(grad, inputs, offsets, gtab, B, L, S, H, gridtype, ac); break;
        case 4: k_grid_bwd<T, D, 4><<<grid, 256, 0, st>>>(grad, inputs, offsets, gtab, B, L, S, H, gridtype, ac); break;
        case 8: k_grid_bwd<T, D, 8><<<grid, 256, 0, st>>>(grad, inputs, offsets, gtab, B, L, S, H, gridtype, ac); break;
        default: set_error("GridEncoding: C must be 1, 2, 4, or 8."); return 2;
    }
    return check_launch("grid_encode_backward");
}
// fp32 tables, batches large enough to amortise zero + flush, largest level <= 160 KB: privatised kernel
static bool bwd_privatised(const float *grad, const float *inputs, const int32_t *offsets, float *gtab, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                           uint32_t H, uint32_t gridtype, bool ac, cudaStream_t st, int &rc) {
    if (B < 16384 || L > 64 || getenv("B2N_GRID_BWD_DIRECT")) return false;
    const uint32_t mx = largest_level_entries(offsets, L, st);
    if (mx == 0 || (size_t)mx * C * sizeof(float) > 160 * 1024) return false;
    const uint32_t fl = mx * C;
#define B2N_PRIV(DD, CC) if (D == DD && C == CC) { rc = launch_bwd_priv<DD, CC>(grad, inputs, offsets, gtab, B, L, S, H, gridtype, ac, fl, st); return true; }
    B2N_PRIV(2, 1) B2N_PRIV(2, 2) B2N_PRIV(2, 4) B2N_PRIV(3, 1) B2N_PRIV(3, 2) B2N_PRIV(3, 4)
#undef B2N_PRIV
    return false;
}

template <typename T>
static int bwd_dispatch(const T *grad, const float *inputs, const int32_t *offsets, T *gtab, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                        uint32_t H, uint32_t gridtype, bool ac, cudaStream_t st) {
    if constexpr (sizeof(T) == 4) {
        int rc = 0;
        if (bwd_privatised((const float *)grad, inputs, offsets, (float *)gtab, B, D, C, L, S, H, gridtype, ac, st, rc)) return rc;
    }
    switch (D) {
        case 1: return bwd_dispatch_c<T, 1>(grad, inputs, offsets, gtab, B, C, L, S, H, gridtype, ac, st);
        case 2: return bwd_dispatch_c<T, 2>(grad, inputs, offsets, gtab, B, C, L, S, H, gridtype, ac, st);
        case 3: return bwd_dispatch_c<T, 3>(grad, inputs, offsets, gtab, B, C, L, S, H, gridtype, ac, st);
        case 4: return bwd_dispatch_c<T, 4>(grad, inputs, offsets, gtab, B, C, L, S, H, gridtype, ac, st);
        case 5: return bwd_dispatch_c<T, 5>(grad, inputs, offsets, gtab, B, C, L, S, H, gridtype, ac, st);
        default: set_error("GridEncoding: D must be 1, 2, 3, 4, or 5"); return 2;
    }
}

}  // namespace b2n

using namespace b2n;

extern "C" {

int b2n_grid_encode_forward(const float *inputs, const void *embeddings, const int32_t *offsets, void *outputs, uint32_t B, uint32_t D, uint32_t C,
                            uint32_t L, float S, uint32_t H, void *dy_dx, uint32_t gridtype, int align_corners, b2n_dtype dtype, void *stream) {
    B2N_REQUIRE(inputs && embeddings && offsets && outputs, "grid_encode_forward: null pointer");
    B2N_REQUIRE(gridtype <= 1, "grid_encode_forward: gridtype must be 0 (hash) or 1 (tiled)");
    B2N_REQUIRE(L >= 1 && L <= 65535, "grid_encode_forward: L=%u out of range", L);
    if (B == 0) return 0;
    if (dtype == B2N_F32)
        return fwd_dispatch<float>(inputs, (const float *)embeddings, offsets, (float *)outputs, B, D, C, L, S, H, (float *)dy_dx, gridtype, align_corners != 0, as_stream(stream));
    if (dtype == B2N_F16)
        return fwd_dispatch<__half>(inputs, (const __half *)embeddings, offsets, (__half *)outputs, B, D, C, L, S, H, (__half *)dy_dx, gridtype, align_corners != 0, as_stream(stream));
    set_error("grid_encode_forward: embeddings must be float32 or float16");
    return 2;
}

// GridEncoder.forward's own layout: outputs [B, L*C] row-major, no input derivative (dy_dx) — same values as b2n_grid_encode_forward followed by the reference's
// permute(1, 0, 2).reshape(B, L*C) (grid.py:52), bit for bit.  D 1..3, C in {1, 2, 4, 8}, L * C small enough for one staged tile (<= ~90 fp32 features).
int b2n_grid_encode_forward_rows(const float *inputs, const void *embeddings, const int32_t *offsets, void *outputs, uint32_t B, uint32_t D, uint32_t C,
                                 uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners, b2n_dtype dtype, void *stream) {
    B2N_REQUIRE(inputs && embeddings && offsets && outputs, "grid_encode_forward_rows: null pointer");
    B2N_REQUIRE(gridtype <= 1, "grid_encode_forward_rows: gridtype must be 0 (hash) or 1 (tiled)");
    B2N_REQUIRE(L >= 1 && L <= 1024, "grid_encode_forward_rows: L=%u out of range", L);
    if (B == 0) return 0;
    if (dtype == B2N_F32)
        return fwd_rows_dispatch<float>(inputs, (const float *)embeddings, offsets, (float *)outputs, B, D, C, L, S, H, gridtype, align_corners != 0, as_stream(stream));
    if (dtype == B2N_F16)
        return fwd_rows_dispatch<__half>(inputs, (const __half *)embeddings, offsets, (__half *)outputs, B, D, C, L, S, H, gridtype, align_corners != 0, as_stream(stream));
    set_error("grid_encode_forward_rows: embeddings must be float32 or float16");
    return 2;
}

int b2n_triplane_grid_backward(const float *grad_planes, const float *xyz, const int32_t *offsets, float *grad_xy, float *grad_yz, float *grad_xz, uint32_t M,
                               uint32_t L, float S, uint32_t H, float bound, void *stream) {
    B2N_REQUIRE(grad_planes && xyz && offsets && grad_xy && grad_yz && grad_xz, "triplane_grid_backward: null pointer");
    B2N_REQUIRE(L >= 1 && L <= 64 && bound > 0.0f, "triplane_grid_backward: L=%u / bound=%f out of range", L, bound);
    if (M == 0) return 0;
    cudaStream_t st = as_stream(stream);
    uint32_t mx = largest_level_entries(offsets, L, st);
    int ex = 0;
    const float two_b = 2.0f * bound;
    const float inv = frexpf(two_b, &ex) == 0.5f ? 1.0f / two_b : 0.0f;
    static const bool fixed_point = !(getenv("B2N_GRID_BWD_FIXED") && getenv("B2N_GRID_BWD_FIXED")[0] == '0');
    if (fixed_point && mx != 0 && (size_t)mx * sizeof(int2) <= 200 * 1024) {
        // fixed-point privatisation: 8 bytes per slot, one 1024-thread CTA per SM
        const size_t smem = sizeof(int2) * (size_t)mx;
        B2N_SMEM(k_triplane_bwd_fix, smem);
        uint32_t slices = 2u * (uint32_t)sm_count() / (3u * L);        // two waves of (slice, level, plane) CTAs
        if (const char *e = getenv("B2N_GRID_BWD_SLICES")) slices = (uint32_t)atoi(e);
        const uint32_t cap = ceil_div<uint32_t>(M, 4096);
        if (slices > cap) slices = cap;
        if (slices < 1) slices = 1;
        const uint32_t per = ceil_div<uint32_t>(M, slices);
        uint32_t nbits = 1;
        while (nbits < 31 && (1ull << nbits) < 4ull * per) nbits++;       // at most 4 contributions per sample and slot
        while (nbits > 24) {                                               // huge batches: more slices keep the per-slot sums inside 32 bits
            slices *= 2;
            nbits = 1;
            while (nbits < 31 && (1ull << nbits) < 4ull * ceil_div<uint32_t>(M, slices)) nbits++;
        }
        const int K = min(62 - 2 * (int)nbits, 30);                        // V = rint(t * 2^(K - e)) must also fit an int32
        k_triplane_bwd_fix<<<dim3(slices, L, 3), GF_THREADS, smem, st>>>(grad_planes, xyz, offsets, grad_xy, grad_yz, grad_xz, M, L, slices, S, H, bound, inv, mx, K,
                                                                           32u - nbits);
        return check_launch("triplane_grid_backward");
    }
    uint32_t fl = (mx != 0 && (size_t)mx * sizeof(float) <= 160 * 1024) ? mx : 0;       // 0: every level scatters straight to global memory
    const size_t smem = sizeof(float) * (size_t)fl;
    B2N_SMEM(k_triplane_bwd_priv, smem);
    uint32_t slices = 3u * (uint32_t)sm_count() / (3u * L);
    if (const char *e = getenv("B2N_GRID_BWD_SLICES")) slices = (uint32_t)atoi(e);
    const uint32_t cap = ceil_div<uint32_t>(M, 2048);
    if (slices > cap) slices = cap;
    if (slices < 1) slices = 1;
    k_triplane_bwd_priv<<<dim3(slices, L, 3), GP_THREADS, smem, st>>>(grad_planes, xyz, offsets, grad_xy, grad_yz, grad_xz, M, L, slices, S, H, bound, inv, fl);
    return check_launch("triplane_grid_backward");
}

int b2n_grid_level_scales(float S, uint32_t H, uint32_t L, float *scales_out, void *stream) {
    B2N_REQUIRE(scales_out, "grid_level_scales: null pointer");
    if (L == 0) return 0;
    k_level_scales<<<ceil_div<uint32_t>(L, 64), 64, 0, as_stream(stream)>>>(S, H, L, scales_out);
    return check_launch("grid_level_scales");
}

int b2n_grid_encode_backward(const void *grad, const float *inputs, const void *embeddings, const int32_t *offsets, void *grad_embeddings, uint32_t B,
                             uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, const void *dy_dx, void *grad_inputs, uint32_t gridtype,
                             int align_corners, b2n_dtype dtype, void *stream) {
    (void)embeddings;
    B2N_REQUIRE(grad && inputs && offsets && grad_embeddings, "grid_encode_backward: null pointer");
    B2N_REQUIRE(gridtype <= 1, "grid_encode_backward: gridtype must be 0 (hash) or 1 (tiled)");
    B2N_REQUIRE(L >= 1 && L <= 65535, "grid_encode_backward: L=%u out of range", L);
    B2N_REQUIRE(!(dtype == B2N_F16 && (C & 1)), "grid_encode_backward: float16 tables need an even C (the reference never runs C=1 in half, grid.py:38)");
    if (B == 0) return 0;
    int rc;
    if (dtype == B2N_F32)
        rc = bwd_dispatch<float>((const float *)grad, inputs, offsets, (float *)grad_embeddings, B, D, C, L, S, H, gridtype, align_corners != 0, as_stream(stream));
    else if (dtype == B2N_F16)
        rc = bwd_dispatch<__half>((const __half *)grad, inputs, offsets, (__half *)grad_embeddings, B, D, C, L, S, H, gridtype, align_corners != 0, as_stream(stream));
    else { set_error("grid_encode_backward: grad must be float32 or float16"); return 2; }
    if (rc) return rc;
    if (dy_dx && grad_inputs) {
        if (dtype == B2N_F32)
            k_grid_input_bwd<float><<<ceil_div<uint32_t>(B * D, 256), 256, 0, as_stream(stream)>>>((const float *)grad, (const float *)dy_dx, (float *)grad_inputs, B, D, C, L);
        else
            k_grid_input_bwd<__half><<<ceil_div<uint32_t>(B * D, 256), 256, 0, as_stream(stream)>>>((const __half *)grad, (const __half *)dy_dx, (__half *)grad_inputs, B, D, C, L);
        return check_launch("grid_encode_backward(inputs)");
    }
    return 0;
}

}  // extern "C"
