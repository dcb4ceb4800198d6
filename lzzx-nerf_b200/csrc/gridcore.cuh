// gridcore.cuh — the D=2, C=1, fp32 grid lookup shared by the fused kernels (tri-plane head model).
// Same float arithmetic, in the same order, as k_grid_fwd<float,2,1> (gridenc.cu) and hence as the reference's
// kernel_grid<float,2,1> (gridencoder.cu:124-175): pos = fma(u, scale, 0.5); w = (1-px|px)*(1-py|py); four fmas in
// corner order (0,0),(1,0),(0,1),(1,1).  Index math is integer-exact and restructured for speed: per-level constants
// are precomputed once per CTA, dense levels need no modulo, hashed levels use a mask when the size is a power of two.
#pragma once
#include "common.cuh"

namespace b2n {

struct Lvl2 {
    float scale;         // exp2f(level*S)*H - 1   (ex2.approx + fma, like the reference)
    uint32_t stride;     // resolution + 1 (align_corners = false)
    uint32_t size;       // entries in this level
    uint32_t off;        // first entry of this level in the table
    uint32_t flags;      // bit0: hashed, bit1: size is a power of two, bit2: index can exceed size (needs modulo)
};

__device__ __forceinline__ Lvl2 make_lvl2(const int32_t *__restrict__ offsets, uint32_t level, float S, uint32_t H, uint32_t gridtype) {
    Lvl2 g;
    g.off = (uint32_t)offsets[level];
    g.size = (uint32_t)offsets[level + 1] - g.off;
    g.scale = __fmaf_rn(exp2f(__fmul_rn((float)level, S)), (float)H, -1.0f);
    const uint32_t res = (uint32_t)ceilf(g.scale) + 1u;
    g.stride = res + 1u;
    // gridencoder.cu:54-72 for D = 2: stride after the loop is (res+1)^2 when res+1 <= size, else res+1
    const uint64_t full = (g.stride <= g.size) ? (uint64_t)g.stride * g.stride : (uint64_t)g.stride;
    const bool hashed = (gridtype == 0) && (full > g.size);
    const bool pow2 = (g.size & (g.size - 1)) == 0;
    g.flags = (hashed ? 1u : 0u) | (pow2 ? 2u : 0u) | ((full > g.size) ? 4u : 0u);
    return g;
}

// generic modulo, kept out of line: only tiled levels whose dense index range exceeds a non-power-of-two size get here
__device__ __noinline__ uint32_t wrap_slow(uint32_t i, uint32_t size) { return i % size; }
__device__ __forceinline__ uint32_t wrap(uint32_t i, const Lvl2 &g) {
    if (!(g.flags & 4u)) return i;
    return (g.flags & 2u) ? (i & (g.size - 1u)) : wrap_slow(i, g.size);
}

// corner indices for cell (x, y): i00, i10, i01, i11   (branch-free select between the hash and the dense/tiled index)
__device__ __forceinline__ void lvl2_corners(const Lvl2 &g, uint32_t x, uint32_t y, uint32_t &i00, uint32_t &i10, uint32_t &i01, uint32_t &i11) {
    const bool hashed = g.flags & 1u;
    const uint32_t mul = hashed ? 2654435761u : ((g.stride <= g.size) ? g.stride : 0u);   // dense: only the x term if stride > size (gridencoder.cu:60)
    const uint32_t m0 = y * mul, m1 = m0 + mul;
    i00 = wrap(hashed ? (x ^ m0) : (x + m0), g); i10 = wrap(hashed ? ((x + 1u) ^ m0) : (x + 1u + m0), g);
    i01 = wrap(hashed ? (x ^ m1) : (x + m1), g); i11 = wrap(hashed ? ((x + 1u) ^ m1) : (x + 1u + m1), g);
}

struct Cell2 { uint32_t x, y; float fx, fy; };
__device__ __forceinline__ Cell2 lvl2_cell(const Lvl2 &g, float u, float v) {
    Cell2 c;
    const float px = __fmaf_rn(u, g.scale, 0.5f), py = __fmaf_rn(v, g.scale, 0.5f);
    c.x = (uint32_t)floorf(px); c.y = (uint32_t)floorf(py);
    c.fx = __fsub_rn(px, (float)c.x); c.fy = __fsub_rn(py, (float)c.y);
    return c;
}

// pair load: when i1 == i0 ^ 1 the two x-neighbours share one aligned 8-byte word (level offsets are multiples of 8).
// Branch-free: the aligned pair around i0 is always fetched with one 8-byte load; the (rare, lane-divergent) other case adds a
// predicated scalar load instead of a divergent branch (BSSY/BSYNC showed up as 20 % of the stall samples in the first profile).
__device__ __forceinline__ void lvl2_load_pair(const float *__restrict__ tab, uint32_t i0, uint32_t i1, float &v0, float &v1) {
    const float2 pr = __ldg(reinterpret_cast<const float2 *>(tab) + (i0 >> 1));
    const bool paired = (i1 == (i0 ^ 1u));
    float other = 0.0f;
    if (!paired) other = __ldg(tab + i1);
    v0 = (i0 & 1u) ? pr.y : pr.x;
    v1 = paired ? ((i0 & 1u) ? pr.x : pr.y) : other;
}

// interpolated feature of one level (inputs u, v already checked to be inside [0,1])
__device__ __forceinline__ float lvl2_interp(const float *__restrict__ table, const Lvl2 &g, float u, float v) {
    const Cell2 c = lvl2_cell(g, u, v);
    uint32_t i00, i10, i01, i11;
    lvl2_corners(g, c.x, c.y, i00, i10, i01, i11);
    const float *tab = table + g.off;                 // g.off is even for every table built by GridEncoder (grid.py:117)
    float v00, v10, v01, v11;
    if (g.off & 1u) { v00 = __ldg(tab + i00); v10 = __ldg(tab + i10); v01 = __ldg(tab + i01); v11 = __ldg(tab + i11); }
    else { lvl2_load_pair(tab, i00, i10, v00, v10); lvl2_load_pair(tab, i01, i11, v01, v11); }
    const float wx0 = __fsub_rn(1.0f, c.fx), wy0 = __fsub_rn(1.0f, c.fy);
    float r = __fmaf_rn(__fmul_rn(wx0, wy0), v00, 0.0f);
    r = __fmaf_rn(__fmul_rn(c.fx, wy0), v10, r);
    r = __fmaf_rn(__fmul_rn(wx0, c.fy), v01, r);
    r = __fmaf_rn(__fmul_rn(c.fx, c.fy), v11, r);
    return r;
}

}  // namespace b2n
