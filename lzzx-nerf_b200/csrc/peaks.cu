// peaks.cu — MEASUREMENT probes (not on the product path): the machine's rate for the two access patterns the grid encoder is bound by, so that its kernels can be
// reported against a measured ceiling instead of against HBM bandwidth, which they barely touch (grid_encode fwd is bound by the L2 / L1 gather rate,
// grid_encode bwd by the L2 reduction rate; MEASURED_PEAKS.json only has HBM copy bandwidth and dense bf16 throughput).
//   * b2n_probe_gather: every thread issues `loads` independent reads of VEC consecutive floats at pseudo-random VEC-aligned positions of a table of n floats
//     (n a power of two; 163 584 x 4 B = 654 KB for one tri-plane table: L2-resident, larger than L1) — a warp's 32 lanes hit 32 different sectors, like
//     random sample points do;
//   * b2n_probe_red: the same positions with red.global.add.f32.
// profiles/kernel_rooflines.py times them with CUDA events and reports loads / s (reductions / s) as the ceilings.
#include "common.cuh"

namespace b2n {

__device__ __forceinline__ uint32_t lcg(uint32_t &x) { x = x * 1664525u + 1013904223u; return x >> 7; }

template <int VEC>
__global__ void __launch_bounds__(256) k_probe_gather(const float *__restrict__ table, uint32_t mask, uint32_t loads, float *__restrict__ sink) {
    uint32_t s[8];
#pragma unroll
    for (int j = 0; j < 8; j++) s[j] = (blockIdx.x * 256u + threadIdx.x) * 8u + j + 12345u * (j + 1);
    float acc = 0.0f;
    for (uint32_t i = 0; i < loads; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t e = (lcg(s[j]) & mask) & ~(uint32_t)(VEC - 1);
            if (VEC == 1) acc += __ldg(table + e);
            else if (VEC == 2) { const float2 v = __ldg(reinterpret_cast<const float2 *>(table + e)); acc += v.x + v.y; }
            else { const float4 v = __ldg(reinterpret_cast<const float4 *>(table + e)); acc += v.x + v.y + v.z + v.w; }
        }
    }
    if (acc == 123.456f) sink[0] = acc;           // keeps the loads alive
}

__global__ void __launch_bounds__(256) k_probe_red(float *__restrict__ table, uint32_t mask, uint32_t loads) {
    uint32_t s[8];
#pragma unroll
    for (int j = 0; j < 8; j++) s[j] = (blockIdx.x * 256u + threadIdx.x) * 8u + j + 12345u * (j + 1);
    for (uint32_t i = 0; i < loads; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(table + (lcg(s[j]) & mask)), "f"(1.0f) : "memory");
    }
}

}  // namespace b2n

using namespace b2n;

extern "C" int b2n_probe_gather(const float *table, uint32_t n_floats, uint32_t vec, uint32_t loads_per_thread, uint32_t blocks, float *sink, void *stream) {
    B2N_REQUIRE(table && sink, "probe_gather: null pointer");
    B2N_REQUIRE(n_floats >= 8 && (n_floats & (n_floats - 1)) == 0, "probe_gather: the table size must be a power of two");
    B2N_REQUIRE(vec == 1 || vec == 2 || vec == 4, "probe_gather: vec must be 1, 2 or 4");
    B2N_REQUIRE(blocks >= 1 && loads_per_thread >= 8 && loads_per_thread % 8 == 0, "probe_gather: loads_per_thread must be a multiple of 8");
    cudaStream_t st = as_stream(stream);
    if (vec == 1) k_probe_gather<1><<<blocks, 256, 0, st>>>(table, n_floats - 1, loads_per_thread, sink);
    else if (vec == 2) k_probe_gather<2><<<blocks, 256, 0, st>>>(table, n_floats - 1, loads_per_thread, sink);
    else k_probe_gather<4><<<blocks, 256, 0, st>>>(table, n_floats - 1, loads_per_thread, sink);
    return check_launch("probe_gather");
}

extern "C" int b2n_probe_red(float *table, uint32_t n_floats, uint32_t reds_per_thread, uint32_t blocks, void *stream) {
    B2N_REQUIRE(table, "probe_red: null pointer");
    B2N_REQUIRE(n_floats >= 8 && (n_floats & (n_floats - 1)) == 0, "probe_red: the table size must be a power of two");
    B2N_REQUIRE(blocks >= 1 && reds_per_thread >= 8 && reds_per_thread % 8 == 0, "probe_red: reds_per_thread must be a multiple of 8");
    k_probe_red<<<blocks, 256, 0, as_stream(stream)>>>(table, n_floats - 1, reds_per_thread);
    return check_launch("probe_red");
}
