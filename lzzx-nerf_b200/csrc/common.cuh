// common.cuh — shared device helpers, launch/error plumbing for libb2nerf.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/b2nerf.h"
#include "../../include/b2nerf_fused.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libb2nerf is written for sm_100a (B200) only"
#endif

namespace b2n {

// ---- host-side error plumbing -----------------------------------------------------------------------
void set_error(const char *fmt, ...);           // thread-local message, returned by b2n_last_error()
extern std::atomic<uint64_t> g_launches;        // kernels launched by this library (bench.py gpu_launches)

inline int check_launch(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
#define B2N_REQUIRE(cond, ...) do { if (!(cond)) { b2n::set_error(__VA_ARGS__); return 2; } } while (0)
#define B2N_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
    b2n::set_error("%s: %s", #call, cudaGetErrorString(e__)); return 3; } } while (0)

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
template <typename T> inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

// SM count of the current device (cached); grids for grid-stride kernels are sized in multiples of it
int sm_count();
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE property of a kernel: remembered per (device, kernel) and raised when a launch asks for more,
// so a process that drives several GPUs configures each of them (a process-wide `static bool` would configure only the first).  0 = ok.
int ensure_dynamic_smem_impl(const void *kernel, size_t bytes);
template <typename K> inline int ensure_dynamic_smem(K kernel, size_t bytes) { return ensure_dynamic_smem_impl(reinterpret_cast<const void *>(kernel), bytes); }
#define B2N_SMEM(kernel, bytes) do { if (int rc__ = b2n::ensure_dynamic_smem(kernel, bytes)) return rc__; } while (0)
// library-internal device scratch (grow-only, per device); never exposed to the caller
void *scratch(size_t bytes, int slot);

// ---- device helpers -----------------------------------------------------------------------------------
// Morton / bit-interleave, 10 bits per axis — same integer function as raymarching.cu:56-81
__host__ __device__ __forceinline__ uint32_t spread3(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t morton_enc(uint32_t x, uint32_t y, uint32_t z) {
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}
__host__ __device__ __forceinline__ uint32_t compact3(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

// streaming (read-once) loads/stores: keep L1 for the tables / bitfield
__device__ __forceinline__ float ld_stream(const float *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float *p, float v) { __stcs(p, v); }

}  // namespace b2n
