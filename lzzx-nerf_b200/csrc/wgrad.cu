// wgrad.cu — weight gradient of a bias-free Linear over a tall activation matrix:  dW[out,in] += dY[M,out]^T · X[M,in]
// (fp16 operands, fp32 accumulation), M = 10^5..10^6 samples, out / in <= 128 — the shape of every head-MLP layer in a training step
// (nerf_triplane/network.py:73-94 through autograd's LinearBackward).
//
// The reduction runs over the SAMPLE dimension, which is the slow (row) dimension of both operands in memory, so both are "MN-major"
// tensor-core operands: a 64-sample chunk of dY / X is copied row by row into shared memory in the canonical SWIZZLE_128B MN-major layout
// (128-byte rows of 64 features for one sample, 8-sample groups of 1024 B, 64-feature column blocks 8 KB apart) and fed to tcgen05.mma
// with a_major = b_major = MN — no transposition anywhere.  Two CTAs per SM walk their share of the chunks with two operand buffers each (the
// loads of chunk c+1 overlap the MMAs of chunk c), accumulates the whole out x in product in TMEM (M = 128 lanes, N = in columns) and
// adds it once to the fp32 result with red.global — so the library's one-wave split (4-6 CTAs, 170-300 us per layer measured) becomes an
// HBM-bound stream over the activations.
#include "common.cuh"
#include "tc5.cuh"
#include <cstdlib>

namespace b2n {
using namespace tc5;

constexpr uint32_t WG_THREADS = 256;
constexpr uint32_t WG_CHUNK = 64;                        // samples per chunk = 4 MMA K-steps
constexpr uint32_t WG_BLOCK_BYTES = WG_CHUNK * 128;      // one 64-feature column block of a chunk: 64 sample rows x 128 B
constexpr uint32_t WG_OPERAND_BYTES = 2 * WG_BLOCK_BYTES;   // up to 128 features
constexpr uint32_t WG_SMEM = 4 * WG_OPERAND_BYTES + 1024 + 64;

// MN-major SWIZZLE_128B descriptor: LBO = distance between 64-element column blocks, SBO = distance between 8-row (K) groups
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(WG_BLOCK_BYTES >> 4) << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// byte offset of element (sample k, feature j) inside an operand buffer
__device__ __forceinline__ uint32_t mn_offset(uint32_t k, uint32_t j) {
    return (j >> 6) * WG_BLOCK_BYTES + (k >> 3) * 1024u + (k & 7u) * 128u + ((((j & 63u) >> 3) ^ (k & 7u)) << 4) + (j & 7u) * 2u;
}

// copy rows [row0, row0 + WG_CHUNK) of a row-major fp16 matrix [M, width] into an operand buffer.  V = halves per access (V | width, pointer
// 2V-aligned).  Threads are laid out as (row, access-in-row) with a power-of-two number of threads per row, so there is no division in the
// loop; loads are issued in batches of 4 rows per thread before the first store.
template <uint32_t V> struct VecOf;
template <> struct VecOf<8> { typedef uint4 type; };
template <> struct VecOf<4> { typedef uint2 type; };
template <> struct VecOf<2> { typedef uint32_t type; };
template <> struct VecOf<1> { typedef uint16_t type; };

template <uint32_t V>
__device__ __forceinline__ void load_chunk(uint8_t *buf, const __half *__restrict__ src, uint32_t row0, uint32_t M, uint32_t width) {
    typedef typename VecOf<V>::type vec_t;
    const uint32_t nv = width / V;                                   // accesses per row (<= 128)
    const uint32_t lg = 32u - __clz(nv - 1u | 0u) ;                  // ceil(log2(nv)) for nv >= 2; nv == 1 -> __clz(0) = 32 -> 0
    const uint32_t tpr = 1u << lg;                                    // threads per row
    const uint32_t col = threadIdx.x & (tpr - 1u), r0 = threadIdx.x >> lg, rstep = WG_THREADS >> lg;
    if (col >= nv) return;
    const uint32_t j = col * V;
    for (uint32_t k0 = r0; k0 < WG_CHUNK; k0 += 4 * rstep) {
        vec_t v[4];
#pragma unroll
        for (uint32_t u = 0; u < 4; u++) {
            const uint32_t k = k0 + u * rstep;
            v[u] = vec_t();
            if (k < WG_CHUNK && row0 + k < M) v[u] = __ldcs(reinterpret_cast<const vec_t *>(src + (size_t)(row0 + k) * width + j));
        }
#pragma unroll
        for (uint32_t u = 0; u < 4; u++) {
            const uint32_t k = k0 + u * rstep;
            if (k < WG_CHUNK) *reinterpret_cast<vec_t *>(buf + mn_offset(k, j)) = v[u];
        }
    }
}
// Odd widths (69, 65, 3, 1: rows are not even 4-byte aligned) with a 16-byte aligned matrix: a chunk starts at a multiple of 64 rows, i.e. at a
// multiple of 128 * width bytes, so the chunk is read as a FLAT stream of aligned 16-byte vectors and only the shared-memory side is
// element-wise (8 two-byte stores per vector, row/column advanced incrementally: one division per vector).
__device__ __forceinline__ void load_chunk_flat(uint8_t *buf, const __half *__restrict__ src, uint32_t row0, uint32_t M, uint32_t width) {
    const uint32_t nvec = WG_CHUNK * width / 8u;                     // 64 * width halves = 8 * width vectors
    const size_t e0 = (size_t)row0 * width, e_end = (size_t)M * width;
    for (uint32_t v0 = threadIdx.x; v0 < nvec; v0 += 2 * WG_THREADS) {
        uint4 q[2];
#pragma unroll
        for (uint32_t u = 0; u < 2; u++) {
            const uint32_t vi = v0 + u * WG_THREADS;
            q[u] = make_uint4(0, 0, 0, 0);
            if (vi < nvec) {
                const size_t e = e0 + (size_t)vi * 8u;
                if (e + 8u <= e_end) q[u] = __ldcs(reinterpret_cast<const uint4 *>(src + e));
                else {                                                // last rows of the matrix: element-wise, zero beyond the end
                    __half h[8];
#pragma unroll
                    for (uint32_t t = 0; t < 8; t++) h[t] = e + t < e_end ? src[e + t] : __float2half_rn(0.0f);
                    q[u] = *reinterpret_cast<uint4 *>(h);
                }
            }
        }
#pragma unroll
        for (uint32_t u = 0; u < 2; u++) {
            const uint32_t vi = v0 + u * WG_THREADS;
            if (vi < nvec) {
                uint32_t k = vi * 8u / width, j = vi * 8u - k * width;
                const uint16_t *h = reinterpret_cast<const uint16_t *>(&q[u]);
#pragma unroll
                for (uint32_t t = 0; t < 8; t++) {
                    *reinterpret_cast<uint16_t *>(buf + mn_offset(k, j)) = h[t];
                    if (++j == width) { j = 0; k++; }
                }
            }
        }
    }
}
__device__ __noinline__ void load_chunk_any(uint32_t v, uint8_t *buf, const __half *src, uint32_t row0, uint32_t M, uint32_t width) {
    if (v == 0) load_chunk_flat(buf, src, row0, M, width);
    else if (v == 8) load_chunk<8>(buf, src, row0, M, width);
    else if (v == 4) load_chunk<4>(buf, src, row0, M, width);
    else if (v == 2) load_chunk<2>(buf, src, row0, M, width);
    else load_chunk<1>(buf, src, row0, M, width);
}

__global__ void __launch_bounds__(WG_THREADS, 3) k_linear_wgrad(const __half *__restrict__ dy, const __half *__restrict__ x, uint32_t M, uint32_t out_dim,
                                                                 uint32_t in_dim, uint32_t va, uint32_t vb, float *__restrict__ dw, uint32_t replicas, uint32_t rstride) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + 4 * WG_OPERAND_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t n_chunks = (M + WG_CHUNK - 1) / WG_CHUNK;
    const uint32_t n_pad = (in_dim + 15u) & ~15u;                      // MMA N
    dw += (size_t)(blockIdx.x % replicas) * rstride;                   // spread the final reductions over `replicas` copies of the result
    const bool vec4 = (in_dim & 3u) == 0 && ((uintptr_t)dw & 15u) == 0;

    // padding features (j >= width) are never written by the loads: zero everything once
    for (uint32_t i = tid; i < 4 * WG_OPERAND_BYTES / 16; i += WG_THREADS) reinterpret_cast<uint4 *>(base)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    if (warp == 1) tmem_alloc(tmem_slot, 128);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = idesc_f16(128, n_pad) | (1u << 15) | (1u << 16);       // A and B MN-major

    uint32_t it = 0;
    for (uint32_t c = blockIdx.x; c < n_chunks; c += gridDim.x, it++) {
        const uint32_t b = it & 1u;
        if (it >= 2) mbar_wait(&bars[b], ((it >> 1) - 1u) & 1u);            // the MMAs that read this buffer two chunks ago are done
        uint8_t *sA = base + b * WG_OPERAND_BYTES, *sB = base + (2u + b) * WG_OPERAND_BYTES;      // two buffers per operand
        load_chunk_any(va, sA, dy, c * WG_CHUNK, M, out_dim);
        load_chunk_any(vb, sB, x, c * WG_CHUNK, M, in_dim);
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            fence_after_sync();
            uint64_t da = smem_desc_mn_sw128(smem_u32(sA)), db = smem_desc_mn_sw128(smem_u32(sB));
#pragma unroll 1
            for (uint32_t k = 0; k < WG_CHUNK / 16; k++, da += 2048u >> 4, db += 2048u >> 4) mma_f16_ss(tmem, da, db, idesc, it > 0 || k > 0);
            mma_commit(&bars[b]);
        }
    }
    if (it > 0) {
        const uint32_t last = it - 1;
        mbar_wait(&bars[last & 1u], (last >> 1) & 1u);                      // a commit covers every MMA issued before it
        fence_after_sync();
        if (warp < 4) {
            const uint32_t o = tid;                                         // TMEM lane = output row
            const uint32_t taddr = tmem + ((warp * 32u) << 16);
            for (uint32_t cb = 0; cb < n_pad; cb += 16) {
                uint32_t acc[16];
                ld16(taddr + cb, acc);
                wait_ld();
                if (o < out_dim) {
                    float *row = dw + (size_t)o * in_dim + cb;
                    if (vec4) {                                          // in_dim % 4 == 0 and dw 16-byte aligned: one vector reduction per 4 columns
#pragma unroll
                        for (uint32_t j = 0; j < 16; j += 4)
                            if (cb + j < in_dim)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + j), "f"(__uint_as_float(acc[j])), "f"(__uint_as_float(acc[j + 1])),
                                             "f"(__uint_as_float(acc[j + 2])), "f"(__uint_as_float(acc[j + 3])) : "memory");
                    } else {
#pragma unroll
                        for (uint32_t j = 0; j < 16; j++)
                            if (cb + j < in_dim) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(row + j), "f"(__uint_as_float(acc[j])) : "memory");
                    }
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 128);
}

// ---- pipelined variant for 16-byte aligned operands (every operand of the fused training path: row pitches are multiples of 8 halves) ------
// One CTA per SM, a ring of WP_STAGES operand stages filled with cp.async (16-byte asynchronous copies, zero-fill past the last row): the
// copies of chunks c+1 .. c+3 are in flight while chunk c is multiplied, so the per-chunk global-load latency that bounds the simple kernel
// (one chunk per CTA in flight, ~2 us per chunk) is hidden.
#ifndef WP_STAGES_N
#define WP_STAGES_N 4
#endif
constexpr uint32_t WP_STAGES = WP_STAGES_N;
#ifndef WP_AHEAD_N
#define WP_AHEAD_N (WP_STAGES_N - 1)
#endif
constexpr uint32_t WP_AHEAD = WP_AHEAD_N;                              // chunks in flight ahead of the one being multiplied
constexpr uint32_t WP_STAGE_BYTES = 2 * WG_OPERAND_BYTES;                 // A (<= 128 features) + B (<= 128 features) of one 64-sample chunk
constexpr uint32_t WP_SMEM = WP_STAGES * WP_STAGE_BYTES + 1024 + 128;

__device__ __forceinline__ void cp_async16(uint32_t dst_saddr, const void *src, bool valid) {
    const uint32_t sz = valid ? 16u : 0u;                                 // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_saddr), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct RowMap { uint32_t col, r0, rstep, nv; };                           // thread -> (16-byte column, first row, row step) of an operand
__device__ __forceinline__ RowMap row_map(uint32_t width) {
    RowMap r;
    r.nv = width / 8u;
    const uint32_t lg = 32u - __clz(r.nv - 1u);
    r.col = threadIdx.x & ((1u << lg) - 1u);
    r.r0 = threadIdx.x >> lg;
    r.rstep = WG_THREADS >> lg;
    return r;
}
__device__ __forceinline__ void issue_operand(uint32_t s_base, const __half *__restrict__ src, const RowMap &r, uint32_t row0, uint32_t M, uint32_t width,
                                              uint32_t chunk = WG_CHUNK) {
    if (r.col >= r.nv) return;
    const uint32_t j = r.col * 8u;
#pragma unroll
    for (uint32_t u = 0; u < 4; u++) {                                    // <= 4 rows per thread (width <= 128 with 64-row chunks, <= 64 with 128-row chunks)
        const uint32_t k = r.r0 + u * r.rstep;
        if (k < chunk) {
            const uint32_t row = row0 + k;
            const bool ok = row < M;
            cp_async16(s_base + mn_offset(k, j), src + (size_t)(ok ? row : 0) * width + j, ok);
        }
    }
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Warp-specialised: 8 producer warps keep the cp.async ring full (each thread signals "my part of chunk i has landed" on full[stage]), one extra
// warp issues the MMAs (waits full[stage], commits to empty[stage]); nobody executes a CTA-wide barrier inside the loop.
constexpr uint32_t WP_THREADS = WG_THREADS + 32;

__global__ void __launch_bounds__(WP_THREADS, 1) k_linear_wgrad_pipe(const __half *__restrict__ dy, const __half *__restrict__ x, uint32_t M, uint32_t out_dim,
                                                                      uint32_t in_dim, float *__restrict__ dw, uint32_t replicas, uint32_t rstride) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full = reinterpret_cast<uint64_t *>(base + WP_STAGES * WP_STAGE_BYTES), *empty = full + WP_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(empty + WP_STAGES);
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const bool producer = tid < WG_THREADS;
    const uint32_t n_chunks = (M + WG_CHUNK - 1) / WG_CHUNK;
    const uint32_t n_pad = (in_dim + 15u) & ~15u;
    dw += (size_t)(blockIdx.x % replicas) * rstride;
    const bool vec4 = (in_dim & 3u) == 0 && ((uintptr_t)dw & 15u) == 0;

    for (uint32_t i = tid; i < WP_STAGES * WP_STAGE_BYTES / 16; i += WP_THREADS) reinterpret_cast<uint4 *>(base)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (uint32_t s = 0; s < WP_STAGES; s++) { mbar_init(&full[s], WG_THREADS); mbar_init(&empty[s], 1); }
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 128);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t base_a = smem_u32(base);
    const uint32_t my_n = blockIdx.x < n_chunks ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    constexpr uint32_t AHEAD = WP_AHEAD;                             // chunks in flight per producer
    if (producer) {
        const RowMap ra = row_map(out_dim), rb = row_map(in_dim);
        auto issue = [&](uint32_t i) {                                    // chunk i of this CTA -> stage i % WP_STAGES
            const uint32_t st = base_a + (i % WP_STAGES) * WP_STAGE_BYTES, row0 = (blockIdx.x + i * gridDim.x) * WG_CHUNK;
            issue_operand(st, dy, ra, row0, M, out_dim);
            issue_operand(st + WG_OPERAND_BYTES, x, rb, row0, M, in_dim);
        };
        for (uint32_t i = 0; i < AHEAD; i++) { if (i < my_n) issue(i); cp_async_commit(); }
        for (uint32_t i = 0; i < my_n; i++) {
            cp_async_wait<AHEAD - 1>();                                   // this thread's copies of chunk i have landed
            fence_proxy_async();                                          // ... and are visible to the tensor pipe (async proxy)
            mbar_arrive(&full[i % WP_STAGES]);
            const uint32_t nxt = i + AHEAD;                               // refills the stage chunk nxt - WP_STAGES (= i - 1) was multiplied from
            if (nxt < my_n) {
                if (nxt >= WP_STAGES) mbar_wait(&empty[nxt % WP_STAGES], ((nxt / WP_STAGES) - 1u) & 1u);
                issue(nxt);
            }
            cp_async_commit();
        }
    } else if (tid == WG_THREADS) {                                       // the MMA thread
        const uint32_t idesc = idesc_f16(128, n_pad) | (1u << 15) | (1u << 16);
        for (uint32_t i = 0; i < my_n; i++) {
            const uint32_t s = i % WP_STAGES;
            mbar_wait(&full[s], (i / WP_STAGES) & 1u);
            fence_after_sync();
            uint64_t da = smem_desc_mn_sw128(base_a + s * WP_STAGE_BYTES), db = smem_desc_mn_sw128(base_a + s * WP_STAGE_BYTES + WG_OPERAND_BYTES);
#pragma unroll 1
            for (uint32_t k = 0; k < WG_CHUNK / 16; k++, da += 2048u >> 4, db += 2048u >> 4) mma_f16_ss(tmem, da, db, idesc, i > 0 || k > 0);
            mma_commit(&empty[s]);
        }
    }
    if (my_n > 0 && warp < 4) {                                           // epilogue warps (TMEM lanes 32 * warp ..): the last commit covers every MMA before it
        const uint32_t last = my_n - 1;
        mbar_wait(&empty[last % WP_STAGES], (last / WP_STAGES) & 1u);
        fence_after_sync();
        {
            const uint32_t o = tid;
            const uint32_t taddr = tmem + ((warp * 32u) << 16);
            for (uint32_t cb = 0; cb < n_pad; cb += 16) {
                uint32_t acc[16];
                ld16(taddr + cb, acc);
                wait_ld();
                if (o < out_dim) {
                    float *row = dw + (size_t)o * in_dim + cb;
                    if (vec4) {
#pragma unroll
                        for (uint32_t j = 0; j < 16; j += 4)
                            if (cb + j < in_dim)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + j), "f"(__uint_as_float(acc[j])), "f"(__uint_as_float(acc[j + 1])),
                                             "f"(__uint_as_float(acc[j + 2])), "f"(__uint_as_float(acc[j + 3])) : "memory");
                    } else {
#pragma unroll
                        for (uint32_t j = 0; j < 16; j++)
                            if (cb + j < in_dim) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(row + j), "f"(__uint_as_float(acc[j])) : "memory");
                    }
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 128);
}

// ---- all weight gradients of one backward in ONE launch ---------------------------------------------------------------------------------
// The 11-13 products of a step share M and differ only in their (narrow) operands; launched one by one each pays ~5 us of launch / ramp / tail on
// 10-25 us of streaming.  Here every CTA walks the jobs back to back through the same cp.async ring: the producers never drain between jobs, the MMA
// warp alternates between two TMEM accumulators, and four epilogue warps reduce job j into memory while job j + 1 is being multiplied.
struct WgradJobs { b2n_wgrad_job j[16]; uint32_t n, M, replicas, rstride; };
constexpr uint32_t WM_THREADS = WG_THREADS + 32 + 128;                    // 8 producer warps, the MMA warp (8), epilogue warps 9..12
constexpr uint32_t WM_SMEM = WP_STAGES * WP_STAGE_BYTES + 1024 + 256;

__global__ void __launch_bounds__(WM_THREADS, 1) k_linear_wgrad_multi(const __grid_constant__ WgradJobs J) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full = reinterpret_cast<uint64_t *>(base + WP_STAGES * WP_STAGE_BYTES), *empty = full + WP_STAGES, *acc_full = empty + WP_STAGES, *acc_free = acc_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_free + 2);
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t n_chunks = (J.M + WG_CHUNK - 1) / WG_CHUNK;
    if (tid == 0) {
        for (uint32_t s = 0; s < WP_STAGES; s++) { mbar_init(&full[s], WG_THREADS / 32); mbar_init(&empty[s], 1); }      // one arrival per producer warp
        for (uint32_t a = 0; a < 2; a++) { mbar_init(&acc_full[a], 1); mbar_init(&acc_free[a], 128); }
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 256);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t base_a = smem_u32(base);
    // Chunk length per job: a stage holds 64 samples of up to 128 features per operand, or — when both operands are at most 64 features wide (one
    // 128-byte row per sample) — 128 samples in the same 16 KB.  The stream is bound by (load latency / chunks in flight) per chunk, whatever the
    // chunk's size, so the narrow jobs (10 of the 13 products of a step) take half as many ring trips.
    auto chunk_of = [&](uint32_t job) { return (J.j[job].out_dim <= 64u && J.j[job].in_dim <= 64u) ? 2u * WG_CHUNK : WG_CHUNK; };
    auto my_chunks = [&](uint32_t job) {                                  // chunks of this job walked by this CTA: blockIdx.x, + gridDim.x, ...
        const uint32_t nc = (J.M + chunk_of(job) - 1) / chunk_of(job);
        return blockIdx.x < nc ? (nc - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    };
    uint32_t total = 0;
    for (uint32_t job = 0; job < J.n; job++) total += my_chunks(job);
    (void)n_chunks;
    constexpr uint32_t AHEAD = WP_AHEAD;
    // stale bytes of a wider previous job may sit beyond a job's operand widths: harmless, D[o, i] only depends on column o of A and column i of B,
    // and the epilogue writes o < out_dim, i < in_dim only
    if (tid < WG_THREADS) {
        // Everything about a thread's copies that does not change from chunk to chunk of a job — which rows / 16-byte column it owns, their swizzled
        // shared-memory offsets, their element offsets inside a chunk — is computed when the issue cursor enters the job; a chunk then costs one
        // multiply-add and one compare per cp.async (the per-chunk index arithmetic was what bound this kernel: ~260 instructions per warp and chunk).
        struct Side { const __half *src; uint32_t width, k[4], soff[4], goff[4]; };
        Side A, B;
        uint32_t pj = 0, pi = 0, pn = 0, ch = WG_CHUNK;                    // cursor of the next chunk to issue
        auto setup_side = [&](Side &sd, const void *ptr, uint32_t width) {
            const RowMap r = row_map(width);
            sd.src = reinterpret_cast<const __half *>(ptr); sd.width = width;
            const uint32_t j = r.col * 8u;
#pragma unroll
            for (uint32_t u = 0; u < 4; u++) {
                const uint32_t k = r.r0 + u * r.rstep;
                const bool act = r.col < r.nv && k < ch;
                sd.k[u] = act ? k : 0xffffffffu;
                sd.soff[u] = act ? mn_offset(k, j) : 0u;
                sd.goff[u] = k * width + j;
            }
        };
        auto setup_job = [&]() {
            ch = chunk_of(pj); pn = my_chunks(pj);
            setup_side(A, J.j[pj].dy, J.j[pj].out_dim);
            setup_side(B, J.j[pj].x, J.j[pj].in_dim);
        };
        auto issue_side = [&](const Side &sd, uint32_t st, uint32_t row0) {
            const __half *p0 = sd.src + (size_t)row0 * sd.width;
#pragma unroll
            for (uint32_t u = 0; u < 4; u++)
                if (sd.k[u] != 0xffffffffu) {
                    const bool ok = row0 + sd.k[u] < J.M;
                    cp_async16(st + sd.soff[u], ok ? p0 + sd.goff[u] : sd.src, ok);
                }
        };
        setup_job();
        auto issue = [&](uint32_t g) {
            while (pi >= pn) { pj++; pi = 0; setup_job(); }
            const uint32_t st = base_a + (g % WP_STAGES) * WP_STAGE_BYTES, row0 = (blockIdx.x + pi * gridDim.x) * ch;
            issue_side(A, st, row0);
            issue_side(B, st + WG_OPERAND_BYTES, row0);
            pi++;
        };
        for (uint32_t g = 0; g < AHEAD; g++) { if (g < total) issue(g); cp_async_commit(); }
        for (uint32_t g = 0; g < total; g++) {
            cp_async_wait<AHEAD - 1>();
            fence_proxy_async();
            __syncwarp();
            if ((tid & 31u) == 0) mbar_arrive(&full[g % WP_STAGES]);         // 8 arrivals per chunk instead of 256 on one shared-memory word
            const uint32_t nxt = g + AHEAD;
            if (nxt < total) {
                if (nxt >= WP_STAGES) mbar_wait(&empty[nxt % WP_STAGES], ((nxt / WP_STAGES) - 1u) & 1u);
                issue(nxt);
            }
            cp_async_commit();
        }
    } else if (tid == WG_THREADS) {                                       // the MMA thread
        uint32_t g = 0;
        for (uint32_t job = 0; job < J.n; job++) {
            const uint32_t a = job & 1u, n_pad = (J.j[job].in_dim + 15u) & ~15u;
            const uint32_t idesc = idesc_f16(128, n_pad) | (1u << 15) | (1u << 16);
            const uint32_t my_n = my_chunks(job), ksteps = chunk_of(job) / 16u;
            if (job >= 2) { mbar_wait(&acc_free[a], ((job >> 1) - 1u) & 1u); fence_after_sync(); }      // the epilogue of job - 2 has drained this accumulator
            for (uint32_t i = 0; i < my_n; i++, g++) {
                const uint32_t s = g % WP_STAGES;
                mbar_wait(&full[s], (g / WP_STAGES) & 1u);
                fence_after_sync();
                uint64_t da = smem_desc_mn_sw128(base_a + s * WP_STAGE_BYTES), db = smem_desc_mn_sw128(base_a + s * WP_STAGE_BYTES + WG_OPERAND_BYTES);
#pragma unroll 1
                for (uint32_t k = 0; k < ksteps; k++, da += 2048u >> 4, db += 2048u >> 4) mma_f16_ss(tmem + a * 128u, da, db, idesc, i > 0 || k > 0);
                mma_commit(&empty[s]);
            }
            mma_commit(&acc_full[a]);                                     // arrives when every MMA of this job has completed
        }
    } else if (tid >= WG_THREADS + 32) {                                  // epilogue warps: TMEM lane quarter = warp & 3
        const uint32_t o = (warp & 3u) * 32u + (tid & 31u);
        for (uint32_t job = 0; job < J.n; job++) {
            const uint32_t a = job & 1u;
            const b2n_wgrad_job &q = J.j[job];
            const uint32_t n_pad = (q.in_dim + 15u) & ~15u;
            float *dw = q.dw + (size_t)(blockIdx.x % J.replicas) * J.rstride;
            const bool vec4 = (q.in_dim & 3u) == 0 && ((uintptr_t)dw & 15u) == 0;
            const bool any = my_chunks(job) > 0;                          // no chunk of this job here: the accumulator holds nothing (keep the handshake)
            mbar_wait(&acc_full[a], (job >> 1) & 1u);
            fence_after_sync();
            const uint32_t taddr = tmem + a * 128u + (((warp & 3u) * 32u) << 16);
            for (uint32_t cb = 0; any && cb < n_pad; cb += 16) {
                uint32_t acc[16];
                ld16(taddr + cb, acc);
                wait_ld();
                if (o < q.out_dim) {
                    float *row = dw + (size_t)o * q.in_dim + cb;
                    if (vec4) {
#pragma unroll
                        for (uint32_t jj = 0; jj < 16; jj += 4)
                            if (cb + jj < q.in_dim)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + jj), "f"(__uint_as_float(acc[jj])), "f"(__uint_as_float(acc[jj + 1])),
                                             "f"(__uint_as_float(acc[jj + 2])), "f"(__uint_as_float(acc[jj + 3])) : "memory");
                    } else {
#pragma unroll
                        for (uint32_t jj = 0; jj < 16; jj++)
                            if (cb + jj < q.in_dim) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(row + jj), "f"(__uint_as_float(acc[jj])) : "memory");
                    }
                }
            }
            fence_before_sync();
            mbar_arrive(&acc_free[a]);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 256);
}

// halves per global access: 8/4/2 when rows allow it, 0 = flat 16-byte stream (odd width, 16-byte aligned matrix), 1 = element-wise
static uint32_t vec_width(const void *p, uint32_t width) {
    for (uint32_t v = 8; v > 1; v >>= 1)
        if (width % v == 0 && ((uintptr_t)p % (2 * v)) == 0) return v;
    return ((uintptr_t)p % 16) == 0 ? 0 : 1;
}

}  // namespace b2n

using namespace b2n;

static int wgrad_launch(const void *dy, const void *x, uint32_t M, uint32_t out_dim, uint32_t in_dim, float *dw, uint32_t replicas, uint32_t rstride, void *stream);

extern "C" int b2n_linear_wgrad(const void *dy, const void *x, uint32_t M, uint32_t out_dim, uint32_t in_dim, float *dw, void *stream) {
    return wgrad_launch(dy, x, M, out_dim, in_dim, dw, 1, out_dim * in_dim, stream);
}
extern "C" int b2n_linear_wgrad_replicated(const void *dy, const void *x, uint32_t M, uint32_t out_dim, uint32_t in_dim, float *dw, uint32_t replicas,
                                           uint32_t replica_stride, void *stream) {
    B2N_REQUIRE(replicas >= 1 && replicas <= 64, "linear_wgrad_replicated: replicas=%u out of range (1..64)", replicas);
    B2N_REQUIRE(replica_stride == 0 || replica_stride >= out_dim * in_dim, "linear_wgrad_replicated: replica stride smaller than the matrix");
    return wgrad_launch(dy, x, M, out_dim, in_dim, dw, replicas, replica_stride ? replica_stride : out_dim * in_dim, stream);
}

static int wgrad_launch(const void *dy, const void *x, uint32_t M, uint32_t out_dim, uint32_t in_dim, float *dw, uint32_t replicas, uint32_t rstride, void *stream) {
    B2N_REQUIRE(dy && x && dw, "linear_wgrad: null pointer");
    B2N_REQUIRE(out_dim >= 1 && out_dim <= 128 && in_dim >= 1 && in_dim <= 128, "linear_wgrad: out=%u / in=%u unsupported (1..128)", out_dim, in_dim);
    B2N_REQUIRE(((uintptr_t)dy & 1) == 0 && ((uintptr_t)x & 1) == 0, "linear_wgrad: operands must be 2-byte aligned");
    if (M == 0) return 0;
    uint32_t ctas_per_sm = 3;           // 66 KB of operand buffers + 34 registers: three resident CTAs per SM hide the per-chunk load latency
    B2N_SMEM(k_linear_wgrad, WG_SMEM);
    if (const char *g = getenv("B2N_WGRAD_CTAS_PER_SM")) ctas_per_sm = (uint32_t)atoi(g);
    const uint32_t n_chunks = ceil_div<uint32_t>(M, WG_CHUNK);
    uint32_t ctas = ctas_per_sm * (uint32_t)sm_count();
    if (ctas > n_chunks) ctas = n_chunks;
    const uint32_t va = vec_width(dy, out_dim), vb = vec_width(x, in_dim);
    if (va == 8 && vb == 8 && !getenv("B2N_WGRAD_SIMPLE")) {              // 16-byte aligned operands: cp.async ring, one CTA per SM
        B2N_SMEM(k_linear_wgrad_pipe, WP_SMEM);
        uint32_t g = (uint32_t)sm_count();
        if (g > n_chunks) g = n_chunks;
        k_linear_wgrad_pipe<<<g, WP_THREADS, WP_SMEM, as_stream(stream)>>>((const __half *)dy, (const __half *)x, M, out_dim, in_dim, dw, replicas, rstride);
        return check_launch("linear_wgrad");
    }
    k_linear_wgrad<<<ctas, WG_THREADS, WG_SMEM, as_stream(stream)>>>((const __half *)dy, (const __half *)x, M, out_dim, in_dim, va, vb, dw, replicas, rstride);
    return check_launch("linear_wgrad");
}

extern "C" int b2n_linear_wgrad_batch(const b2n_wgrad_job *jobs, uint32_t n_jobs, uint32_t M, uint32_t replicas, uint32_t replica_stride, void *stream) {
    B2N_REQUIRE(jobs && n_jobs >= 1 && n_jobs <= 16, "linear_wgrad_batch: 1..16 jobs");
    B2N_REQUIRE(replicas >= 1 && replicas <= 64, "linear_wgrad_batch: replicas=%u out of range (1..64)", replicas);
    if (M == 0) return 0;
    WgradJobs J = {};
    J.n = n_jobs; J.M = M; J.replicas = replicas; J.rstride = replica_stride;
    for (uint32_t i = 0; i < n_jobs; i++) {
        const b2n_wgrad_job &q = jobs[i];
        B2N_REQUIRE(q.dy && q.x && q.dw, "linear_wgrad_batch: null pointer in job %u", i);
        B2N_REQUIRE(q.out_dim >= 8 && q.out_dim <= 128 && q.in_dim >= 8 && q.in_dim <= 128 && q.out_dim % 8 == 0 && q.in_dim % 8 == 0,
                    "linear_wgrad_batch: job %u: out=%u / in=%u must be multiples of 8 in 8..128 (16-byte rows)", i, q.out_dim, q.in_dim);
        B2N_REQUIRE(((uintptr_t)q.dy & 15) == 0 && ((uintptr_t)q.x & 15) == 0, "linear_wgrad_batch: job %u: operands must be 16-byte aligned", i);
        B2N_REQUIRE(replicas == 1 || replica_stride >= q.out_dim * q.in_dim, "linear_wgrad_batch: replica stride smaller than matrix %u", i);
        J.j[i] = q;
    }
    B2N_SMEM(k_linear_wgrad_multi, WM_SMEM);
    const uint32_t n_chunks = ceil_div<uint32_t>(M, WG_CHUNK);
    uint32_t g = (uint32_t)sm_count();
    if (g > n_chunks) g = n_chunks;
    k_linear_wgrad_multi<<<g, WM_THREADS, WM_SMEM, as_stream(stream)>>>(J);
    return check_launch("linear_wgrad_batch");
}

// ---- finishing pass of the replicated weight gradients: sum the replicas and ACCUMULATE rectangular blocks straight into the parameters' .grad
// storage (row rotations / column slices of the padded products are expressed as blocks), so no slicing / cat / add kernels follow a backward.
namespace b2n {
struct ScatterJobs { b2n_wgrad_block j[32]; uint32_t n, replicas, rstride; };
__global__ void __launch_bounds__(256) k_wgrad_scatter(const __grid_constant__ ScatterJobs J, const float *__restrict__ src) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t k = 0; k < J.n; k++) {
        const b2n_wgrad_block &q = J.j[k];
        const uint32_t n = q.rows * q.cols;
        if (e < n) {
            const uint32_t i = e / q.cols, c = e - i * q.cols;
            const float *p = src + q.src_off + (size_t)i * q.src_ld + c;
            float a = 0.0f;
            for (uint32_t r = 0; r < J.replicas; r++) a += __ldg(p + (size_t)r * J.rstride);
            q.dst[(size_t)i * q.dst_ld + c] += a;
            return;
        }
        e -= n;
    }
}
}  // namespace b2n

extern "C" int b2n_wgrad_scatter(const float *src, uint32_t replicas, uint32_t replica_stride, const b2n_wgrad_block *blocks, uint32_t n_blocks, void *stream) {
    B2N_REQUIRE(src && blocks, "wgrad_scatter: null pointer");
    B2N_REQUIRE(n_blocks >= 1 && n_blocks <= 32, "wgrad_scatter: 1..32 blocks");
    B2N_REQUIRE(replicas >= 1 && replicas <= 64, "wgrad_scatter: replicas=%u out of range (1..64)", replicas);
    ScatterJobs J = {};
    J.n = n_blocks; J.replicas = replicas; J.rstride = replica_stride;
    uint32_t total = 0;
    for (uint32_t i = 0; i < n_blocks; i++) {
        B2N_REQUIRE(blocks[i].dst && blocks[i].cols >= 1 && blocks[i].rows >= 1, "wgrad_scatter: block %u is empty", i);
        J.j[i] = blocks[i];
        total += blocks[i].rows * blocks[i].cols;
    }
    k_wgrad_scatter<<<ceil_div<uint32_t>(total, 256), 256, 0, as_stream(stream)>>>(J, src);
    return check_launch("wgrad_scatter");
}
