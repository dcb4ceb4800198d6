// peer_allreduce.cu — the data-parallel gradient exchange of a training step (ONE all-reduce of the flat fp32 gradient buffer, 683 509 floats =
// 2.73 MB) as ONE kernel over NVLink / NVSwitch peer memory.
//
// The payload is latency-bound (2.7 MB against 900 GB/s per direction), and a library all-reduce inside the step's CUDA graph costs ~75 us, fully exposed after
// the last backward kernel (SCALE_r01: 0.88 -> 0.958 ms per step at 8 GPUs).  Here every rank's gradient buffer is a cudaMalloc block opened on all peers through
// CUDA IPC (one process per GPU), and one kernel per rank does a TWO-SHOT all-reduce in place:
//     barrier (all ranks' gradients are final)  ->  rank r sums element i of ALL ranks for the i of its 1/world slice, scales by 1/world and writes the result
//     back into ALL ranks' buffers  ->  barrier (all writes have landed).
// Element i of any buffer is read and then written by exactly one thread in the whole job (the owner of its slice), so the exchange needs no scratch copy.
// Per rank and call: (world - 1)/world * 2.73 MB read + as much written over NVLink, two flag barriers (one store to every peer + a spin on local memory each).
// Barriers use monotonically increasing epochs (no reset races); a spin that sees no progress for ~2 s sets an error word and gives up instead of hanging the GPU.
// An inf / nan produced by any rank propagates through the sum, so every rank's GradScaler skips the same steps (the behaviour of the NCCL path).
#include "common.cuh"

struct b2n_peer_comm {
    int rank, world;
    float *data[16];             // every rank's gradient buffer (data[rank] is local)
    uint32_t *flags[16];         // every rank's flag block: [0..15] start-barrier epochs, [16..31] end-barrier epochs, [32] epoch, [33] done-CTA ticket, [34] error
    void *opened[16];            // IPC mappings to close
    void *local_base;
    uint64_t bytes;
};

namespace b2n {

constexpr uint32_t PF_START = 0, PF_END = 16, PF_EPOCH = 32, PF_TICKET = 33, PF_ERROR = 34, PF_WORDS = 64;

struct PeerArgs { float *data[16]; uint32_t *flags[16]; int rank, world; uint64_t n4; float inv_world; };

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ float4 ld_sys_f4(const float4 *p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_f4(float4 *p, const float4 &v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// wait until the local flag words [base, base + world) have all reached `epoch` (peers store them); false on timeout
__device__ __forceinline__ bool wait_flags(const uint32_t *flags, uint32_t base, int world, uint32_t epoch, uint32_t *err) {
    if (threadIdx.x < (uint32_t)world) {
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(flags + base + threadIdx.x) - epoch) < 0) {
            if (clock64() - t0 > 4000000000ll) { atomicExch(err, 1u); break; }       // ~2 s at 1.9 GHz: a peer never arrived
        }
    }
    __syncthreads();
    return true;
}

__global__ void __launch_bounds__(256) k_peer_allreduce(const __grid_constant__ PeerArgs a) {
    uint32_t *my = a.flags[a.rank];
    const uint32_t e0 = my[PF_EPOCH];             // stable for the whole launch: only the last CTA to finish advances it
    // ---- barrier 1: every rank's gradients are final (its backward kernels precede this kernel in its stream) -----------------------
    if (blockIdx.x == 0 && threadIdx.x < (uint32_t)a.world) st_release_sys(a.flags[threadIdx.x] + PF_START + a.rank, e0 + 1u);
    wait_flags(my, PF_START, a.world, e0 + 1u, my + PF_ERROR);
    // ---- reduce my slice over all ranks, write it back to all ranks -----------------------------------------------------------------
    const uint64_t per = (a.n4 + a.world - 1) / a.world, lo = per * a.rank, hi = min(lo + per, a.n4);
    for (uint64_t i = lo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < hi; i += (uint64_t)gridDim.x * blockDim.x) {
        float4 v[16];
#pragma unroll
        for (int p = 0; p < 16; p++) if (p < a.world) v[p] = ld_sys_f4(reinterpret_cast<const float4 *>(a.data[p]) + i);
        float4 s = v[0];                           // fixed summation order (rank 0, 1, ...): every rank's copy of the result is the same bits
#pragma unroll
        for (int p = 1; p < 16; p++) if (p < a.world) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }
        s.x *= a.inv_world; s.y *= a.inv_world; s.z *= a.inv_world; s.w *= a.inv_world;
#pragma unroll
        for (int p = 0; p < 16; p++) if (p < a.world) st_sys_f4(reinterpret_cast<float4 *>(a.data[p]) + i, s);
    }
    // ---- barrier 2: all my writes have landed everywhere; all peers' writes have landed here ------------------------------------------
    __threadfence_system();
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) s_last = (atomicAdd(my + PF_TICKET, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if (threadIdx.x < (uint32_t)a.world) st_release_sys(a.flags[threadIdx.x] + PF_END + a.rank, e0 + 2u);
    wait_flags(my, PF_END, a.world, e0 + 2u, my + PF_ERROR);
    if (threadIdx.x == 0) { my[PF_TICKET] = 0; my[PF_EPOCH] = e0 + 2u; __threadfence(); }
}

}  // namespace b2n

using namespace b2n;

extern "C" int b2n_peer_alloc(uint64_t bytes, void **dev_ptr, void *ipc_handle_out) {
    B2N_REQUIRE(dev_ptr && ipc_handle_out && bytes > 0, "peer_alloc: bad argument");
    const size_t total = ((size_t)bytes + 255) / 256 * 256 + PF_WORDS * sizeof(uint32_t);
    void *p = nullptr;
    B2N_CUDA(cudaMalloc(&p, total));
    B2N_CUDA(cudaMemset(p, 0, total));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { (void)cudaGetLastError(); cudaFree(p); set_error("peer_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return 3; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(ipc_handle_out, &h, 64);
    *dev_ptr = p;
    return 0;
}

extern "C" int b2n_peer_free(void *dev_ptr) {
    if (dev_ptr) B2N_CUDA(cudaFree(dev_ptr));
    return 0;
}

extern "C" int b2n_peer_comm_create(b2n_peer_comm **out, int rank, int world, void *local_ptr, const void *all_handles, uint64_t bytes) {
    B2N_REQUIRE(out && local_ptr && all_handles, "peer_comm_create: null pointer");
    B2N_REQUIRE(world >= 1 && world <= 16 && rank >= 0 && rank < world, "peer_comm_create: rank %d of %d (1..16 ranks)", rank, world);
    B2N_REQUIRE(bytes % 16 == 0, "peer_comm_create: the buffer must be a multiple of 16 bytes");
    b2n_peer_comm *c = new b2n_peer_comm();
    c->rank = rank; c->world = world; c->bytes = bytes; c->local_base = local_ptr;
    const size_t flag_off = ((size_t)bytes + 255) / 256 * 256;
    for (int p = 0; p < world; p++) {
        void *base = local_ptr;
        c->opened[p] = nullptr;
        if (p != rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const char *)all_handles + 64 * (size_t)p, 64);
            cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                (void)cudaGetLastError();
                for (int q = 0; q < p; q++) if (c->opened[q]) cudaIpcCloseMemHandle(c->opened[q]);
                delete c;
                set_error("peer_comm_create: cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(e));
                return 3;
            }
            c->opened[p] = base;
        }
        c->data[p] = (float *)base;
        c->flags[p] = (uint32_t *)((char *)base + flag_off);
    }
    *out = c;
    return 0;
}

// Ranks as streams of ONE process (tests on a single GPU, or a single-process multi-GPU driver with peer access enabled): the buffers of all ranks are plain
// device pointers of this process (each from b2n_peer_alloc), no IPC.
extern "C" int b2n_peer_comm_create_local(b2n_peer_comm **out, int rank, int world, void *const *ptrs, uint64_t bytes) {
    B2N_REQUIRE(out && ptrs, "peer_comm_create_local: null pointer");
    B2N_REQUIRE(world >= 1 && world <= 16 && rank >= 0 && rank < world, "peer_comm_create_local: rank %d of %d (1..16 ranks)", rank, world);
    B2N_REQUIRE(bytes % 16 == 0, "peer_comm_create_local: the buffer must be a multiple of 16 bytes");
    b2n_peer_comm *c = new b2n_peer_comm();
    c->rank = rank; c->world = world; c->bytes = bytes; c->local_base = ptrs[rank];
    const size_t flag_off = ((size_t)bytes + 255) / 256 * 256;
    for (int p = 0; p < world; p++) {
        c->opened[p] = nullptr;
        c->data[p] = (float *)ptrs[p];
        c->flags[p] = (uint32_t *)((char *)ptrs[p] + flag_off);
    }
    *out = c;
    return 0;
}

extern "C" void b2n_peer_comm_destroy(b2n_peer_comm *c) {
    if (!c) return;
    for (int p = 0; p < c->world; p++) if (c->opened[p]) cudaIpcCloseMemHandle(c->opened[p]);
    delete c;
}

// in place on every rank's buffer: data[i] <- (sum over ranks of data[i]) / world for the first n_floats floats (rounded up to a multiple of 4; the buffer is
// padded).  Every rank must call it the same number of times with the same n_floats.
extern "C" int b2n_peer_allreduce_mean(b2n_peer_comm *c, uint64_t n_floats, void *stream) {
    B2N_REQUIRE(c, "peer_allreduce_mean: null communicator");
    B2N_REQUIRE(n_floats * 4 <= c->bytes, "peer_allreduce_mean: %llu floats exceed the buffer", (unsigned long long)n_floats);
    if (n_floats == 0 || c->world == 1) return 0;
    PeerArgs a;
    for (int p = 0; p < 16; p++) { a.data[p] = p < c->world ? c->data[p] : nullptr; a.flags[p] = p < c->world ? c->flags[p] : nullptr; }
    a.rank = c->rank; a.world = c->world; a.n4 = (n_floats + 3) / 4; a.inv_world = 1.0f / (float)c->world;
    const uint64_t per = (a.n4 + c->world - 1) / c->world;
    uint32_t blocks = (uint32_t)((per + 255) / 256);
    const uint32_t cap = (uint32_t)sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    k_peer_allreduce<<<blocks, 256, 0, as_stream(stream)>>>(a);
    return check_launch("peer_allreduce_mean");
}

// error word of the local flag block (non-zero: a barrier timed out — some rank did not launch the matching call); synchronises the stream
extern "C" int b2n_peer_error(b2n_peer_comm *c, int32_t *host_out, void *stream) {
    B2N_REQUIRE(c && host_out, "peer_error: null pointer");
    B2N_CUDA(cudaMemcpyAsync(host_out, c->flags[c->rank] + PF_ERROR, sizeof(int32_t), cudaMemcpyDeviceToHost, as_stream(stream)));
    B2N_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return 0;
}
