// core.cu — library plumbing: error string, launch counter, SM count, internal scratch.
#include "common.cuh"
#include <map>
#include <mutex>
#include <utility>
#include <string.h>

namespace b2n {

static thread_local char t_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int ensure_dynamic_smem_impl(const void *kernel, size_t bytes) {
    // (no shortcut for small sizes: static + dynamic shared memory together may exceed the 48 KB default even when the dynamic part alone does not)
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, size_t> configured;
    int dev = 0;
    B2N_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    size_t &have = configured[{dev, kernel}];
    if (have < bytes) {
        B2N_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        have = bytes;
    }
    return 0;
}

// Grow-only per-device scratch slots for the entry points that keep the reference's argument lists (no workspace parameter).  Growth happens only when a call
// needs more than any earlier call (never in steady state, never under stream capture after a warm-up call of the same shape).  A slot is shared by every
// caller on the device: calls that use it must be stream-ordered with each other; concurrent callers use the *_ws entry points with their own workspace.
struct Slot { void *ptr = nullptr; size_t bytes = 0; };
static Slot g_slots[64][4];
static std::mutex g_slot_mu;

void *scratch(size_t bytes, int slot) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || slot < 0 || slot >= 4) return nullptr;
    std::lock_guard<std::mutex> lk(g_slot_mu);
    Slot &s = g_slots[dev][slot];
    if (s.bytes < bytes) {
        size_t want = bytes < 4096 ? 4096 : bytes + bytes / 2;
        void *p = nullptr;
        if (cudaMalloc(&p, want) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
        // the old block is RETIRED, not freed: a CUDA graph captured earlier may still replay kernels that point into it, and another stream may still be
        // running on it.  Blocks grow geometrically, so the retired ones add up to less than twice the live one.
        s.ptr = p; s.bytes = want;
    }
    return s.ptr;
}

}  // namespace b2n

extern "C" {
int b2n_version(void) { return B2N_VERSION; }
const char *b2n_last_error(void) { return b2n::t_err; }
uint64_t b2n_launch_count(void) { return b2n::g_launches.load(std::memory_order_relaxed); }
}
