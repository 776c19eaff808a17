// jz_common.cuh -- shared internals of libjz_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include <atomic>

#include "../../include/jz_b200.h"

namespace jz {

struct Ctx {
    bool inited = false;
    int device = -1;
    int sm_count = 148;
    int cc_major = 0, cc_minor = 0;
    size_t total_mem = 0;
    std::atomic<uint64_t> launches{0};
    int gemm_mode = JZ_GEMM_3XTF32;
    int gemm_last_path = 0;
    int gemm_last_splits = 1;   // k-splits per tail tile of the last tensor-core launch (1 = none)
    int gemm_last_walk = 0;            // 1: the last strided batch was walked (one CTA group per SM over the (member, tile) units)
    int gemm_last_cluster_split = 0;   // 1: those splits met through distributed shared memory (one cluster per tile)
};

Ctx& ctx();
int ensure_init();                                  // lazy jz_init(current device)
int fail(int code, const char* fmt, ...);           // records jz_last_error(), returns code
int cuda_fail(cudaError_t e, const char* what);     // JZ_ERR_CUDA with the CUDA error string

inline cudaStream_t as_stream(jz_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }

// Zero-initialised ticket counters (kTicketSlots unsigned), one block per (device, stream).  Kernels on a stream are
// serialised and every kernel that uses the block leaves its counters at zero again, so the block is never shared by
// two running kernels and never needs a memset after the first use.  nullptr on allocation failure.
constexpr size_t kTicketSlots = 4096;
unsigned* tickets_for(cudaStream_t s);

// temporary workspace from the stream-ordered pool (bytes), released with ws_free
int ws_alloc(void** p, size_t bytes, cudaStream_t s);
int ws_free(void* p, cudaStream_t s);
// scope guard: the workspace goes back to the pool on every exit path (JZ_LAUNCH / JZ_CUDA return early on errors)
struct WsGuard {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    explicit WsGuard(cudaStream_t stream) : s(stream) {}
    ~WsGuard() { if (p) ws_free(p, s); }
    WsGuard(const WsGuard&) = delete;
    WsGuard& operator=(const WsGuard&) = delete;
};

// Programmatic dependent launch for the small kernels.  A step of a training loop is a chain of dependent launches of a
// few microseconds each, and between two of them the GPU idles for the launch latency (measured: 2.8 - 4.1 us per
// dependent launch of a <= 1184-block kernel in stream order, 2.35 us with programmatic stream serialization).  Every
// kernel of this library starts with pdl_enter(): griddepcontrol.wait (all prerequisite grids complete, their memory
// visible) before its first global-memory access, then griddepcontrol.launch_dependents, so the NEXT kernel of the
// stream may be scheduled and run its own preamble while this one executes; it blocks in its own wait until this grid
// has finished.  Kernels without the launch attribute (the reference's own, the header functor kernels) are unaffected.
// Large grids are launched without the attribute: their dependents would only take SM slots from their last wave
// (8192 blocks: 8.2 -> 9.2 us).  JZ_NO_PDL=1 disables it.
bool pdl_on();
constexpr unsigned kPdlMaxBlocks = 148 * 16;
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <class... KArgs, class... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl_on() && size_t(grid.x) * grid.y * grid.z <= kPdlMaxBlocks) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(static_cast<Args&&>(args))...);
}
#endif

}  // namespace jz

// launch + count + error check.  Every kernel of this library goes through here so
// jz_launch_count() is an honest count of OUR launches.
#define JZ_LAUNCH(kernel, grid, block, smem, stream, ...)                          \
    do {                                                                           \
        ::jz::launch_kernel(kernel, dim3(grid), dim3(block), (smem), (stream), __VA_ARGS__); \
        ::jz::ctx().launches.fetch_add(1, std::memory_order_relaxed);              \
        cudaError_t e__ = cudaPeekAtLastError();                                   \
        if (e__ != cudaSuccess) return ::jz::cuda_fail(cudaGetLastError(), #kernel); \
    } while (0)

#define JZ_INIT_OR_RETURN()                      \
    do {                                         \
        int rc__ = ::jz::ensure_init();          \
        if (rc__ != JZ_OK) return rc__;          \
    } while (0)

#define JZ_CUDA(call)                                              \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return ::jz::cuda_fail(e__, #call); \
    } while (0)
