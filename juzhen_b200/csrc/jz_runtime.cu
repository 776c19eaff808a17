// jz_runtime.cu -- lifetime, error reporting, copies and the stream-ordered pool.
//
// Replaces the reference's process-scope setup (cpp/launcher.cu:44-101: cuBLAS handle +
// RAII Memory<CUDAfloat>) and its exact-size free-list pool (cpp/memory.hpp:50-119 over
// cudaMalloc/cudaFree, cpp/cumatrix.cuh:67-85).
//
// Pool design (B200: 180 GB HBM3e, cudaMalloc is ~100 us-1 ms so temporaries must never
// reach the driver in steady state):
//   * sizes are rounded to 512 B; a freed block goes to a per-size free list (the
//     reference's "exact size match" rule, which is what makes rvalue chains hit).
//   * STREAM-ORDERED: a block remembers the stream it was freed on.  Re-allocation on the
//     same stream is immediate (program order on the stream protects it); re-allocation
//     on a different stream first makes that stream wait on an event recorded on the
//     freeing stream, so no host synchronisation is ever needed.
//   * on cudaMalloc failure the cached blocks are released to the driver and the
//     allocation retried once (the reference never trims; with 4 GiB temporaries at
//     32768^2 that matters).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "jz_common.cuh"

namespace jz {

bool pdl_on() {
    static const bool on = [] {
        const char* e = std::getenv("JZ_NO_PDL");
        const char* g = std::getenv("JZ_GEMM_NO_PDL");   // older name, still honoured
        return !((e && e[0] && e[0] != '0') || (g && g[0] && g[0] != '0'));
    }();
    return on;
}


static Ctx g_ctx;
static thread_local char g_err[512] = "";

Ctx& ctx() { return g_ctx; }

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    return fail(e == cudaErrorMemoryAllocation ? JZ_ERR_OOM : JZ_ERR_CUDA, "CUDA error: %s (%s)",
                cudaGetErrorString(e), what);
}

static std::mutex g_init_mu;
static int pool_prepare_device_switch();
static void pool_mark_alive();

static int do_init(int device) {
    std::lock_guard<std::mutex> lk(g_init_mu);
    if (g_ctx.inited && (device < 0 || device == g_ctx.device)) return JZ_OK;
    if (g_ctx.inited) {
        // the pool and the ticket blocks belong to the device they were allocated on (one process per GPU is the
        // model): switching is allowed only when nothing is live, and the cached blocks go back to the old device
        int rc = pool_prepare_device_switch();
        if (rc != JZ_OK) return rc;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(JZ_ERR_CUDA,
                    "no CUDA device available (%s): libjz_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= count) return fail(JZ_ERR_ARG, "device %d out of range (%d devices)", device, count);
    JZ_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    JZ_CUDA(cudaGetDeviceProperties(&p, device));
    g_ctx.device = device;
    g_ctx.sm_count = p.multiProcessorCount;
    g_ctx.cc_major = p.major;
    g_ctx.cc_minor = p.minor;
    g_ctx.total_mem = p.totalGlobalMem;
    // GEMM mode: NVIDIA_TF32=1 keeps the reference's spelling (cpp/launcher.cu:73-77)
    const char* tf32 = std::getenv("NVIDIA_TF32");
    if (tf32 && *tf32 && std::strcmp(tf32, "0") != 0) g_ctx.gemm_mode = JZ_GEMM_TF32;
    const char* gm = std::getenv("JZ_GEMM_MODE");
    if (gm && *gm) {
        if (!std::strcmp(gm, "3xtf32")) g_ctx.gemm_mode = JZ_GEMM_3XTF32;
        else if (!std::strcmp(gm, "tf32")) g_ctx.gemm_mode = JZ_GEMM_TF32;
        else if (!std::strcmp(gm, "fp32")) g_ctx.gemm_mode = JZ_GEMM_FP32_SIMT;
    }
    pool_mark_alive();
    g_ctx.inited = true;
    return JZ_OK;
}

int ensure_init() {
    if (g_ctx.inited) {
        // a host thread other than the one that called jz_init starts on device 0: bind it once
        static thread_local bool bound = false;
        if (!bound) {
            int cur = -1;
            if (cudaGetDevice(&cur) != cudaSuccess || cur != g_ctx.device) JZ_CUDA(cudaSetDevice(g_ctx.device));
            bound = true;
        }
        return JZ_OK;
    }
    return do_init(-1);
}

// ------------------------------------------------------------------ pool
struct Block {
    void* ptr;
    size_t bytes;
    cudaStream_t freed_on;
    cudaEvent_t freed_ev;   // recorded on freed_on when the block was released (nullptr: could not be recorded)
};

struct Pool {
    std::mutex mu;
    std::unordered_map<void*, size_t> live;                 // ptr -> rounded bytes
    std::unordered_map<size_t, std::vector<Block>> cached;  // rounded bytes -> free blocks
    std::vector<cudaEvent_t> events;                        // recycled events
    size_t live_bytes = 0, cached_bytes = 0, n_device_allocs = 0, n_hits = 0;
    bool torn_down = false;  // set by destroy(): late frees from static destructors are then ignored
    bool have_stream = false, multi_stream = false;
    cudaStream_t first_stream = nullptr;
    void note_stream(cudaStream_t s) {
        if (!have_stream) { have_stream = true; first_stream = s; }
        else if (s != first_stream) multi_stream = true;
    }

    static size_t round_up(size_t bytes) {
        if (bytes == 0) bytes = 4;  // count==0 -> 1 element (cpp/cumatrix.cuh:71)
        return (bytes + 511) & ~size_t(511);
    }

    int trim_locked() {
        for (auto& kv : cached)
            for (auto& b : kv.second) {
                cudaFree(b.ptr);
                if (b.freed_ev) events.push_back(b.freed_ev);
            }
        cached.clear();
        cached_bytes = 0;
        return JZ_OK;
    }

    int alloc(void** out, size_t bytes, cudaStream_t s) {
        const size_t rb = round_up(bytes);
        std::lock_guard<std::mutex> lk(mu);
        auto it = cached.find(rb);
        if (it != cached.end() && !it->second.empty()) {
            // prefer a block freed on the same stream (no cross-stream wait needed)
            auto& vec = it->second;
            size_t pick = vec.size() - 1;
            for (size_t i = vec.size(); i-- > 0;)
                if (vec[i].freed_on == s) { pick = i; break; }
            Block b = vec[pick];
            note_stream(s);
            if (b.freed_on != s) {
                // another stream: wait for the work that was queued on the freeing stream when the block was released.
                // Once the pool has seen two streams every release records that point (below); blocks released
                // earlier fall back to an event recorded on the freeing stream now.  On any failure the block stays
                // cached and the caller gets a fresh allocation -- nothing is lost.
                cudaEvent_t ev = b.freed_ev;
                if (!ev) {
                    if (!events.empty()) { ev = events.back(); events.pop_back(); }
                    else if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); ev = nullptr; }
                    if (ev && cudaEventRecord(ev, b.freed_on) != cudaSuccess) { cudaGetLastError(); events.push_back(ev); ev = nullptr; }
                    if (ev) vec[pick].freed_ev = ev;
                }
                if (!ev || cudaStreamWaitEvent(s, ev, 0) != cudaSuccess) {
                    cudaGetLastError();
                    goto fresh;
                }
                b.freed_ev = ev;
            }
            vec.erase(vec.begin() + pick);
            if (b.freed_ev) events.push_back(b.freed_ev);
            cached_bytes -= rb;
            live_bytes += rb;
            live.emplace(b.ptr, rb);
            n_hits++;
            *out = b.ptr;
            return JZ_OK;
        }
    fresh:
        note_stream(s);
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, rb);
        if (e != cudaSuccess) {
            cudaGetLastError();
            cudaDeviceSynchronize();
            trim_locked();
            e = cudaMalloc(&p, rb);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return fail(JZ_ERR_OOM, "device allocation of %zu bytes failed (%s); live %zu bytes",
                            rb, cudaGetErrorString(e), live_bytes);
            }
        }
        n_device_allocs++;
        live_bytes += rb;
        live.emplace(p, rb);
        *out = p;
        return JZ_OK;
    }

    int release(void* p, cudaStream_t s) {
        if (!p) return JZ_OK;
        std::lock_guard<std::mutex> lk(mu);
        auto it = live.find(p);
        if (it == live.end()) {
            if (torn_down) return JZ_OK;  // block already returned to the driver by jz_shutdown()
            return fail(JZ_ERR_ARG, "jz_free: pointer %p is not a live pool allocation", p);
        }
        const size_t rb = it->second;
        live.erase(it);
        live_bytes -= rb;
        cached_bytes += rb;
        // the point in the freeing stream after which the block may be reused from ANOTHER stream: recorded only once
        // the pool has seen a second stream (an event per release costs a single-stream program -- every
        // Matrix<CUDAfloat> program -- about a microsecond of host time per temporary for nothing)
        note_stream(s);
        cudaEvent_t ev = nullptr;
        if (multi_stream) {
            if (!events.empty()) { ev = events.back(); events.pop_back(); }
            else if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); ev = nullptr; }
            if (ev && cudaEventRecord(ev, s) != cudaSuccess) {   // e.g. a capturing stream: same-stream reuse only
                cudaGetLastError();
                events.push_back(ev);
                ev = nullptr;
            }
        }
        cached[rb].push_back(Block{p, rb, s, ev});
        return JZ_OK;
    }

    int destroy() {
        std::lock_guard<std::mutex> lk(mu);
        cudaDeviceSynchronize();
        trim_locked();
        for (auto& kv : live) cudaFree(kv.first);
        live.clear();
        live_bytes = 0;
        torn_down = true;   // late frees (static destructors) of blocks that are already gone are ignored
        for (auto ev : events) cudaEventDestroy(ev);
        events.clear();
        return JZ_OK;
    }
};

static Pool& pool() {
    static Pool* p = new Pool();  // intentionally leaked: outlives static destructors of callers
    return *p;
}

static int pool_prepare_device_switch() {
    Pool& p = pool();
    std::lock_guard<std::mutex> lk(p.mu);
    if (!p.live.empty()) return fail(JZ_ERR_ARG, "jz_init: cannot switch device while %zu pool blocks are live", p.live.size());
    cudaDeviceSynchronize();
    p.trim_locked();
    return JZ_OK;
}
static void pool_mark_alive() {
    Pool& p = pool();
    std::lock_guard<std::mutex> lk(p.mu);
    p.torn_down = false;   // after jz_shutdown + jz_init, unknown pointers are errors again
}

unsigned* tickets_for(cudaStream_t s) {
    static std::mutex mu;
    static std::unordered_map<unsigned long long, unsigned*> blocks;   // key: (device, stream)
    std::lock_guard<std::mutex> lock(mu);
    const unsigned long long key = (unsigned long long)reinterpret_cast<uintptr_t>(s) * 64ull + (unsigned long long)(ctx().device & 63);
    auto it = blocks.find(key);
    if (it != blocks.end()) return it->second;
    unsigned* p = nullptr;
    if (cudaMalloc(&p, kTicketSlots * sizeof(unsigned)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    // zeroed on the SAME stream: ordered before the first kernel that uses it, whatever kind of stream s is
    if (cudaMemsetAsync(p, 0, kTicketSlots * sizeof(unsigned), s) != cudaSuccess) { cudaGetLastError(); cudaFree(p); return nullptr; }
    blocks[key] = p;
    return p;
}

int ws_alloc(void** p, size_t bytes, cudaStream_t s) { return pool().alloc(p, bytes, s); }
int ws_free(void* p, cudaStream_t s) { return pool().release(p, s); }

}  // namespace jz

using namespace jz;

extern "C" {

int jz_abi_version(void) { return JZ_ABI_VERSION; }

int jz_init(int device) { return do_init(device); }

int jz_shutdown(void) {
    if (!g_ctx.inited) return JZ_OK;
    pool().destroy();
    g_ctx.inited = false;
    return JZ_OK;
}

const char* jz_last_error(void) { return g_err; }

int jz_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
    JZ_INIT_OR_RETURN();
    if (sm_count) *sm_count = g_ctx.sm_count;
    if (cc_major) *cc_major = g_ctx.cc_major;
    if (cc_minor) *cc_minor = g_ctx.cc_minor;
    if (total_mem) *total_mem = g_ctx.total_mem;
    return JZ_OK;
}

int jz_sync(jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    JZ_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return JZ_OK;
}

uint64_t jz_launch_count(void) { return g_ctx.launches.load(); }

int jz_set_gemm_mode(int mode) {
    if (mode < JZ_GEMM_3XTF32 || mode > JZ_GEMM_FP32_SIMT) return fail(JZ_ERR_ARG, "bad gemm mode %d", mode);
    g_ctx.gemm_mode = mode;
    return JZ_OK;
}
int jz_get_gemm_mode(void) { return g_ctx.gemm_mode; }
int jz_gemm_last_path(void) { return g_ctx.gemm_last_path; }

int jz_malloc(float** ptr, size_t count, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    if (!ptr) return fail(JZ_ERR_ARG, "jz_malloc: null out pointer");
    void* p = nullptr;
    int rc = pool().alloc(&p, count * sizeof(float), as_stream(stream));
    *ptr = static_cast<float*>(p);
    return rc;
}

int jz_free(float* ptr, jz_stream_t stream) { return pool().release(ptr, as_stream(stream)); }

int jz_pool_trim(void) {
    JZ_INIT_OR_RETURN();
    std::lock_guard<std::mutex> lk(pool().mu);
    cudaDeviceSynchronize();
    return pool().trim_locked();
}

int jz_pool_stats(size_t* live_bytes, size_t* cached_bytes, size_t* n_device_allocs, size_t* n_hits) {
    Pool& p = pool();
    std::lock_guard<std::mutex> lk(p.mu);
    if (live_bytes) *live_bytes = p.live_bytes;
    if (cached_bytes) *cached_bytes = p.cached_bytes;
    if (n_device_allocs) *n_device_allocs = p.n_device_allocs;
    if (n_hits) *n_hits = p.n_hits;
    return JZ_OK;
}

int jz_memcpy_h2d(float* dst, const float* src, size_t count, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    if (count == 0) return JZ_OK;
    if (!dst || !src) return fail(JZ_ERR_ARG, "jz_memcpy_h2d: null pointer");
    JZ_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(float), cudaMemcpyHostToDevice, as_stream(stream)));
    return JZ_OK;
}

// Upload from PAGEABLE host memory without draining the stream.  cudaMemcpyAsync on pageable memory larger than
// the driver's inline limit makes the host wait until the stream reaches the copy (the driver stages it chunk by
// chunk at execution time) -- for a training loop that uploads a batch per step (examples/demo_mnist.cu:108-109)
// that serialises host and device once per step.  Here the host copies into a slot of a pinned ring first (the
// caller's buffer is free again when this returns, as after a synchronous cudaMemcpy), then the DMA is queued
// behind whatever the stream is doing.  A slot is reused only after the event recorded behind its DMA completed.
namespace {
struct UploadRing {
    static constexpr int kSlots = 16;
    static constexpr size_t kSlotBytes = size_t(1) << 19;   // 512 KiB: a 784 x 32 fp32 batch is 98 KiB
    void* base = nullptr;
    cudaEvent_t done[kSlots] = {};
    bool used[kSlots] = {};
    int next = 0;
    int device = -1;
};
UploadRing g_ring;
std::mutex g_ring_mu;
}  // namespace

int jz_upload(float* dst, const float* src, size_t count, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    if (count == 0) return JZ_OK;
    if (!dst || !src) return fail(JZ_ERR_ARG, "jz_upload: null pointer");
    const size_t bytes = count * sizeof(float);
    if (bytes > UploadRing::kSlotBytes) {   // large: the plain path (host waits for the stream, data safe on return)
        JZ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
        return JZ_OK;
    }
    std::lock_guard<std::mutex> lock(g_ring_mu);
    UploadRing& r = g_ring;
    if (!r.base || r.device != g_ctx.device) {
        if (r.base) {
            cudaFreeHost(r.base);
            for (auto& e : r.done) if (e) { cudaEventDestroy(e); e = nullptr; }
        }
        JZ_CUDA(cudaHostAlloc(&r.base, UploadRing::kSlots * UploadRing::kSlotBytes, cudaHostAllocDefault));
        for (auto& e : r.done) JZ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& u : r.used) u = false;
        r.next = 0;
        r.device = g_ctx.device;
    }
    const int slot = r.next;
    r.next = (r.next + 1) % UploadRing::kSlots;
    if (r.used[slot]) JZ_CUDA(cudaEventSynchronize(r.done[slot]));
    void* stage = static_cast<char*>(r.base) + size_t(slot) * UploadRing::kSlotBytes;
    std::memcpy(stage, src, bytes);
    JZ_CUDA(cudaMemcpyAsync(dst, stage, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
    JZ_CUDA(cudaEventRecord(r.done[slot], as_stream(stream)));
    r.used[slot] = true;
    return JZ_OK;
}

int jz_memcpy_d2h(float* dst, const float* src, size_t count, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    if (count == 0) return JZ_OK;
    if (!dst || !src) return fail(JZ_ERR_ARG, "jz_memcpy_d2h: null pointer");
    JZ_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(float), cudaMemcpyDeviceToHost, as_stream(stream)));
    JZ_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return JZ_OK;
}

int jz_memcpy_d2d(float* dst, const float* src, size_t count, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    if (count == 0) return JZ_OK;
    if (!dst || !src) return fail(JZ_ERR_ARG, "jz_memcpy_d2d: null pointer");
    JZ_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(float), cudaMemcpyDeviceToDevice, as_stream(stream)));
    return JZ_OK;
}

}  // extern "C"
