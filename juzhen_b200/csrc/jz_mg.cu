// jz_mg.cu -- multi-GPU entry points of the C ABI (SURVEY 8b "jz_mg_*", 8e): one process per GPU, peers' buffers mapped
// through CUDA IPC (P2P loads / stores over NVLink), no torch, no NCCL, no MPI inside the library.  The caller moves
// the opaque handle bytes between its processes by whatever it has (a pipe, MPI, torch.distributed, shared memory).
// The reference has no collectives at all; this is what section 8(e) adds for C++ callers:
//
//   jz_mg_export / jz_mg_import / jz_mg_release   map a peer's buffer (cudaIpc*, offset inside the allocation kept)
//   jz_mg_barrier                                 device-side barrier over flag words in the ranks' exported memory
//   jz_mg_gemm_allgather                          column-sharded C = chain(alpha * op(A) * B[:, j0:j1]); the tcgen05
//                                                 epilogue stores every finished tile into EVERY rank's image of C
//   jz_mg_allreduce_sum                           sum of per-rank partial vectors (column sums over row-sharded data),
//                                                 every rank reads every image (P2P loads), fixed rank order
#include <cuda.h>

#include <cstring>

#include "jz_common.cuh"

namespace jz {

struct MgHandle {          // what travels between processes: fits JZ_MG_HANDLE_BYTES
    cudaIpcMemHandle_t ipc;
    unsigned long long offset;     // bytes from the base of the exported allocation
    int device;
    int magic;
};
static_assert(sizeof(MgHandle) <= JZ_MG_HANDLE_BYTES, "handle does not fit");
constexpr int kMagic = 0x4A5A4D47;   // "JZMG"

struct PeerPtrs { float* p[JZ_MAX_PEERS + 1]; };
struct FlagPtrs { unsigned* p[JZ_MAX_PEERS + 1]; };

// rank `rank` announces `epoch` in every rank's flag array, then waits until every rank has announced it here.
// One thread per rank; system-scope release / acquire so that everything the stream did before (kernel boundaries)
// is visible to the peers' later kernels.
__global__ void mg_barrier_kernel(FlagPtrs flags, int world, int rank, unsigned epoch) {
    pdl_enter();
    const int r = threadIdx.x;
    if (r < world) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags.p[r] + rank), "r"(epoch) : "memory");
        unsigned v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags.p[rank] + r) : "memory");
            if (int(v - epoch) < 0) __nanosleep(100);
        } while (int(v - epoch) < 0);   // wrap-safe: epochs only grow
        __threadfence_system();
    }
}

// out[i] = sum over ranks (in rank order: every rank computes the same bits) of image_r[i]
__global__ void __launch_bounds__(256) mg_allreduce_kernel(float* out, PeerPtrs img, int world, size_t n) {
    pdl_enter();
    for (size_t i = size_t(blockIdx.x) * 256 + threadIdx.x; i < n; i += size_t(gridDim.x) * 256) {
        float s = img.p[0][i];
        for (int r = 1; r < world; r++) s = __fadd_rn(s, img.p[r][i]);
        out[i] = s;
    }
}

}  // namespace jz

using namespace jz;

extern "C" {

int jz_mg_block_range(size_t n, int world, int rank, size_t* begin, size_t* end) {
    if (world <= 0 || rank < 0 || rank >= world || !begin || !end) return fail(JZ_ERR_ARG, "jz_mg_block_range: bad arguments");
    const size_t base = n / size_t(world), extra = n % size_t(world);
    const size_t b = size_t(rank) * base + (size_t(rank) < extra ? size_t(rank) : extra);
    *begin = b;
    *end = b + base + (size_t(rank) < extra ? 1 : 0);
    return JZ_OK;
}

int jz_mg_export(const float* dev_ptr, void* handle) {
    JZ_INIT_OR_RETURN();
    if (!dev_ptr || !handle) return fail(JZ_ERR_ARG, "jz_mg_export: null pointer");
    MgHandle h;
    std::memset(&h, 0, sizeof(h));
    CUdeviceptr base = 0;
    size_t size = 0;
    // the IPC handle names the whole allocation: keep the offset of dev_ptr inside it
    typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
    static RangeFn range_fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<RangeFn>(p);
    }();
    if (!range_fn || range_fn(&base, &size, reinterpret_cast<CUdeviceptr>(dev_ptr)) != CUDA_SUCCESS)
        return fail(JZ_ERR_CUDA, "jz_mg_export: cuMemGetAddressRange failed (not a device allocation?)");
    JZ_CUDA(cudaIpcGetMemHandle(&h.ipc, reinterpret_cast<void*>(base)));
    h.offset = reinterpret_cast<CUdeviceptr>(dev_ptr) - base;
    h.device = ctx().device;
    h.magic = kMagic;
    std::memset(handle, 0, JZ_MG_HANDLE_BYTES);
    std::memcpy(handle, &h, sizeof(h));
    return JZ_OK;
}

int jz_mg_import(const void* handle, float** peer_ptr) {
    JZ_INIT_OR_RETURN();
    if (!handle || !peer_ptr) return fail(JZ_ERR_ARG, "jz_mg_import: null pointer");
    MgHandle h;
    std::memcpy(&h, handle, sizeof(h));
    if (h.magic != kMagic) return fail(JZ_ERR_ARG, "jz_mg_import: not a jz_mg_export handle");
    void* base = nullptr;
    JZ_CUDA(cudaIpcOpenMemHandle(&base, h.ipc, cudaIpcMemLazyEnablePeerAccess));
    *peer_ptr = reinterpret_cast<float*>(static_cast<char*>(base) + h.offset);
    return JZ_OK;
}

int jz_mg_release(float* peer_ptr) {
    JZ_INIT_OR_RETURN();
    if (!peer_ptr) return JZ_OK;
    // cudaIpcCloseMemHandle wants the base the open returned: recover it from the mapping's address range
    CUdeviceptr base = 0;
    size_t size = 0;
    typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p ||
        reinterpret_cast<RangeFn>(p)(&base, &size, reinterpret_cast<CUdeviceptr>(peer_ptr)) != CUDA_SUCCESS) {
        cudaGetLastError();
        return fail(JZ_ERR_CUDA, "jz_mg_release: unknown mapping");
    }
    JZ_CUDA(cudaIpcCloseMemHandle(reinterpret_cast<void*>(base)));
    return JZ_OK;
}

int jz_mg_barrier(unsigned* const* flags, int world, int rank, unsigned epoch, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    if (world < 1 || world > JZ_MAX_PEERS + 1 || rank < 0 || rank >= world || !flags) return fail(JZ_ERR_ARG, "jz_mg_barrier: bad arguments");
    if (world == 1) return JZ_OK;
    FlagPtrs f;
    for (int r = 0; r <= JZ_MAX_PEERS; r++) f.p[r] = r < world ? flags[r] : nullptr;
    for (int r = 0; r < world; r++)
        if (!f.p[r]) return fail(JZ_ERR_ARG, "jz_mg_barrier: null flag array for rank %d", r);
    JZ_LAUNCH(mg_barrier_kernel, 1, 32, 0, as_stream(stream), f, world, rank, epoch);
    return JZ_OK;
}

int jz_mg_gemm_allgather(int transA, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda,
                         const float* B_block, size_t ldb, float* const* c_images, int world, int rank,
                         const jz_step* steps, int nsteps, int mode, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    if (world < 1 || world > JZ_MAX_PEERS + 1 || rank < 0 || rank >= world || !c_images) return fail(JZ_ERR_ARG, "jz_mg_gemm_allgather: bad arguments");
    size_t j0 = 0, j1 = 0;
    jz_mg_block_range(n, world, rank, &j0, &j1);
    float* peers[JZ_MAX_PEERS];
    int np = 0;
    for (int r = 0; r < world; r++) {
        if (!c_images[r]) return fail(JZ_ERR_ARG, "jz_mg_gemm_allgather: null image for rank %d", r);
        if (r != rank) peers[np++] = c_images[r] + j0 * m;
    }
    // ldc = m: the images are dense m x n column-major matrices, this rank owns columns [j0, j1)
    return jz_gemm_chain_bcast(transA, 0, m, j1 - j0, k, alpha, A, lda, B_block, ldb, c_images[rank] + j0 * m, m, peers, np, steps,
                               nsteps, mode, stream);
}

int jz_mg_allreduce_sum(float* out, float* const* partial_images, size_t n, int world, int rank, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    if (world < 1 || world > JZ_MAX_PEERS + 1 || rank < 0 || rank >= world || !partial_images || !out) return fail(JZ_ERR_ARG, "jz_mg_allreduce_sum: bad arguments");
    if (n == 0) return JZ_OK;
    PeerPtrs img;
    for (int r = 0; r <= JZ_MAX_PEERS; r++) img.p[r] = r < world ? partial_images[r] : nullptr;
    for (int r = 0; r < world; r++)
        if (!img.p[r]) return fail(JZ_ERR_ARG, "jz_mg_allreduce_sum: null image for rank %d", r);
    const size_t cap = size_t(ctx().sm_count) * 4, blocks = ceil_div(n, size_t(256));
    JZ_LAUNCH(mg_allreduce_kernel, unsigned(blocks < cap ? blocks : cap), 256, 0, as_stream(stream), out, img, world, n);
    return JZ_OK;
}

}  // extern "C"
