// jz_gemm.cu -- the GEMM behind Matrix<CUDAfloat>::dot / operator* (SURVEY 8a row a17):
//     C(m x n, ldc) = alpha * op(A)(m x k) * op(B)(k x n) + beta * C        (column-major)
// replacing cublasSgemm (cpp/cumatrix.cu:177-197).  No cuBLAS anywhere.
//
// Main path (sm_100a): TMA -> shared memory (128B swizzle) -> tcgen05.mma.kind::tf32 with the
// fp32 accumulator in TMEM -> tcgen05.ld -> coalesced column-major stores.
//   * warp-specialised CTA: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc),
//     warps 2..5 = epilogue (one TMEM lane quarter each);
//   * CG = 2: a CTA pair (cluster 2x1x1, cta_group::2) computes one 256 x 256 tile; each CTA
//     loads its 128 rows of A and its 128-column half of B, the leader issues M=256 N=256 K=8
//     MMAs that read both CTAs' shared memory.  CG = 1 is the single-SM 128 x 256 variant;
//   * operands are read by TMA straight from the caller's column-major storage, in either
//     major: column-major B and flagged-transpose A are K-major (128B swizzle); plain A and
//     flagged-transpose B are MN-major (32 contiguous rows x 32 k boxes, the 128B swizzle with
//     32-byte atoms that tf32 MN-major operands require) -- no re-layout pass, no workspace;
//   * 3xTF32 (default, fp32 accuracy): every k-block issues A_lo*B_hi + A_hi*B_lo + A_hi*B_hi
//     into the same TMEM accumulator.  kind::tf32 reads fp32 words and drops the low 13
//     mantissa bits, so the raw tile already IS the "hi" operand (hi = trunc_tf32(x)); the
//     eight epilogue warps, idle between accumulator drains, compute
//     lo = rna_tf32(x - trunc_tf32(x)) from each landed tile into a second shared-memory
//     buffer while the tensor core works on the previous stage (MODE_XFORM).  The older pre-split variant (MODE_PRESPLIT: a pre-pass writes K-major
//     hi/lo images to workspace, 4 TMA loads per stage) is kept for operands TMA cannot
//     address in place and for A/B measurement (JZ_GEMM_PRESPLIT=1);
//
// Other paths: shapes the tensor path does not take (m or n < 64, k < 32, tiny, or operands it cannot
// address) go to the latency-oriented warp-per-tile fp32 kernel of jz_gemm_small.cu when the product
// is at most 2^26 multiply-adds (a training step at batch 32), else to a bounds-checked
// shared-memory fp32 FMA (SIMT) kernel, which also serves mode JZ_GEMM_FP32_SIMT; rank-1 products (k == 1: the reference's broadcast idiom) are a streaming
// outer-product kernel.
#include <cuda.h>

#include <cstring>
#include <mutex>

#include "jz_common.cuh"
#include "jz_math.cuh"

namespace jz {

// jz_gemm_small.cu
bool gemm_small_wants(size_t m, size_t n, size_t k);
int launch_gemm_small(int ta, int tb, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda,
                      const float* B, size_t ldb, float beta, float* C, size_t ldc, const ChainParams& chain,
                      cudaStream_t s, size_t batch = 1, size_t strideA = 0, size_t strideB = 0, size_t strideC = 0);

// ======================================================================= SIMT fallback
constexpr int SBM = 64, SBN = 64, SBK = 16;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) sgemm_simt_kernel(size_t m, size_t n, size_t k, float alpha, const float* A,
                                                         size_t lda, const float* B, size_t ldb, float beta, float* C,
                                                         size_t ldc, ChainParams chain_p) {
    __shared__ ChainParams chain;
    stage_chain(&chain, chain_p, threadIdx.x);
    __syncthreads();
    __shared__ __align__(16) float As[SBK][SBM + 4];  // read back as float4
    __shared__ __align__(16) float Bs[SBK][SBN + 4];
    const int t = threadIdx.x;
    const size_t i0 = size_t(blockIdx.x) * SBM, j0 = size_t(blockIdx.y) * SBN;
    const int tx = t & 15, ty = t >> 4;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[r][c] = 0.0f;

    for (size_t k0 = 0; k0 < k; k0 += SBK) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            // A tile: op(A)(i, kk)
            int ii, kk;
            if (!TA) { ii = t & 63; kk = (t >> 6) + 4 * r; }
            else { kk = t & 15; ii = (t >> 4) + 16 * r; }
            const size_t gi = i0 + ii, gk = k0 + kk;
            float v = 0.0f;
            if (gi < m && gk < k) v = TA ? A[gi * lda + gk] : A[gk * lda + gi];
            As[kk][ii] = v;
            // B tile: op(B)(kk, j)
            int jj, kb;
            if (!TB) { kb = t & 15; jj = (t >> 4) + 16 * r; }
            else { jj = t & 63; kb = (t >> 6) + 4 * r; }
            const size_t gj = j0 + jj, gk2 = k0 + kb;
            float w = 0.0f;
            if (gj < n && gk2 < k) w = TB ? B[gk2 * ldb + gj] : B[gj * ldb + gk2];
            Bs[kb][jj] = w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SBK; kk++) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][tx * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][ty * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const size_t gj = j0 + ty * 4 + c;
        if (gj >= n) continue;
        float v[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const size_t gi = i0 + tx * 4 + r;
            float x = alpha * acc[r][c];
            if (beta != 0.0f && gi < m) x += beta * C[gj * ldc + gi];
            v[r] = x;
        }
        if (chain.n) apply_chain<4>(v, chain);
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const size_t gi = i0 + tx * 4 + r;
            if (gi < m) C[gj * ldc + gi] = v[r];
        }
    }
}

static int launch_simt(int ta, int tb, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda,
                       const float* B, size_t ldb, float beta, float* C, size_t ldc, const ChainParams& chain,
                       cudaStream_t s) {
    const size_t gx = ceil_div(m, SBM), gy = ceil_div(n, SBN);
    if (gy > 65535) return fail(JZ_ERR_UNSUPPORTED, "simt gemm: n too large for the fallback kernel");
    const dim3 grid((unsigned)gx, (unsigned)gy, 1);
    if (!ta && !tb) JZ_LAUNCH((sgemm_simt_kernel<false, false>), grid, 256, 0, s, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain);
    else if (ta && !tb) JZ_LAUNCH((sgemm_simt_kernel<true, false>), grid, 256, 0, s, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain);
    else if (!ta && tb) JZ_LAUNCH((sgemm_simt_kernel<false, true>), grid, 256, 0, s, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain);
    else JZ_LAUNCH((sgemm_simt_kernel<true, true>), grid, 256, 0, s, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain);
    ctx().gemm_last_path = 2;
    return JZ_OK;
}

// rank-1: C = alpha*u*v^T (+beta*C) ; u = op(A)(:,0) (stride su), v = op(B)(0,:) (stride sv).
// u == nullptr means the product term is absent (k == 0).
__global__ void __launch_bounds__(256) rank1_kernel(size_t m, size_t n, float alpha, const float* u, size_t su,
                                                    const float* v, size_t sv, float beta, float* C, size_t ldc,
                                                    ChainParams chain_p) {
    __shared__ ChainParams chain;
    stage_chain(&chain, chain_p, threadIdx.x);
    __syncthreads();
    for (size_t j = blockIdx.y; j < n; j += gridDim.y) {
        const float vj = u ? alpha * v[j * sv] : 0.0f;
        for (size_t i = size_t(blockIdx.x) * 256 + threadIdx.x; i < m; i += size_t(gridDim.x) * 256) {
            float x[1] = {u ? u[i * su] * vj : 0.0f};
            if (beta != 0.0f) x[0] += beta * C[j * ldc + i];
            if (chain.n) apply_chain<1>(x, chain);
            C[j * ldc + i] = x[0];
        }
    }
}

// ======================================================================= tcgen05 path
namespace tc {

constexpr int BK = 32;               // fp32 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 8;            // tf32: 32 bytes per MMA k-step
constexpr int TILE_M = 128;          // rows of A per CTA (TMEM lanes)
// accumulator columns per tile: TN = 256 (default) or 128 (chosen when 256-wide tiles leave a large partial wave)
constexpr int A_BYTES = TILE_M * BK * 4;  // 16 KB
constexpr int NUM_EPI_WARPS = 8;
constexpr int FIRST_EPI_WARP = 2;
constexpr int NUM_THREADS = 32 * (FIRST_EPI_WARP + NUM_EPI_WARPS);  // TMA warp + MMA warp + 8 epilogue/transform warps
constexpr int MODE_TF32 = 0;      // single pass over the raw fp32 tiles
constexpr int MODE_PRESPLIT = 1;  // 3xTF32, hi/lo images written by a pre-pass (K-major only)
constexpr int MODE_XFORM = 2;     // 3xTF32, lo computed in shared memory by the transform warps
constexpr int MN_BOX_BYTES = 32 * BK * 4;  // one MN-major TMA box: 32 contiguous rows x 32 k = 4 KB

template <int CG, int TN> __host__ __device__ constexpr int b_rows() { return TN / CG; }
template <int CG, int TN> __host__ __device__ constexpr int b_bytes() { return b_rows<CG, TN>() * BK * 4; }
template <int CG, int MODE, int TN> __host__ __device__ constexpr int stage_bytes() { return (MODE != MODE_TF32 ? 2 : 1) * (A_BYTES + b_bytes<CG, TN>()); }
template <int CG, int MODE, int TN> __host__ __device__ constexpr int num_stages() {
    constexpr int s = (227 * 1024 - 2048) / stage_bytes<CG, MODE, TN>();
    return s > 8 ? 8 : s;
}
template <int CG, int MODE, int TN> __host__ __device__ constexpr int smem_bytes() {
    return num_stages<CG, MODE, TN>() * stage_bytes<CG, MODE, TN>() + 1024 /*align slack*/ + 256 /*barriers*/;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// TO_LEADER: the copy (issued by either CTA of a pair) signals the LEADER CTA's barrier (peer bit cleared);
// otherwise it signals the issuing CTA's own barrier.
template <bool TO_LEADER>
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    if constexpr (TO_LEADER) {
        asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
            : "memory");
    } else {
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
            : "memory");
    }
}
// One operand tile of `ROWS` rows x BK k starting at (row0, kc).
//   K-major source: one box {BK, ROWS}: ROWS rows of 128 B, 128B swizzle.
//   MN-major source: ROWS/32 boxes {32 rows, BK}: BK rows of 128 B each holding 32 consecutive operand rows
//   (4 KB per box, 128B swizzle with 32-byte atoms).
template <bool MN, int ROWS, bool TO_LEADER>
__device__ __forceinline__ void tma_load_tile(uint32_t dst, const CUtensorMap* map, uint32_t bar, int kc, int row0) {
    if constexpr (MN) {
#pragma unroll
        for (int i = 0; i < ROWS / 32; i++) tma_load_2d<TO_LEADER>(dst + i * MN_BOX_BYTES, map, bar, row0 + 32 * i, kc);
    } else {
        tma_load_2d<TO_LEADER>(dst, map, bar, kc, row0);
    }
}
template <int CG>
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    if constexpr (CG == 2) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc),
            "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc),
            "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// tcgen05.commit: arrive on `bar` (in every CTA of the pair for CG == 2) when all prior MMAs retire
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    if constexpr (CG == 2) {
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
            "h"((uint16_t)3)
            : "memory");
    } else {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    if constexpr (CG == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    if constexpr (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory operand descriptors (sm_100 format: version 1 at bit 46).
//   K-major tile: rows of 128 B, 128B swizzle (layout type 2), 8-row groups 1024 B apart (SBO); one UMMA_K step
//   (8 tf32 = 32 B) advances the start address by 32 B inside the swizzled row.
//   MN-major tile: the tf32-only canonical layout "128B swizzle, 32B atoms" (layout type 1): an atom is 32
//   consecutive operand rows (128 B) x 4 k; atoms of the next 4 k follow 512 B later (SBO), the next 32 rows
//   start one TMA box = 4096 B later (LBO); one UMMA_K step covers two k-atoms = 1024 B.
template <bool MN>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= uint64_t((saddr & 0x3FFFFu) >> 4);   // start address, 16-byte units
    if constexpr (MN) {
        d |= uint64_t(MN_BOX_BYTES >> 4) << 16;   // leading byte offset: between 32-row atoms
        d |= uint64_t(512 >> 4) << 32;            // stride byte offset: between 4-k atoms
        d |= uint64_t(1) << 46;
        d |= uint64_t(1) << 61;                   // SWIZZLE_128B_BASE32B
    } else {
        d |= uint64_t(0) << 16;                   // leading byte offset: unused for swizzled K-major
        d |= uint64_t(1024 >> 4) << 32;           // stride byte offset between 8-row groups
        d |= uint64_t(1) << 46;
        d |= uint64_t(2) << 61;                   // SWIZZLE_128B
    }
    return d;
}
template <bool MN> __host__ __device__ constexpr uint32_t kstep_bytes() { return MN ? 1024u : 32u; }
// instruction descriptor: tf32 x tf32 -> f32; bit 15 / 16 = A / B is MN-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn ? 1u << 15 : 0u) | (b_mn ? 1u << 16 : 0u) |
           (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

struct GemmArgs {
    size_t m, n, k;
    float alpha, beta;
    float* C;
    size_t ldc;
    unsigned tiles_m, tiles_n;  // in units of (CG*128) x 256 tiles
    int kb_per_chunk;           // k-blocks accumulated inside TMEM before promotion to registers
    int n_peers;                // additional destinations (peer-GPU images of C, same ldc)
    float* peers[JZ_MAX_PEERS];
    ChainParams chain;
};

__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// lo part of the 3xTF32 split against the hardware's own hi: kind::tf32 drops the low 13 mantissa bits of the
// fp32 word it reads, so hi = x & 0xFFFFE000 and x - hi is exact in fp32; lo is that remainder rounded to tf32
// (nearest, ties away: add half an ulp to the magnitude, truncate), so the tensor core reads it unchanged.
// inf/nan keep their semantics through hi alone (lo = 0 avoids inf - inf).
__device__ __forceinline__ float tf32_lo_of(float x) {
    const uint32_t b = __float_as_uint(x);
    const float d = __fsub_rn(x, __uint_as_float(b & 0xFFFFE000u));
    const uint32_t r = (__float_as_uint(d) + 0x1000u) & 0xFFFFE000u;
    return d == d ? __uint_as_float(r) : 0.0f;   // x = inf/nan gives d = nan
}

// One (CG*128) x TN output tile per CTA group.  tmA*/tmB*: tensor maps of the operands (see make_map).
//
// Accumulation is two-level: the tensor core adds each MMA's partial product into the TMEM
// accumulator with truncation (measured: relative error 6.9e-9 * k, i.e. biased, linear in the
// length of the chain), so only `kb_per_chunk` k-blocks are chained inside TMEM; the epilogue
// warps then promote the chunk into fp32 REGISTER accumulators with round-to-nearest adds while
// the MMA warp is already filling the other TMEM buffer (2 x TN columns).
//
// Barriers (per smem stage): MODE_TF32 / MODE_PRESPLIT: TMA of both CTAs -> full (leader) -> MMA -> empty (both).
// MODE_XFORM: TMA -> full (own CTA) -> epilogue warps write lo -> ready (leader) -> MMA -> empty (both).
template <int CG, int MODE, int TN, bool AMN, bool BMN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                    const GemmArgs args) {
    static_assert(MODE != MODE_PRESPLIT || (!AMN && !BMN), "pre-split images are K-major");
    constexpr bool SPLIT = MODE != MODE_TF32;
    constexpr bool XFORM = MODE == MODE_XFORM;
    constexpr bool TO_LEADER = CG == 2 && !XFORM;   // whose `full` barrier the TMA copies signal
    constexpr int TILE_N = TN;
    constexpr int HALF_N = TN / 2;       // columns drained by one epilogue warp
    constexpr int STAGES = num_stages<CG, MODE, TN>();
    constexpr int STAGE_BYTES = stage_bytes<CG, MODE, TN>();
    constexpr int B_ROWS = b_rows<CG, TN>();
    constexpr int B_BYTES = b_bytes<CG, TN>();
    constexpr int RAW_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t IDESC = make_idesc_tf32(CG * 128, TILE_N, AMN, BMN);
    constexpr uint32_t KA = kstep_bytes<AMN>(), KB = kstep_bytes<BMN>();

    extern __shared__ uint8_t smem_raw[];
    __shared__ ChainParams s_chain;
    __shared__ float* s_peers[JZ_MAX_PEERS];
    stage_chain(&s_chain, args.chain, threadIdx.x);  // visible after the setup barrier below
    if (threadIdx.x == 32) {  // static indices: direct constant-bank reads (a runtime index would spill the array)
#pragma unroll
        for (int q = 0; q < JZ_MAX_PEERS; q++) s_peers[q] = args.peers[q];
    }
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
    // barriers: full[STAGES], empty[STAGES], ready[STAGES], tmem_full[2], tmem_empty[2], then the TMEM base slot
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto ready_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tmem_full_bar = [&](int b) { return bar_base + 8u * (3 * STAGES + b); };
    auto tmem_empty_bar = [&](int b) { return bar_base + 8u * (3 * STAGES + 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (3 * STAGES + 4);
    uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(gen_base + STAGES * STAGE_BYTES + 8 * (3 * STAGES + 4));

    const int warp = threadIdx.x >> 5;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;

    // tile coordinates (grouped rasterisation for L2 reuse of the A / B panels)
    const unsigned tile = CG == 2 ? blockIdx.x >> 1 : blockIdx.x;
    constexpr unsigned GROUP = 8;
    const unsigned per_group = GROUP * args.tiles_n;
    const unsigned group_id = tile / per_group;
    const unsigned first_m = group_id * GROUP;
    const unsigned group_m = args.tiles_m - first_m < GROUP ? args.tiles_m - first_m : GROUP;
    const unsigned tm = first_m + (tile % per_group) % group_m;
    const unsigned tn = (tile % per_group) / group_m;
    const int m0 = int(tm) * (CG * TILE_M) + int(rank) * TILE_M;  // first row of this CTA
    const int n0 = int(tn) * TILE_N;
    const int nb0 = n0 + (CG == 2 ? int(rank) * (TILE_N / 2) : 0);         // first B row (= C column) this CTA loads
    const int num_kb = int((args.k + BK - 1) / BK);
    const int kbc = args.kb_per_chunk;
    const int num_chunks = (num_kb + kbc - 1) / kbc;

    if (warp == 0 && elect_one()) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_hi) : "memory");
        if (MODE == MODE_PRESPLIT) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_lo) : "memory");
        }
    }
    if (warp == 1) {
        if (elect_one()) {
            for (int s = 0; s < STAGES; s++) {
                mbar_init(full_bar(s), 1);
                mbar_init(empty_bar(s), 1);
                mbar_init(ready_bar(s), NUM_EPI_WARPS * CG);       // every transform (= epilogue) warp of the pair
            }
            for (int b = 0; b < 2; b++) {
                mbar_init(tmem_full_bar(b), 1);
                mbar_init(tmem_empty_bar(b), NUM_EPI_WARPS * CG);  // every epilogue warp of the pair
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<CG>(tmem_slot, 2 * TILE_N);
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (int kb = 0; kb < num_kb; kb++) {
                mbar_wait(empty_bar(stage), phase ^ 1);
                if (XFORM) mbar_arrive_expect_tx(full_bar(stage), uint32_t(RAW_BYTES));
                else if (leader) mbar_arrive_expect_tx(full_bar(stage), uint32_t(STAGE_BYTES) * CG);
                const uint32_t sa = smem_base + stage * STAGE_BYTES;
                const int kc = kb * BK;
                tma_load_tile<AMN, TILE_M, TO_LEADER>(sa, &tmA_hi, full_bar(stage), kc, m0);
                tma_load_tile<BMN, B_ROWS, TO_LEADER>(sa + A_BYTES, &tmB_hi, full_bar(stage), kc, nb0);
                if (MODE == MODE_PRESPLIT) {
                    tma_load_tile<false, TILE_M, TO_LEADER>(sa + RAW_BYTES, &tmA_lo, full_bar(stage), kc, m0);
                    tma_load_tile<false, B_ROWS, TO_LEADER>(sa + RAW_BYTES + A_BYTES, &tmB_lo, full_bar(stage), kc, nb0);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            uint32_t stage = 0, phase = 0;
            int chunk = 0, in_chunk = 0;
            for (int kb = 0; kb < num_kb; kb++) {
                const uint32_t buf = uint32_t(chunk) & 1u;
                if (in_chunk == 0) {  // this TMEM buffer must have been drained by every epilogue warp
                    mbar_wait(tmem_empty_bar(buf), ((uint32_t(chunk) >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                }
                mbar_wait(XFORM ? ready_bar(stage) : full_bar(stage), phase);
                tc_fence_after();
                const bool chunk_end = (in_chunk == kbc - 1) || (kb == num_kb - 1);
                if (elect_one()) {
                    const uint32_t d = tmem_base + buf * TILE_N;
                    const uint32_t sa = smem_base + stage * STAGE_BYTES;
                    const uint32_t a_hi = sa, b_hi = sa + A_BYTES;
                    const uint32_t a_lo = sa + RAW_BYTES, b_lo = sa + RAW_BYTES + A_BYTES;
                    uint32_t acc = in_chunk == 0 ? 0u : 1u;
                    if (SPLIT) {
#pragma unroll
                        for (int ks = 0; ks < BK / UMMA_K; ks++) {
                            umma_tf32<CG>(d, make_smem_desc<AMN>(a_lo + ks * KA), make_smem_desc<BMN>(b_hi + ks * KB), IDESC, acc);
                            acc = 1u;
                        }
#pragma unroll
                        for (int ks = 0; ks < BK / UMMA_K; ks++)
                            umma_tf32<CG>(d, make_smem_desc<AMN>(a_hi + ks * KA), make_smem_desc<BMN>(b_lo + ks * KB), IDESC, 1u);
                    }
#pragma unroll
                    for (int ks = 0; ks < BK / UMMA_K; ks++) {
                        umma_tf32<CG>(d, make_smem_desc<AMN>(a_hi + ks * KA), make_smem_desc<BMN>(b_hi + ks * KB), IDESC, acc);
                        acc = 1u;
                    }
                    umma_commit<CG>(empty_bar(stage));                   // frees this smem stage (both CTAs)
                    if (chunk_end) umma_commit<CG>(tmem_full_bar(buf));  // chunk accumulator complete
                }
                __syncwarp();
                if (chunk_end) { chunk++; in_chunk = 0; } else { in_chunk++; }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue warps: lo-part transform of landed stages (MODE_XFORM), =====================
        // ===================== TMEM chunks -> fp32 registers (RN) -> global                     =====================
        const int e = warp - FIRST_EPI_WARP;
        const int quarter = warp & 3;   // TMEM lane quarter this warp may access (hardware: warp id % 4)
        const int half = e >> 2;        // which half of the accumulator columns
        const int lane = threadIdx.x & 31;
        float acc[HALF_N];
#pragma unroll
        for (int i = 0; i < HALF_N; i++) acc[i] = 0.0f;
        const uint32_t lane_addr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(half * HALF_N);
        const uint32_t empty0 = tmem_empty_bar(0) & 0xFEFFFFFFu, empty1 = tmem_empty_bar(1) & 0xFEFFFFFFu;
        const uint32_t ready0 = ready_bar(0) & 0xFEFFFFFFu;  // on the leader CTA
        // Transform and drain interleave in ONE instruction stream per warp, ordered so that neither can starve
        // the other: k-block j reuses the smem stage of k-block j - STAGES, so it cannot land before the MMAs of
        // k-block j - STAGES have retired; chunk c (k-blocks c*kbc .. (c+1)*kbc - 1) is therefore complete by the
        // time k-block (c+1)*kbc + STAGES - 1 lands, and is drained right before that k-block is transformed --
        // after every k-block the chunk itself (and the next chunk's first STAGES - 1) has been handed to the MMA.
        constexpr int XT = 32 * NUM_EPI_WARPS;            // transform threads per CTA
        constexpr int N4 = RAW_BYTES / 16;                // float4 words per stage (A tile then B tile, contiguous)
        constexpr int PER = N4 / XT;                      // float4 words per thread per stage
        constexpr int BATCH = PER <= 8 ? PER : (PER % 4 == 0 ? 4 : 3);   // all of a thread's loads in flight together
        static_assert(N4 % XT == 0 && PER % BATCH == 0, "stage size must divide evenly among the transform threads");
        const int te = threadIdx.x - 32 * FIRST_EPI_WARP;
        uint32_t stage = 0, phase = 0;
        int next_drain = 0;
        const int kb_end = XFORM ? num_kb : 0;
        for (int kb = 0; kb <= kb_end; kb++) {
            while (next_drain < num_chunks && (kb >= kb_end || (next_drain + 1) * kbc + STAGES - 1 <= kb)) {
                const int chunk = next_drain++;
                const uint32_t buf = uint32_t(chunk) & 1u;
                mbar_wait(tmem_full_bar(buf), (uint32_t(chunk) >> 1) & 1u);
                tc_fence_after();
#pragma unroll
                for (int p = 0; p < HALF_N / 32; p++) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(lane_addr + buf * TILE_N + p * 32, r);
#pragma unroll
                    for (int c = 0; c < 32; c++) acc[p * 32 + c] = __fadd_rn(acc[p * 32 + c], __uint_as_float(r[c]));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(buf ? empty1 : empty0);  // on the leader CTA's barrier
            }
            if (XFORM && kb < kb_end) {
                mbar_wait(full_bar(stage), phase);
                const float4* src = reinterpret_cast<const float4*>(gen_base + stage * STAGE_BYTES) + te;
                float4* dst = reinterpret_cast<float4*>(gen_base + stage * STAGE_BYTES + RAW_BYTES) + te;
#pragma unroll
                for (int i0 = 0; i0 < PER; i0 += BATCH) {
                    float4 v[BATCH];
#pragma unroll
                    for (int u = 0; u < BATCH; u++) v[u] = src[(i0 + u) * XT];
#pragma unroll
                    for (int u = 0; u < BATCH; u++)
                        dst[(i0 + u) * XT] = make_float4(tf32_lo_of(v[u].x), tf32_lo_of(v[u].y), tf32_lo_of(v[u].z), tf32_lo_of(v[u].w));
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(ready0 + 8u * stage);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        const size_t row = size_t(m0) + quarter * 32 + lane;
        const bool row_ok = row < args.m;
        const size_t ncol0 = size_t(n0) + half * HALF_N;
        const size_t ldc = args.ldc;
#pragma unroll
        for (int p = 0; p < HALF_N / 32; p++) {
            const size_t colp = ncol0 + p * 32;
            if (colp < args.n) {  // warp-uniform
                const int ncols = args.n - colp < 32 ? int(args.n - colp) : 32;
                float v[32];
#pragma unroll
                for (int c = 0; c < 32; c++) v[c] = args.alpha * acc[p * 32 + c];
                if (args.beta != 0.0f && row_ok) {
                    const float* src = args.C + row + colp * ldc;  // running pointer: no 32 hoisted addresses
#pragma unroll
                    for (int c = 0; c < 32; c++) {
                        if (c < ncols) v[c] += args.beta * *src;
                        src += ldc;
                    }
                }
                if (s_chain.n) apply_chain<32>(v, s_chain);
                // a warp writes 32 consecutive floats (128 B) per column; destination 0 is the local C, the rest
                // are the peer GPUs' images of C (fused all-gather: P2P stores over NVLink, tile by tile while
                // other tiles are still computing)
                if (row_ok) {
                    for (int d = 0; d <= args.n_peers; d++) {
                        float* dst = (d == 0 ? args.C : s_peers[d - 1]) + row + colp * ldc;
#pragma unroll
                        for (int c = 0; c < 32; c++) {
                            if (c < ncols) *dst = v[c];
                            dst += ldc;
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc<CG>(tmem_base, 2 * TILE_N);
}

// ----------------------------------------------------------------------- operand pre-pass
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = tf32_rna(x);
    lo = tf32_rna(__fsub_rn(x, hi));  // x - hi is exact in fp32
    if (!isfinite(hi)) { hi = x; lo = 0.0f; }  // keep inf/nan semantics, avoid inf - inf
}

// source already K-major: src(r, kk) at r*ld + kk ; dst[r*kp + kk].
// One thread per quad of k: a 128-bit load, the split, one or two 128-bit stores (VEC), grid-stride.
template <bool SPLIT, bool VEC>
__global__ void __launch_bounds__(256) prep_kmajor_kernel(float* hi, float* lo, size_t kp, const float* src, size_t ld,
                                                          size_t rows, size_t k) {
    const size_t kq = kp >> 2;  // quads per row (kp is a multiple of 4)
    const size_t total = rows * kq;
    for (size_t idx = size_t(blockIdx.x) * 256 + threadIdx.x; idx < total; idx += size_t(gridDim.x) * 256) {
        const size_t r = idx / kq, kk = (idx - r * kq) << 2;
        float x[4];
        if (VEC && kk + 4 <= k) {
            const float4 v = *reinterpret_cast<const float4*>(src + r * ld + kk);
            x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) x[i] = kk + i < k ? src[r * ld + kk + i] : 0.0f;
        }
        float h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (SPLIT) split_tf32(x[i], h[i], l[i]);
            else h[i] = x[i];
        }
        *reinterpret_cast<float4*>(hi + r * kp + kk) = make_float4(h[0], h[1], h[2], h[3]);
        if (SPLIT) *reinterpret_cast<float4*>(lo + r * kp + kk) = make_float4(l[0], l[1], l[2], l[3]);
    }
}

// source MN-major: src(r, kk) at kk*ld + r ; dst[r*kp + kk]   (tiled transpose).
// 64 x 64 tiles through shared memory; interior tiles use 128-bit accesses on both sides when VEC.
template <bool SPLIT, bool VEC>
__global__ void __launch_bounds__(256) prep_transpose_kernel(float* hi, float* lo, size_t kp, const float* src,
                                                             size_t ld, size_t rows, size_t k, size_t tiles_r,
                                                             size_t tiles_k) {
    __shared__ float tile[64][65];  // tile[kk][r]
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const size_t ntiles = tiles_r * tiles_k;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const size_t tr = t % tiles_r, tk = t / tiles_r;
        const size_t r0 = tr * 64, k0 = tk * 64;
        const bool full = VEC && r0 + 64 <= rows && k0 + 64 <= k;
        if (full) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int kl = ty + 16 * q;
                const float4 v = *reinterpret_cast<const float4*>(src + (k0 + kl) * ld + r0 + 4 * tx);
                tile[kl][4 * tx + 0] = v.x; tile[kl][4 * tx + 1] = v.y;
                tile[kl][4 * tx + 2] = v.z; tile[kl][4 * tx + 3] = v.w;
            }
        } else {
            for (int e = threadIdx.x; e < 64 * 64; e += 256) {
                const int kl = e >> 6, rl = e & 63;
                tile[kl][rl] = (k0 + kl < k && r0 + rl < rows) ? src[(k0 + kl) * ld + r0 + rl] : 0.0f;
            }
        }
        __syncthreads();
        // write: row r of the image, 4 consecutive k per thread (k0 and kp are multiples of 4: always in bounds)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int rl = ty + 16 * q;
            const size_t r = r0 + rl, kk = k0 + 4 * tx;
            if (r < rows && kk < kp) {
                float h[4], l[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float x = tile[4 * tx + i][rl];
                    if (SPLIT) split_tf32(x, h[i], l[i]);
                    else h[i] = x;
                }
                *reinterpret_cast<float4*>(hi + r * kp + kk) = make_float4(h[0], h[1], h[2], h[3]);
                if (SPLIT) *reinterpret_cast<float4*>(lo + r * kp + kk) = make_float4(l[0], l[1], l[2], l[3]);
            }
        }
        __syncthreads();
    }
}

// ----------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    });
    return fn;
}

// Tensor map of one operand over its source storage, zero OOB fill.
//   K-major ([rows][k], row stride `stride_elems`): dims {k, rows}, box {32, box_rows}, 128B swizzle.
//   MN-major ([k][rows], k stride `stride_elems`): dims {rows, k}, box {32 rows, 32 k}, 128B swizzle / 32B atoms.
static int make_map(CUtensorMap* map, const float* base, size_t rows, size_t k, size_t stride_elems, int box_rows,
                    bool mn_major) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(JZ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {cuuint64_t(mn_major ? rows : k), cuuint64_t(mn_major ? k : rows)};
    cuuint64_t strides[1] = {cuuint64_t(stride_elems) * sizeof(float)};
    cuuint32_t box[2] = {cuuint32_t(BK), cuuint32_t(mn_major ? BK : box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(JZ_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
    return JZ_OK;
}

static int g_cg = 0;  // 0 = auto (2), else forced via JZ_GEMM_CG
static int pick_cg() {
    if (g_cg) return g_cg;
    const char* e = std::getenv("JZ_GEMM_CG");
    g_cg = (e && e[0] == '1') ? 1 : 2;
    return g_cg;
}
static bool env_flag(const char* name) {
    const char* e = std::getenv(name);
    return e && e[0] && e[0] != '0';
}

// Tile width.  A CTA pair owns a 256 x TN tile.  256-wide tiles are preferred: per k-block the A tile is loaded
// (and, in 3xTF32, transformed) once per tile whatever its width, so a 128-wide tile costs well over half a
// 256-wide one (measured, profiles/r01o_gemm_sweep.log: 2048^2 3xTF32 66 us with 64 wide tiles in one wave against
// 97 us with 128 narrow tiles in two; 4096^2 0.52 against 0.62 ms although the narrow tiling has the fuller last
// wave).  Narrow tiles only pay when the wide tiling cannot occupy half the SM pairs (1024^2: 16 tiles on 74 pair
// slots).  JZ_GEMM_TN=128|256 forces it.
static int pick_tile_n(size_t m, size_t n, int cg) {
    static const int forced = [] {
        const char* e = std::getenv("JZ_GEMM_TN");
        return e ? std::atoi(e) : 0;
    }();
    if (cg != 2) return 256;
    if (forced == 128 || forced == 256) return forced;
    const size_t slots = size_t(ctx().sm_count) / 2;
    const size_t t256 = ceil_div(m, size_t(256)) * ceil_div(n, size_t(256));
    return t256 * 2 <= slots ? 128 : 256;
}

// k-blocks (of 32) chained inside TMEM before RN promotion: 3xTF32 keeps the chain short for
// fp32-grade accuracy; TF32 mode is input-rounding dominated so long chains are harmless.
static int chunk_kb(bool split) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = std::getenv("JZ_GEMM_CHUNK_KB");
        forced = e ? std::atoi(e) : 0;
    }
    if (forced > 0) return forced;
    return split ? 4 : 32;
}

struct Operand {
    const float* hi = nullptr;   // raw fp32 (MODE_TF32 / MODE_XFORM) or the tf32 hi image (MODE_PRESPLIT)
    const float* lo = nullptr;   // MODE_PRESPLIT only
    size_t stride = 0;           // elements between consecutive rows (K-major) / consecutive k (MN-major)
    bool mn = false;             // MN-major: element (r, kk) at kk*stride + r
    void* owned = nullptr;       // workspace to release
};

// Describe an operand for the kernel.  kmajor_src: src(r,kk) at r*ld+kk, else at kk*ld+r.
// In place (no copy, either major) whenever TMA can address the source: 16-byte aligned base and row pitch.
// Otherwise -- or when `presplit` asks for hi/lo images -- a pre-pass writes K-major image(s) to workspace.
static int prepare_operand(Operand& op, const float* src, size_t ld, bool kmajor_src, size_t rows, size_t k,
                           bool presplit, cudaStream_t s) {
    static const bool no_mn = env_flag("JZ_GEMM_NO_MN");   // debugging: never consume MN-major sources in place
    const bool tma_ok = ld % 4 == 0 && aligned16(src);
    if (!presplit && tma_ok && (kmajor_src || !no_mn)) {
        op.hi = src;
        op.stride = ld;
        op.mn = !kmajor_src;
        return JZ_OK;
    }
    const bool split = presplit;
    const size_t kp = (k + 3) & ~size_t(3);
    const size_t elems = rows * kp;
    int rc = ws_alloc(&op.owned, (split ? 2 : 1) * elems * sizeof(float), s);
    if (rc != JZ_OK) return rc;
    float* hi = static_cast<float*>(op.owned);
    float* lo = split ? hi + elems : nullptr;
    op.hi = hi;
    op.lo = lo;
    op.stride = kp;
    op.mn = false;
    const size_t cap = size_t(ctx().sm_count) * 8;
    const bool vec = tma_ok;
    if (kmajor_src) {
        const size_t blocks = ceil_div(rows * (kp >> 2), size_t(256));
        const unsigned grid = unsigned(blocks < cap * 2 ? (blocks ? blocks : 1) : cap * 2);
        if (split && vec) JZ_LAUNCH((prep_kmajor_kernel<true, true>), grid, 256, 0, s, hi, lo, kp, src, ld, rows, k);
        else if (split) JZ_LAUNCH((prep_kmajor_kernel<true, false>), grid, 256, 0, s, hi, lo, kp, src, ld, rows, k);
        else if (vec) JZ_LAUNCH((prep_kmajor_kernel<false, true>), grid, 256, 0, s, hi, lo, kp, src, ld, rows, k);
        else JZ_LAUNCH((prep_kmajor_kernel<false, false>), grid, 256, 0, s, hi, lo, kp, src, ld, rows, k);
    } else {
        const size_t tiles_r = ceil_div(rows, 64), tiles_k = ceil_div(k, 64);
        const size_t nt = tiles_r * tiles_k;
        const unsigned grid = unsigned(nt < cap ? (nt ? nt : 1) : cap);
        if (split && vec) JZ_LAUNCH((prep_transpose_kernel<true, true>), grid, 256, 0, s, hi, lo, kp, src, ld, rows, k, tiles_r, tiles_k);
        else if (split) JZ_LAUNCH((prep_transpose_kernel<true, false>), grid, 256, 0, s, hi, lo, kp, src, ld, rows, k, tiles_r, tiles_k);
        else if (vec) JZ_LAUNCH((prep_transpose_kernel<false, true>), grid, 256, 0, s, hi, lo, kp, src, ld, rows, k, tiles_r, tiles_k);
        else JZ_LAUNCH((prep_transpose_kernel<false, false>), grid, 256, 0, s, hi, lo, kp, src, ld, rows, k, tiles_r, tiles_k);
    }
    return JZ_OK;
}

template <int CG, int MODE, int TN, bool AMN, bool BMN>
static int launch_tc(const Operand& a, const Operand& b, const GemmArgs& args_in, cudaStream_t s) {
    GemmArgs args = args_in;
    args.tiles_m = unsigned(ceil_div(args.m, size_t(CG * TILE_M)));
    args.tiles_n = unsigned(ceil_div(args.n, size_t(TN)));
    alignas(64) CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    int rc;
    if ((rc = make_map(&ma_hi, a.hi, args.m, args.k, a.stride, TILE_M, AMN)) != JZ_OK) return rc;
    if ((rc = make_map(&mb_hi, b.hi, args.n, args.k, b.stride, b_rows<CG, TN>(), BMN)) != JZ_OK) return rc;
    if (MODE == MODE_PRESPLIT) {
        if ((rc = make_map(&ma_lo, a.lo, args.m, args.k, a.stride, TILE_M, false)) != JZ_OK) return rc;
        if ((rc = make_map(&mb_lo, b.lo, args.n, args.k, b.stride, b_rows<CG, TN>(), false)) != JZ_OK) return rc;
    } else {
        ma_lo = ma_hi;
        mb_lo = mb_hi;
    }
    auto kern = gemm_tcgen05_kernel<CG, MODE, TN, AMN, BMN>;
    constexpr int SMEM = smem_bytes<CG, MODE, TN>();
    static bool attr_done = false;
    if (!attr_done) {
        JZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_done = true;
    }
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(args.tiles_m * args.tiles_n * CG, 1, 1);
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ma_hi, ma_lo, mb_hi, mb_lo, args);
    ctx().launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return cuda_fail(e, "gemm_tcgen05_kernel launch");
    return JZ_OK;
}

// operand majors are compile-time (they select TMA box shapes and descriptor layouts)
template <int CG, int MODE, int TN>
static int launch_tc_major(const Operand& a, const Operand& b, const GemmArgs& args, cudaStream_t s) {
    if constexpr (MODE == MODE_PRESPLIT) {
        return launch_tc<CG, MODE, TN, false, false>(a, b, args, s);
    } else {
        if (a.mn) return b.mn ? launch_tc<CG, MODE, TN, true, true>(a, b, args, s) : launch_tc<CG, MODE, TN, true, false>(a, b, args, s);
        return b.mn ? launch_tc<CG, MODE, TN, false, true>(a, b, args, s) : launch_tc<CG, MODE, TN, false, false>(a, b, args, s);
    }
}
template <int MODE>
static int launch_tc_shape(int cg, int tn, const Operand& a, const Operand& b, const GemmArgs& args, cudaStream_t s) {
    if (cg == 2 && tn == 256) return launch_tc_major<2, MODE, 256>(a, b, args, s);
    if (cg == 2) return launch_tc_major<2, MODE, 128>(a, b, args, s);
    return launch_tc_major<1, MODE, 256>(a, b, args, s);
}

static int gemm_tc(int ta, int tb, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda,
                   const float* B, size_t ldb, float beta, float* C, size_t ldc, bool split, const ChainParams& chain,
                   float* const* peers, int n_peers, cudaStream_t s) {
    static const bool force_presplit = env_flag("JZ_GEMM_PRESPLIT");   // A/B measurement of the older variant
    const bool presplit = split && force_presplit;
    Operand a, b;
    int rc = prepare_operand(a, A, lda, /*kmajor_src=*/ta != 0, m, k, presplit, s);
    if (rc == JZ_OK) rc = prepare_operand(b, B, ldb, /*kmajor_src=*/tb == 0, n, k, presplit, s);
    if (rc == JZ_OK) {
        GemmArgs args;
        args.m = m; args.n = n; args.k = k;
        args.alpha = alpha; args.beta = beta;
        args.C = C; args.ldc = ldc;
        args.tiles_m = args.tiles_n = 0;
        args.kb_per_chunk = chunk_kb(split);
        args.n_peers = n_peers;
        for (int q = 0; q < JZ_MAX_PEERS; q++) args.peers[q] = q < n_peers ? peers[q] : nullptr;
        args.chain = chain;
        const int cg = pick_cg();
        const int tn = pick_tile_n(m, n, cg);
        if (!split) rc = launch_tc_shape<MODE_TF32>(cg, tn, a, b, args, s);
        else if (presplit) rc = launch_tc_shape<MODE_PRESPLIT>(cg, tn, a, b, args, s);
        else rc = launch_tc_shape<MODE_XFORM>(cg, tn, a, b, args, s);
    }
    if (a.owned) ws_free(a.owned, s);
    if (b.owned) ws_free(b.owned, s);
    if (rc == JZ_OK) ctx().gemm_last_path = 1;
    return rc;
}

}  // namespace tc

static int gemm_local(int ta, int tb, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda,
                      const float* B, size_t ldb, float beta, float* C, size_t ldc, const ChainParams& chain, int mode,
                      float* const* peers, int n_peers, bool* peers_done, cudaStream_t s);

static int gemm_entry(int ta, int tb, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda,
                      const float* B, size_t ldb, float beta, float* C, size_t ldc, const ChainParams& chain, int mode,
                      cudaStream_t s, float* const* peers = nullptr, int n_peers = 0) {
    if (n_peers < 0 || n_peers > JZ_MAX_PEERS || (n_peers && !peers)) return fail(JZ_ERR_ARG, "jz_gemm: bad peer list");
    bool peers_done = false;
    int rc = gemm_local(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain, mode, peers, n_peers, &peers_done, s);
    if (rc != JZ_OK || peers_done || m == 0 || n == 0) return rc;
    // paths without a fused store (SIMT / rank-1): replicate the finished block with strided P2P copies
    for (int q = 0; q < n_peers; q++) {
        rc = jz_copy2d(peers[q], ldc, C, ldc, m, n, 0, s);
        if (rc != JZ_OK) return rc;
    }
    return JZ_OK;
}

static int gemm_local(int ta, int tb, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda,
                      const float* B, size_t ldb, float beta, float* C, size_t ldc, const ChainParams& chain, int mode,
                      float* const* peers, int n_peers, bool* peers_done, cudaStream_t s) {
    if (m == 0 || n == 0) return JZ_OK;
    if (!C) return fail(JZ_ERR_ARG, "jz_gemm: null C");
    if (ldc < m) return fail(JZ_ERR_SHAPE, "jz_gemm: ldc < m");
    if (mode < 0) mode = ctx().gemm_mode;
    if (mode > JZ_GEMM_BF16) return fail(JZ_ERR_ARG, "jz_gemm: bad mode %d", mode);
    if (k > 0) {
        if (!A || !B) return fail(JZ_ERR_ARG, "jz_gemm: null operand");
        if (lda < (ta ? k : m) || ldb < (tb ? n : k)) return fail(JZ_ERR_SHAPE, "jz_gemm: leading dimension too small");
    }
    const size_t cap = size_t(ctx().sm_count) * 8;
    if (k <= 1) {  // k == 0: C = chain(beta*C); k == 1: rank-1 update (the reference's broadcast idiom)
        const float* u = k ? A : nullptr;
        const float* v = k ? B : nullptr;
        const size_t su = ta ? lda : 1, sv = tb ? 1 : ldb;
        size_t gx = ceil_div(m, size_t(256));
        if (gx > cap) gx = cap;
        size_t gy = ceil_div(cap * 4, gx);
        if (gy > n) gy = n;
        if (gy > 65535) gy = 65535;
        const dim3 grid((unsigned)gx, (unsigned)gy, 1);
        JZ_LAUNCH(rank1_kernel, grid, 256, 0, s, m, n, alpha, u, su, v, sv, beta, C, ldc, chain);
        ctx().gemm_last_path = 3;
        return JZ_OK;
    }
    const bool want_tc = (mode == JZ_GEMM_3XTF32 || mode == JZ_GEMM_TF32 || mode == JZ_GEMM_BF16) && ctx().cc_major == 10;
    // tensor path: a reasonably filled tile grid (m, n >= 64, k >= 32) from 2^22 multiply-adds up, and EVERY product
    // beyond the small-product kernel's range (2^26): skinny ones too -- a 4096 x 4096 x 48 product wastes most of its
    // 256-wide tiles and is still an order of magnitude faster there than on the fp32 SIMT kernel
    const double macs = double(m) * double(n) * double(k);
    const bool big_enough = (m >= 64 && n >= 64 && k >= 32 && macs >= double(1 << 22)) || macs > double(1 << 26);
    const bool fits_i32 = m < (size_t(1) << 31) && n < (size_t(1) << 31) && k < (size_t(1) << 31);
    static const bool force_simt = std::getenv("JZ_GEMM_FORCE_SIMT") != nullptr;
    if (want_tc && big_enough && fits_i32 && !force_simt) {
        // BF16 mode currently rides the TF32 kernel (a strict accuracy superset of bf16 inputs)
        const bool split = mode == JZ_GEMM_3XTF32;
        *peers_done = true;
        return tc::gemm_tc(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, split, chain, peers, n_peers, s);
    }
    if (gemm_small_wants(m, n, k))   // latency-bound shapes: the warp-per-tile fp32 kernel (jz_gemm_small.cu)
        return launch_gemm_small(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain, s);
    return launch_simt(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain, s);
}

}  // namespace jz

using namespace jz;

extern "C" {

int jz_gemm(int transA, int transB, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda,
            const float* B, size_t ldb, float beta, float* C, size_t ldc, int mode, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    ChainParams chain;
    chain.n = 0;
    return gemm_entry(transA, transB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain, mode, as_stream(stream));
}

/* strided batch (cublasSgemmStridedBatched in TransformerLayer, ml/layer.hpp:2896-2926): member i uses
   A + i*strideA, B + i*strideB, C + i*strideC.  Attention-sized members (seq x seq x head_dim) run as ONE launch of
   the small-product kernel with the batch on grid.z; large members go through the single-product paths one by one. */
int jz_gemm_strided_batched(int transA, int transB, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda,
                            size_t strideA, const float* B, size_t ldb, size_t strideB, float beta, float* C, size_t ldc,
                            size_t strideC, size_t batch, int mode, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    if (batch == 0 || m == 0 || n == 0) return JZ_OK;
    if (!C) return fail(JZ_ERR_ARG, "jz_gemm_strided_batched: null C");
    if (ldc < m) return fail(JZ_ERR_SHAPE, "jz_gemm_strided_batched: ldc < m");
    if (k > 0) {
        if (!A || !B) return fail(JZ_ERR_ARG, "jz_gemm_strided_batched: null operand");
        if (lda < (transA ? k : m) || ldb < (transB ? n : k)) return fail(JZ_ERR_SHAPE, "jz_gemm_strided_batched: leading dimension too small");
    }
    ChainParams chain;
    chain.n = 0;
    cudaStream_t s = as_stream(stream);
    const bool tc_shape = m >= 64 && n >= 64 && k >= 32 && double(m) * double(n) * double(k) >= double(1 << 24);
    if (k >= 2 && !tc_shape && gemm_small_wants(m, n, k) && batch <= 65535)
        return launch_gemm_small(transA, transB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain, s, batch, strideA, strideB, strideC);
    for (size_t i = 0; i < batch; i++) {
        int rc = gemm_entry(transA, transB, m, n, k, alpha, A + i * strideA, lda, B + i * strideB, ldb, beta, C + i * strideC, ldc,
                            chain, mode, s);
        if (rc != JZ_OK) return rc;
    }
    return JZ_OK;
}

int jz_gemm_chain(int transA, int transB, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda,
                  const float* B, size_t ldb, float* C, size_t ldc, const jz_step* steps, int nsteps, int mode,
                  jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    ChainParams chain;
    if (make_chain(chain, steps, nsteps) != JZ_OK) return fail(JZ_ERR_ARG, "jz_gemm_chain: bad step list");
    return gemm_entry(transA, transB, m, n, k, alpha, A, lda, B, ldb, 0.0f, C, ldc, chain, mode, as_stream(stream));
}

int jz_gemm_chain_bcast(int transA, int transB, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda,
                        const float* B, size_t ldb, float* C, size_t ldc, float* const* peer_C, int n_peers,
                        const jz_step* steps, int nsteps, int mode, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    ChainParams chain;
    if (make_chain(chain, steps, nsteps) != JZ_OK) return fail(JZ_ERR_ARG, "jz_gemm_chain_bcast: bad step list");
    return gemm_entry(transA, transB, m, n, k, alpha, A, lda, B, ldb, 0.0f, C, ldc, chain, mode, as_stream(stream), peer_C,
                      n_peers);
}

}  // extern "C"
