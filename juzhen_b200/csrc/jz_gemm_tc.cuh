// jz_gemm_tc.cuh -- the tensor-core GEMM kernel behind Matrix<CUDAfloat>::dot / operator* (SURVEY 8a row a17,
// replacing cublasSgemm, cpp/cumatrix.cu:177-197, and cublasSgemmStridedBatched, ml/layer.hpp:2896-2926).
// Included by jz_gemm.cu (dispatch) and by the jz_gemm_tc_*.cu translation units that instantiate it (one per
// arithmetic mode and CTA-group size, so the instantiations compile in parallel).
//
// sm_100a: TMA -> shared memory (128B swizzle) -> tcgen05.mma.kind::tf32 with the fp32 accumulator in TMEM ->
// tcgen05.ld -> coalesced column-major stores.
//   * warp-specialised CTA: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..9 = epilogue
//     (one TMEM lane quarter x one column half each) which, in 3xTF32 mode, also compute the lo operand tiles;
//   * CG = 2: a CTA pair (cluster 2x1x1, cta_group::2) computes one 256 x TN tile; each CTA loads its 128 rows of
//     A and its half of the B rows, the leader issues M=256 MMAs that read both CTAs' shared memory.  CG = 1 is the
//     single-SM 128 x TN variant used for narrow / short / batched products (TN = 64, 128, 256);
//   * operands are read by TMA straight from the caller's column-major storage, in either major (K-major: box of
//     32 k x rows, 128B swizzle; MN-major: boxes of 32 rows x 32 k, 128B swizzle with 32-byte atoms); the tensor
//     maps are 3-D, the third coordinate is the member of a strided batch (blockIdx.z);
//   * 3xTF32 (MODE_XFORM, fp32 accuracy): every k-block issues A_lo*B_hi + A_hi*B_lo + A_hi*B_hi into the same
//     TMEM accumulator.  kind::tf32 reads fp32 words and drops the low 13 mantissa bits, so the raw tile already IS
//     the hi operand; the epilogue warps compute lo = rna_tf32(x - trunc_tf32(x)) into a second shared buffer;
//   * 3xTF32 with A in TENSOR MEMORY (MODE_XFORM_TS, CG = 1, TN <= 128: narrow / short outputs and small products of
//     few tiles, where the shared-memory transform and not the tensor pipe sets the pace): the transform warps read A's
//     rows from the swizzled tile, write raw words and lo words to TMEM (tcgen05.st, one slot per stage beside the two
//     accumulator buffers) and the MMAs take A from TMEM (tcgen05.mma [d], [a], b_desc); B keeps the shared form;
//   * two-level accumulation: only kb_per_chunk k-blocks are chained inside TMEM (the tensor core accumulates with
//     truncation), the epilogue warps add each chunk into fp32 registers with round-to-nearest while the MMA warp
//     fills the other TMEM half;
//   * SPLIT-K UNITS: the launch is a list of units in block order -- first `full_tiles` whole tiles (whole waves of
//     the tile grid, hardware-dispatched in lockstep so the tiles of a wave share their A/B panels in L2), then the
//     tiles of the partial last wave, each split along k into `splits` units so the tail fills every SM.  The units
//     of a split tile write their register accumulators to workspace, take a ticket, wait until all `splits` partials
//     of the tile are there, and each finishes a 1/splits column slice of the tile: partials summed in split order
//     (bitwise deterministic whatever the arrival order), alpha/beta/chain, store.  All units of a split tile sit in
//     consecutive blocks of one launch, so the wait cannot deadlock under in-order block dispatch;
//   * CLUSTER SPLIT-K (products of fewer tiles than a quarter of the SMs, 2 or 4 splits): the units of a tile are the
//     CTAs (CTA pairs) of one thread-block cluster; after its mainloop every unit writes the column slices it does not
//     own straight into the owner's shared memory (st.shared::cluster, the operand stages are idle by then), and the
//     owner adds the partials in split order -- same bits as the workspace form, without the trip through L2;
//   * WALKED BATCHES (strided batches with more (member, tile) units than SMs): a CTA keeps its barriers, its tensor-memory
//     allocation and its TMA pipeline and walks the units u, u + P, u + 2P ...: the producer runs ahead into the next unit's
//     operands while the epilogue warps store the current one, and the per-CTA prologue (a large part of a unit of 4 k-blocks)
//     is paid once per SM instead of once per unit;
//   * programmatic dependent launch: everything before the first global-memory access (barrier init, TMEM
//     allocation, tensor-map prefetch) may overlap the tail of the previous kernel in the stream;
//   * fused all-gather epilogue (multi-GPU): finished elements are also stored to peer images of C, either one
//     P2P store per peer or ONE multimem.st to an NVSwitch multicast address that lands in every GPU's image.
#pragma once
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "jz_common.cuh"
#include "jz_math.cuh"

namespace jz {
namespace tc {

constexpr int BK = 32;               // fp32 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 8;            // tf32: 32 bytes per MMA k-step
constexpr int TILE_M = 128;          // rows of A per CTA (TMEM lanes)
constexpr int A_BYTES = TILE_M * BK * 4;  // 16 KB
constexpr int NUM_EPI_WARPS = 8;
constexpr int FIRST_EPI_WARP = 2;
constexpr int NUM_THREADS = 32 * (FIRST_EPI_WARP + NUM_EPI_WARPS);  // TMA warp + MMA warp + 8 epilogue/transform warps
constexpr int MODE_TF32 = 0;      // single pass over the raw fp32 tiles
constexpr int MODE_XFORM = 2;     // 3xTF32, lo computed in shared memory by the transform warps
constexpr int MODE_XFORM_TS = 3;  // 3xTF32, A operand (hi and lo) written to TENSOR MEMORY by the transform warps (CG = 1, TN <= 128)
constexpr int TS_SLOT_COLS = 2 * BK;   // TMEM columns of one k-block of A: 32 hi + 32 lo
constexpr int MN_BOX_BYTES = 32 * BK * 4;  // one MN-major TMA box: 32 contiguous rows x 32 k = 4 KB
constexpr int MAX_SPLITS = 16;

template <int CG, int TN> __host__ __device__ constexpr int b_rows() { return TN / CG; }
template <int CG, int TN> __host__ __device__ constexpr int b_bytes() { return b_rows<CG, TN>() * BK * 4; }
template <int CG, int MODE, int TN> __host__ __device__ constexpr int stage_bytes() {
    if (MODE == MODE_XFORM_TS) return A_BYTES + 2 * b_bytes<CG, TN>();   // raw A, raw B, lo(B): lo(A) lives in TMEM
    return (MODE != MODE_TF32 ? 2 : 1) * (A_BYTES + b_bytes<CG, TN>());
}
template <int CG, int MODE, int TN> __host__ __device__ constexpr int num_stages() {
    int s = (227 * 1024 - 2048) / stage_bytes<CG, MODE, TN>();
    if (MODE == MODE_XFORM_TS) {   // one TMEM slot per smem stage, next to the two accumulator buffers
        const int slots = (512 - 2 * TN) / TS_SLOT_COLS;
        if (s > slots) s = slots;
    }
    return s > 8 ? 8 : s;
}
// TMEM-A variant: two transform groups of four warps (one per TMEM lane quarter) take the k-blocks alternately.  (A third,
// transform-only group of four more warps was measured at TN = 64: 73.6 against 73.0 us at 8192 x 32 x 8192 -- the
// transform is bound by what the groups share, tensor-memory store and integer-pipe throughput, not by a warp's latency.)
constexpr int TS_GROUPS = 2;
template <int CG, int MODE, int TN> __host__ __device__ constexpr int smem_bytes() {
    return num_stages<CG, MODE, TN>() * stage_bytes<CG, MODE, TN>() + 1024 /*align slack*/ + 256 /*barriers*/;
}

struct GemmArgs {
    size_t m, n, k;
    float alpha, beta;
    float* C;
    size_t ldc;
    size_t strideC;             // elements between batch members of C
    unsigned tiles_m, tiles_n;  // in units of (CG*128) x TN tiles
    unsigned full_tiles;        // units [0, full_tiles) are whole tiles; later units are k-splits of the remaining tiles
    int splits;                 // k-splits per split tile (>= 2 when full_tiles < tiles_m*tiles_n)
    int kb_per_split;           // k-blocks per split unit
    int kb_per_chunk;           // k-blocks accumulated inside TMEM before promotion to registers
    unsigned walk_units;        // strided batches of many small members: a CTA (pair) WALKS the (member, tile) units u, u + P,
                                // u + 2P ... (P = CTA groups in the grid) instead of exiting after one -- barriers, tensor
                                // memory and the TMA pipeline live across units (0: one unit per CTA group)
    int cluster_split;          // the `splits` units of a tile form ONE thread-block cluster and exchange their partial tiles
                                // through distributed shared memory (no workspace, no tickets); needs full_tiles == 0
    float* ws;                  // split-K partial tiles: [split tile][split][rank][TN columns][128 rows]
    unsigned* tickets;          // per-stream self-resetting counters: arrivals at [tile], departures at [kTicketSlots/2 + tile]
    int n_peers;                // additional destinations (peer-GPU images of C, same ldc)
    float* peers[JZ_MAX_PEERS];
    float* mc;                  // multicast (NVSwitch) image of C: when set, every element is stored by ONE multimem.st
    ChainParams chain;
#ifdef JZ_GEMM_PROFILE
    // instrumented build (scripts/build_prof_lib.sh): cycle counts of CTA 5, read back by jz_debug_gemm_prof
    //   [0..2] TMA: k-blocks, total, waiting for a free stage      [3..5] MMA: total, waiting for operands, for a drained accumulator
    //   [8 + 8 g ..] transform group g (warp e = 4 g): mainloop, waiting for TMA, transform, waiting for a chunk, drain
    //   [24..26] epilogue: cycles, split, splits      [29..31] split epilogue: partial tile out, fence + barrier, ticket + wait
    long long* prof;
#endif
};

struct Operand {
    const float* ptr = nullptr;  // raw fp32
    size_t stride = 0;           // elements between consecutive rows (K-major) / consecutive k (MN-major)
    size_t batch_stride = 0;     // elements between batch members
    bool mn = false;             // MN-major: element (r, kk) at kk*stride + r
    void* owned = nullptr;       // workspace to release
};

// jz_gemm_tc_*.cu: one definition per MODE x CG
template <int MODE, int CG>
int launch_tc_cg(int tn, const Operand& a, const Operand& b, const GemmArgs& args, unsigned batch, cudaStream_t s);
// jz_gemm_tc_xform_ts.cu: 3xTF32 with A staged in tensor memory (CG = 1, tn = 64 | 128)
int launch_tc_ts(int tn, const Operand& a, const Operand& b, const GemmArgs& args, unsigned batch, cudaStream_t s);
// jz_gemm_tc_tf32_persist.cu: the persistent single-pass TF32 kernel (CTA pairs, 256 x 256 tiles, single products)
int launch_tc_tf32_persistent(const Operand& a, const Operand& b, const GemmArgs& args, cudaStream_t s);

#ifdef JZ_GEMM_TC_IMPL   // ------------------------------------------------------------------ kernel side


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
// programmatic dependent launch: wait for the prerequisite grids (and their memory) / let the dependent grid start
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// TO_LEADER: the copy (issued by either CTA of a pair) signals the LEADER CTA's barrier (peer bit cleared);
// otherwise it signals the issuing CTA's own barrier.  c2 = batch member.
template <bool TO_LEADER>
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    if constexpr (TO_LEADER) {
        asm volatile(
            "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
            : "memory");
    } else {
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
            : "memory");
    }
}
// One operand tile of `ROWS` rows x BK k starting at (row0, kc).
//   K-major source: one box {BK, ROWS}: ROWS rows of 128 B, 128B swizzle.
//   MN-major source: ROWS/32 boxes {32 rows, BK}: BK rows of 128 B each holding 32 consecutive operand rows
//   (4 KB per box, 128B swizzle with 32-byte atoms).
template <bool MN, int ROWS, bool TO_LEADER>
__device__ __forceinline__ void tma_load_tile(uint32_t dst, const CUtensorMap* map, uint32_t bar, int kc, int row0, int bz) {
    if constexpr (MN) {
#pragma unroll
        for (int i = 0; i < ROWS / 32; i++) tma_load_3d<TO_LEADER>(dst + i * MN_BOX_BYTES, map, bar, row0 + 32 * i, kc, bz);
    } else {
        tma_load_3d<TO_LEADER>(dst, map, bar, kc, row0, bz);
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
// A operand from tensor memory (lane = row, one fp32 word per column = k), B from shared memory; CTA group 1 only
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc),
        "r"(idesc), "r"(accumulate)
        : "memory");
}
// tcgen05.commit: arrive on `bar` (in every CTA of the pair for CG == 2) when all prior MMAs retire
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar, uint16_t pair_mask = 3) {
    if constexpr (CG == 2) {   // pair_mask: the two CTAs of this pair inside the cluster (3 << even rank)
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
            "h"(pair_mask)
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

// 16 consecutive TMEM columns of this thread's lane (32 lanes x 32 bit per warp); caller issues tmem_st_wait()
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

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

__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// named barrier over the 8 epilogue warps only (barrier 0 belongs to __syncthreads)
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// one store that NVSwitch replicates into every GPU's image of the buffer (multicast address)
__device__ __forceinline__ void multimem_st(float* mc_addr, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc_addr), "f"(v) : "memory");
}

// lo part of the 3xTF32 split against the hardware's own hi: kind::tf32 drops the low 13 mantissa bits of the
// fp32 word it reads, so hi = x & 0xFFFFE000 and x - hi is exact in fp32; lo is that remainder rounded to tf32
// (nearest, ties away: add half an ulp to the magnitude, truncate), so the tensor core reads it unchanged.
// inf/nan keep their semantics through hi alone (lo = 0 avoids inf - inf); the one corner left is hi(a) = inf against an
// element of the other operand whose lo is exactly 0 (one fp32 value in 8192): inf * 0 = nan where fp32 gives inf.
__device__ __forceinline__ float tf32_lo_of(float x) {
    const uint32_t b = __float_as_uint(x);
    const float d = __fsub_rn(x, __uint_as_float(b & 0xFFFFE000u));
    const uint32_t r = __float_as_uint(d) + 0x1000u;   // the low 13 bits that remain are dropped by the tensor core
    return d == d ? __uint_as_float(r) : 0.0f;   // x = inf/nan gives d = nan
}

// ---- cluster split-K: exchange of partial tiles through distributed shared memory
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// A thread holds row (quarter*32 + lane) x columns [half*TN/2, (half+1)*TN/2) of its unit's partial tile in acc[].  Slice d
// (columns [d*W, (d+1)*W), W = TN/S) is finished by unit d: every other unit writes its part of that slice into unit d's
// shared memory at [slot][column of the slice][128 rows], slot = the sender's split index with the owner's skipped.
template <int CG, int TN, int S>
__device__ __forceinline__ void cluster_split_send(const float (&acc)[TN / 2], uint32_t recv_base, int split, int half, int row,
                                                   uint32_t rank) {
    constexpr int W = TN / S, DPH = (TN / 2) / W;
    if constexpr (W >= 32 && DPH >= 1) {
#pragma unroll
        for (int j = 0; j < DPH; j++) {
            const int d = half * DPH + j;
            if (d == split) continue;
            const uint32_t slot = uint32_t(split < d ? split : split - 1);
            const uint32_t dst = mapa_shared(recv_base, uint32_t(d) * CG + rank) + ((slot * W) * TILE_M + uint32_t(row)) * 4u;
#pragma unroll
            for (int c = 0; c < W; c++) st_cluster_f32(dst + uint32_t(c) * (TILE_M * 4u), acc[j * W + c]);
        }
    }
}
// The owner's side: partials added in split order (its own from registers at its place in the order), so the sum has
// the same bits as the workspace form and does not depend on arrival order.
template <int CG, int TN, int S>
__device__ __forceinline__ void cluster_split_reduce(float (&acc)[TN / 2], const float* recv, int split, int half, int row) {
    constexpr int W = TN / S, DPH = (TN / 2) / W;
    if constexpr (W >= 32 && DPH >= 1) {
#pragma unroll
        for (int j = 0; j < DPH; j++) {
            if (half * DPH + j != split) continue;
#pragma unroll   // fully: acc[] must keep static indices (registers)
            for (int c = 0; c < W; c++) {
                float in[S - 1];
#pragma unroll
                for (int t = 0; t < S - 1; t++) in[t] = recv[(size_t(t) * W + c) * TILE_M + row];
                const float own = acc[j * W + c];
                float v = split == 0 ? own : in[0];
#pragma unroll
                for (int pos = 1; pos < S; pos++) {
                    const float term = pos < split ? in[pos < S - 1 ? pos : S - 2] : (pos == split ? own : in[pos - 1]);
                    v = __fadd_rn(v, term);
                }
                acc[j * W + c] = v;
            }
        }
    }
}

// One unit (a whole (CG*128) x TN output tile, or one k-split of it) per CTA group.
//
// Barriers (per smem stage): MODE_TF32: TMA of both CTAs -> full (leader) -> MMA -> empty (both).
// MODE_XFORM: TMA -> full (own CTA) -> epilogue warps write lo -> ready (leader) -> MMA -> empty (both).
template <int CG, int MODE, int TN, bool AMN, bool BMN, bool WALK>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs args) {
    constexpr bool TS = MODE == MODE_XFORM_TS;       // A operand (hi, lo) staged in tensor memory
    constexpr bool XFORM = MODE == MODE_XFORM || TS;
    static_assert(!TS || (CG == 1 && TN <= 128), "the TMEM-A variant is single-CTA and needs TMEM columns beside the accumulators");
    constexpr bool TO_LEADER = CG == 2 && !XFORM;   // whose `full` barrier the TMA copies signal
    constexpr int TILE_N = TN;
    constexpr int HALF_N = TN / 2;       // columns drained by one epilogue warp
    constexpr int STAGES = num_stages<CG, MODE, TN>();
    constexpr int STAGE_BYTES = stage_bytes<CG, MODE, TN>();
    constexpr int B_ROWS = b_rows<CG, TN>();
    constexpr int B_BYTES = b_bytes<CG, TN>();
    constexpr int RAW_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t IDESC = make_idesc_tf32(CG * 128, TILE_N, TS ? false : AMN, BMN);   // A from TMEM is always [row][k]
    constexpr uint32_t KA = kstep_bytes<AMN>(), KB = kstep_bytes<BMN>();
    constexpr int TILE_ELEMS = TILE_M * TN;   // one CTA's share of a tile (workspace layout: [column][128 rows])
    constexpr int TMEM_COLS = TS ? 512 : 2 * TILE_N;   // two accumulator buffers (+ STAGES slots of A for TS)
    constexpr int GROUPS = TS ? TS_GROUPS : 1;   // transform groups taking k-blocks round-robin
    constexpr bool DEFER = TS && STAGES >= 6;   // hand a k-block to the MMA warp only when the group's next one is in registers

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
    const uint32_t crank = cluster_ctarank();          // rank in the cluster: pair rank, or split * CG + pair rank (cluster split)
    const uint32_t rank = CG == 2 ? crank & 1u : 0u;   // rank in the CTA pair
    const bool leader = rank == 0;
    const bool cs = args.cluster_split != 0;           // uniform over the launch
    const uint16_t pair_mask = uint16_t(3u << (crank & ~1u));

    // unit -> (tile, k range).  Whole tiles first, then the k-splits of the remaining tiles.
    const unsigned unit = CG == 2 ? blockIdx.x >> 1 : blockIdx.x;
    const unsigned walk = WALK ? args.walk_units : 0u;           // > 0: this CTA group walks units unit, unit + ustride, ...
                                                                 // (a compile-time variant: the one-unit kernels keep their registers)
    const unsigned ustride = (CG == 2 ? gridDim.x >> 1 : gridDim.x);
    const int num_kb_total = int((args.k + BK - 1) / BK);
    unsigned tile = unit;
    int kb0 = 0, num_kb = num_kb_total, split = -1;
    unsigned split_tile = 0;
    if (!walk && unit >= args.full_tiles) {
        const unsigned r = unit - args.full_tiles;
        split_tile = r / unsigned(args.splits);
        split = int(r % unsigned(args.splits));
        tile = args.full_tiles + split_tile;
        kb0 = split * args.kb_per_split;
        num_kb = num_kb_total - kb0 < args.kb_per_split ? num_kb_total - kb0 : args.kb_per_split;
    }
    // tile coordinates (grouped rasterisation for L2 reuse of the A / B panels); a walked unit u is tile u % tiles of
    // member u / tiles.  Every warp role keeps its own copy and re-places it per unit.
    struct Place { int m0, n0, nb0, bz; };
    auto place = [&](unsigned u) -> Place {
        unsigned t = u;
        int z = int(blockIdx.z);
        if (walk) {
            const unsigned tiles = args.tiles_m * args.tiles_n;
            z = int(u / tiles);
            t = u % tiles;
        }
        constexpr unsigned GROUP = 8;
        const unsigned per_group = GROUP * args.tiles_n;
        const unsigned group_id = t / per_group;
        const unsigned first_m = group_id * GROUP;
        const unsigned group_m = args.tiles_m - first_m < GROUP ? args.tiles_m - first_m : GROUP;
        const unsigned tm = first_m + (t % per_group) % group_m;
        const unsigned tn = (t % per_group) / group_m;
        const int n0_ = int(tn) * TILE_N;
        return Place{int(tm) * (CG * TILE_M) + int(rank) * TILE_M,                 // first row of this CTA
                     n0_, n0_ + (CG == 2 ? int(rank) * (TILE_N / 2) : 0), z};       // first B row (= C column) this CTA loads
    };
    const Place first_place = place(walk ? unit : tile);
    const int kbc = args.kb_per_chunk;
    const int num_chunks = (num_kb + kbc - 1) / kbc;

    if (warp == 0 && elect_one()) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1) {
        if (elect_one()) {
            for (int s = 0; s < STAGES; s++) {
                mbar_init(full_bar(s), 1);
                mbar_init(empty_bar(s), 1);
                mbar_init(ready_bar(s), TS ? NUM_EPI_WARPS / 2 : NUM_EPI_WARPS * CG);   // every transform (= epilogue) warp of the pair; TS: of one group
            }
            for (int b = 0; b < 2; b++) {
                mbar_init(tmem_full_bar(b), 1);
                mbar_init(tmem_empty_bar(b), NUM_EPI_WARPS * CG);  // every epilogue warp of the pair
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<CG>(tmem_slot, TMEM_COLS);
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // everything above touched no global memory: it may run while the previous kernel of the stream drains
    pdl_wait();
    pdl_launch_dependents();

    // hand a transformed k-block to the MMA warp: tensor-memory stores retired, generic-proxy writes of lo(B) visible
    // to the tensor core, one arrive per warp on the stage's `ready` barrier
    auto ts_hand_over = [&](uint32_t& pend_ready) {
        tmem_st_wait();
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive_cluster(pend_ready);
        pend_ready = 0;
    };
    // TMEM-A transform of one landed stage by one group of four warps (`quarter` = the warp's TMEM lane quarter, `tg` = the
    // thread's index in the group).  A: this thread's row (quarter * 32 + lane) from the swizzled tile -> raw words + lo
    // words -> the stage's TMEM slot.  The slot is free: the TMA that filled this stage waited for the MMAs of its previous
    // use.  With DEFER the hand-over of a k-block waits until the first half of the group's NEXT k-block sits in registers,
    // so the tensor-memory store latency hides under that work instead of ending the per-k-block dependency chain
    // (ablation: the stores + their wait were 28 % of the mainloop).
    auto ts_transform = [&](uint32_t stage, int quarter, int tg, uint32_t ready0, uint32_t& pend_ready) {
        if constexpr (TS) {
            const int lane = threadIdx.x & 31;
            const uint8_t* const sA = gen_base + stage * STAGE_BYTES;
            const uint32_t a_t = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(2 * TILE_N + int(stage) * TS_SLOT_COLS);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint32_t hi[16], lo[16];
                if constexpr (AMN) {
                    // box `quarter` = these 32 rows; line k = 128 B of 32 consecutive rows, its four 32-byte atoms
                    // XOR-swizzled by (k & 3): a warp reads one line per k, conflict-free
                    const uint8_t* const box = sA + quarter * MN_BOX_BYTES;
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const int k = h * 16 + i;
                        hi[i] = *reinterpret_cast<const uint32_t*>(box + k * 128 + ((((lane >> 3) ^ (k & 3)) << 5) | ((lane & 7) << 2)));
                    }
                } else {
                    // row r = 128 B of 32 k, its eight 16-byte chunks XOR-swizzled by (r & 7): a quarter warp reads eight
                    // different chunk positions, conflict-free
                    const int r = quarter * 32 + lane;
                    const uint8_t* const rowp = sA + r * 128;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint4 w = *reinterpret_cast<const uint4*>(rowp + (((h * 4 + q) ^ (r & 7)) << 4));
                        hi[4 * q] = w.x; hi[4 * q + 1] = w.y; hi[4 * q + 2] = w.z; hi[4 * q + 3] = w.w;
                    }
                }
#pragma unroll
                for (int i = 0; i < 16; i++) lo[i] = __float_as_uint(tf32_lo_of(__uint_as_float(hi[i])));
                if (DEFER && h == 0 && pend_ready) ts_hand_over(pend_ready);   // the group's previous k-block
                tmem_st_32x32b_x16(a_t + h * 16, hi);
                tmem_st_32x32b_x16(a_t + BK + h * 16, lo);
            }
            // B: lo(B) beside the raw tile, by the 128 threads of this group
            constexpr int XG = 128, PERG = B_BYTES / 16 / XG;
            static_assert(PERG % 4 == 0, "B tile must divide among the group's threads");
            const float4* src = reinterpret_cast<const float4*>(sA + A_BYTES) + tg;
            float4* dst = reinterpret_cast<float4*>(gen_base + stage * STAGE_BYTES + A_BYTES + B_BYTES) + tg;
#pragma unroll
            for (int i0 = 0; i0 < PERG; i0 += 4) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = src[(i0 + u) * XG];
#pragma unroll
                for (int u = 0; u < 4; u++)
                    dst[(i0 + u) * XG] = make_float4(tf32_lo_of(v[u].x), tf32_lo_of(v[u].y), tf32_lo_of(v[u].z), tf32_lo_of(v[u].w));
            }
            pend_ready = ready0 + 8u * stage;
            if (!DEFER) ts_hand_over(pend_ready);   // few stages: a late hand-over would starve the TMA of free stages
        }
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
#ifdef JZ_GEMM_PROFILE
            long long pf_wait = 0, pf_t0 = clock64();
#endif
            Place pl = first_place;
            for (unsigned u = unit;;) {
            for (int kb = 0; kb < num_kb; kb++) {
#ifdef JZ_GEMM_PROFILE
                const long long pf_a = clock64();
#endif
                mbar_wait(empty_bar(stage), phase ^ 1);
#ifdef JZ_GEMM_PROFILE
                pf_wait += clock64() - pf_a;
                if (kb == num_kb - 1 && blockIdx.x == 5 && blockIdx.z == 0)
                    { args.prof[0] = num_kb; args.prof[1] = clock64() - pf_t0; args.prof[2] = pf_wait; }
#endif
                if (XFORM) mbar_arrive_expect_tx(full_bar(stage), uint32_t(RAW_BYTES));
                else if (leader) mbar_arrive_expect_tx(full_bar(stage), uint32_t(STAGE_BYTES) * CG);
                const uint32_t sa = smem_base + stage * STAGE_BYTES;
                const int kc = (kb0 + kb) * BK;
                tma_load_tile<AMN, TILE_M, TO_LEADER>(sa, &tmA, full_bar(stage), kc, pl.m0, pl.bz);
                tma_load_tile<BMN, B_ROWS, TO_LEADER>(sa + A_BYTES, &tmB, full_bar(stage), kc, pl.nb0, pl.bz);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (!walk || (u += ustride) >= walk) break;
            pl = place(u);   // next unit of the walk: the pipeline simply continues into its operands
            }
        }
        if (cs) { __syncwarp(); cluster_sync_all(); cluster_sync_all(); }   // the exchange's two cluster barriers (below)
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            uint32_t stage = 0, phase = 0;
            int chunk = 0, in_chunk = 0;
#ifdef JZ_GEMM_PROFILE
            long long pf_ready = 0, pf_tmem = 0, pf_t0 = clock64();
#endif
            for (unsigned u = unit;;) {   // (chunk counts on across the units of a walk: the two TMEM buffers keep alternating)
            for (int kb = 0; kb < num_kb; kb++) {
                const uint32_t buf = uint32_t(chunk) & 1u;
#ifdef JZ_GEMM_PROFILE
                const long long pf_a = clock64();
#endif
                if (in_chunk == 0) {  // this TMEM buffer must have been drained by every epilogue warp
                    mbar_wait(tmem_empty_bar(buf), ((uint32_t(chunk) >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                }
#ifdef JZ_GEMM_PROFILE
                const long long pf_b = clock64();
#endif
                mbar_wait(XFORM ? ready_bar(stage) : full_bar(stage), phase);
                tc_fence_after();
#ifdef JZ_GEMM_PROFILE
                pf_tmem += pf_b - pf_a;
                pf_ready += clock64() - pf_b;
                if (kb == num_kb - 1 && blockIdx.x == 5 && blockIdx.z == 0 && (threadIdx.x & 31) == 0)
                    { args.prof[3] = clock64() - pf_t0; args.prof[4] = pf_ready; args.prof[5] = pf_tmem; }
#endif
                const bool chunk_end = (in_chunk == kbc - 1) || (kb == num_kb - 1);
                if (elect_one()) {
                    const uint32_t d = tmem_base + buf * TILE_N;
                    const uint32_t sa = smem_base + stage * STAGE_BYTES;
                    uint32_t acc = in_chunk == 0 ? 0u : 1u;
                    if constexpr (TS) {
                        // A: TMEM slot of this stage (columns [0,32) = raw words, the tensor core reads their top 19
                        // bits = hi; [32,64) = lo); B: raw tile and lo(B) in the smem stage
                        const uint32_t a_t = tmem_base + uint32_t(2 * TILE_N + int(stage) * TS_SLOT_COLS);
                        const uint32_t b_hi = sa + A_BYTES, b_lo = sa + A_BYTES + B_BYTES;
#pragma unroll
                        for (int ks = 0; ks < BK / UMMA_K; ks++) {
                            umma_tf32_ts(d, a_t + BK + ks * UMMA_K, make_smem_desc<BMN>(b_hi + ks * KB), IDESC, acc);
                            acc = 1u;
                        }
#pragma unroll
                        for (int ks = 0; ks < BK / UMMA_K; ks++)
                            umma_tf32_ts(d, a_t + ks * UMMA_K, make_smem_desc<BMN>(b_lo + ks * KB), IDESC, 1u);
#pragma unroll
                        for (int ks = 0; ks < BK / UMMA_K; ks++)
                            umma_tf32_ts(d, a_t + ks * UMMA_K, make_smem_desc<BMN>(b_hi + ks * KB), IDESC, 1u);
                    } else {
                        const uint32_t a_hi = sa, b_hi = sa + A_BYTES;
                        const uint32_t a_lo = sa + RAW_BYTES, b_lo = sa + RAW_BYTES + A_BYTES;
                        if (XFORM) {
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
                    }
                    umma_commit<CG>(empty_bar(stage), pair_mask);                   // frees this smem stage (both CTAs)
                    if (chunk_end) umma_commit<CG>(tmem_full_bar(buf), pair_mask);  // chunk accumulator complete
                }
                __syncwarp();
                if (chunk_end) { chunk++; in_chunk = 0; } else { in_chunk++; }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (!walk || (u += ustride) >= walk) break;
            }
        }
        if (cs) { __syncwarp(); cluster_sync_all(); cluster_sync_all(); }   // the exchange's two cluster barriers (below)
    } else {
        // ===================== epilogue warps: lo-part transform of landed stages (MODE_XFORM), =====================
        // ===================== TMEM chunks -> fp32 registers (RN) -> global                     =====================
        const int e = warp - FIRST_EPI_WARP;
        const int quarter = warp & 3;   // TMEM lane quarter this warp may access (hardware: warp id % 4)
        const int half = e >> 2;        // which half of the accumulator columns
        const int lane = threadIdx.x & 31;
        float acc[HALF_N];
        const uint32_t lane_addr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(half * HALF_N);
        // on the leader CTA of the pair (bit 24 of a shared-window address = low bit of the CTA's rank in its cluster); a
        // single-CTA unit keeps its own barriers -- in a cluster-split launch it has a rank of its own
        constexpr uint32_t LEADER_MASK = CG == 2 ? 0xFEFFFFFFu : 0xFFFFFFFFu;
        const uint32_t empty0 = tmem_empty_bar(0) & LEADER_MASK, empty1 = tmem_empty_bar(1) & LEADER_MASK;
        const uint32_t ready0 = ready_bar(0) & LEADER_MASK;
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
        uint32_t pend_ready = 0;   // TS: `ready` barrier of the group's transformed but not yet handed-over k-block
        int chunk_base = 0;        // chunks of the units this CTA has finished (a walk: the TMEM buffers keep alternating)
        const int kb_end = XFORM ? num_kb : 0;
#ifdef JZ_GEMM_PROFILE
        long long pf_full = 0, pf_xf = 0, pf_dw = 0, pf_dr = 0, pf_t0 = clock64(), pf_own = 0, pf_skip = 0;
#endif
        Place pl = first_place;
        for (unsigned u = unit;;) {   // one iteration unless this CTA walks a batch
        const int m0 = pl.m0, n0 = pl.n0, bz = pl.bz;
        int next_drain = 0;
#pragma unroll
        for (int i = 0; i < HALF_N; i++) acc[i] = 0.0f;
        for (int kb = 0; kb <= kb_end; kb++) {
#ifdef JZ_GEMM_PROFILE
            const long long pf_it = clock64();
#endif
            if (TS && pend_ready && kb + (half - kb % GROUPS + GROUPS) % GROUPS >= kb_end) {
                // the group has no further k-block: nothing left to hide the hand-over of its last one under.  (Drains of
                // earlier iterations never wait on a pending hand-over: chunk c is drained at kb >= 4c + 4 + STAGES - 1 and
                // a pending k-block is at most GROUPS behind kb, i.e. > 4c + 3 for STAGES > GROUPS.)
                ts_hand_over(pend_ready);
            }
            while (next_drain < num_chunks && (kb >= kb_end || (next_drain + 1) * kbc + STAGES - 1 <= kb)) {
                const int chunk = chunk_base + next_drain++;
                const uint32_t buf = uint32_t(chunk) & 1u;
#ifdef JZ_GEMM_PROFILE
                const long long pf_a = clock64();
#endif
                mbar_wait(tmem_full_bar(buf), (uint32_t(chunk) >> 1) & 1u);
                tc_fence_after();
#ifdef JZ_GEMM_PROFILE
                pf_dw += clock64() - pf_a;
                pf_dr -= clock64();
#endif
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
#ifdef JZ_GEMM_PROFILE
                pf_dr += clock64();
#endif
            }
            if (XFORM && kb < kb_end) {
                // TS: the two warp groups (half = 0 / 1, each with all four TMEM lane quarters) take the k-blocks alternately,
                // so two transforms are in flight and one's shared-memory / tcgen05.st latency hides under the other's
                if (!TS || kb % GROUPS == half) {
#ifdef JZ_GEMM_PROFILE
                    const long long pf_a = clock64();
#endif
                    mbar_wait(full_bar(stage), phase);
#ifdef JZ_GEMM_PROFILE
                    pf_full += clock64() - pf_a;
                    pf_xf -= clock64();
#endif
                    if constexpr (TS) {
                        ts_transform(stage, quarter, te & 127, ready0, pend_ready);
                    } else {
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
                    }
                    if constexpr (!TS) {
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(ready0 + 8u * stage);
                    }
#ifdef JZ_GEMM_PROFILE
                    pf_xf += clock64();
#endif
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
#ifdef JZ_GEMM_PROFILE
            if (!TS || kb % GROUPS == half) pf_own += clock64() - pf_it; else pf_skip += clock64() - pf_it;
#endif
        }
#ifdef JZ_GEMM_PROFILE
        if (blockIdx.x == 5 && blockIdx.z == 0 && lane == 0 && (e & 3) == 0)
        {
            long long* g = args.prof + 8 + 8 * (e >> 2);
            g[0] = clock64() - pf_t0; g[1] = pf_full; g[2] = pf_xf; g[3] = pf_dw; g[4] = pf_dr; g[5] = pf_own; g[6] = pf_skip;
        }
        const long long pf_e0 = clock64();
#endif
        float* const Cb = args.C + size_t(bz) * args.strideC;
        const size_t ldc = args.ldc;
        if (cs) {
            // ---- cluster split-K: the units of this tile are the CTAs (pairs) of this cluster.  Barrier 1: every unit has
            // left its mainloop (its last chunk is drained, so its MMAs have retired and its operand stages are idle);
            // then each thread writes the column slices it does not own into the owners' shared memory; barrier 2
            // (release / acquire at cluster scope) makes them visible; the owner adds the partials in split order and
            // finishes its slice through the whole-tile path below.
            const int S = args.splits;
            const int row_in_cta = quarter * 32 + lane;
            cluster_sync_all();
            if (S == 2) cluster_split_send<CG, TN, 2>(acc, smem_base, split, half, row_in_cta, rank);
            else cluster_split_send<CG, TN, 4>(acc, smem_base, split, half, row_in_cta, rank);
            cluster_sync_all();
            const float* const recv = reinterpret_cast<const float*>(gen_base);
            if (S == 2) cluster_split_reduce<CG, TN, 2>(acc, recv, split, half, row_in_cta);
            else cluster_split_reduce<CG, TN, 4>(acc, recv, split, half, row_in_cta);
            if (s_chain.n) epi_bar_sync();   // the program's per-warp scratch (below) overlays the receive area
        }
        if (split < 0 || cs) {
            // ---- whole tile: alpha/beta/chain on the register accumulators, coalesced column stores
            const size_t row = size_t(m0) + quarter * 32 + lane;
            const bool row_ok = row < args.m;
            const size_t ncol0 = size_t(n0) + half * HALF_N;
            const int slice_w = cs ? TILE_N / args.splits : TILE_N;   // cluster split: this unit finishes columns [split*w, (split+1)*w)
            // The general column block below (beta, broadcast stage, program, peer / multicast destinations, ragged edges)
            // is ~3.8k instructions per 32 columns, executed once per unit: ncu on a walked batch of k = 128 members
            // showed a third of all warp samples stalled on instruction fetch there.  The common case -- C = alpha * acc
            // into a full block of 32 columns -- takes a compact path of its own (~100 instructions).
            const bool plain = args.beta == 0.0f && !s_chain.bias && s_chain.n == 0 && !args.mc && args.n_peers == 0;
#pragma unroll
            for (int p = 0; p < HALF_N / 32; p++) {
                const size_t colp = ncol0 + p * 32;
                if (cs && (half * HALF_N + p * 32) / slice_w != split) continue;   // warp-uniform
                if (plain && colp + 32 <= args.n) {   // warp-uniform
                    if (row_ok) {
                        float* dst = Cb + row + colp * ldc;
#pragma unroll
                        for (int c = 0; c < 32; c++) {
                            *dst = args.alpha * acc[p * 32 + c];
                            dst += ldc;
                        }
                    }
                    continue;
                }
                if (colp < args.n) {  // warp-uniform
                    const int ncols = args.n - colp < 32 ? int(args.n - colp) : 32;
                    float v[32];
#pragma unroll
                    for (int c = 0; c < 32; c++) v[c] = args.alpha * acc[p * 32 + c];
                    if (args.beta != 0.0f && row_ok) {
                        const float* src = Cb + row + colp * ldc;  // running pointer: no 32 hoisted addresses
#pragma unroll
                        for (int c = 0; c < 32; c++) {
                            if (c < ncols) v[c] += args.beta * *src;
                            src += ldc;
                        }
                    }
                    if (s_chain.bias) {   // x = s1*x + s2*bias[row | column]: W*x + b*ones(1,N) as part of the product
                        const float br = s_chain.bias_dim == 1 ? s_chain.bias[row_ok ? row : 0] : 0.0f;
#pragma unroll
                        for (int c = 0; c < 32; c++) {
                            const float bv = s_chain.bias_dim == 1 ? br : s_chain.bias[c < ncols ? colp + c : colp];
                            v[c] = __fadd_rn(__fmul_rn(s_chain.bias_s1, v[c]), __fmul_rn(s_chain.bias_s2, bv));
                        }
                    }
                    if (s_chain.n) {
                        // Elementwise program: NOT unrolled over the 32 columns in registers -- that is ~100 KB of
                        // straight-line code per tile, executed once, and the epilogue then waits on instruction
                        // fetch (ncu: a third of all samples stalled on no_instruction, profiles/r02_config1_*).  The
                        // row goes through a per-warp 32 x 32 scratch in the (now idle) operand stages and a compact
                        // loop applies the program eight columns at a time: same per-element arithmetic, same bits.
                        float* const sw = reinterpret_cast<float*>(gen_base) + e * 1024;
                        __syncwarp();
#pragma unroll
                        for (int c = 0; c < 32; c++) sw[c * 32 + lane] = v[c];
                        __syncwarp();
#pragma unroll 1
                        for (int c0 = 0; c0 < ncols; c0 += 8) {   // 8 independent dependency chains per thread per step
                            float w[8];
#pragma unroll
                            for (int q = 0; q < 8; q++) w[q] = sw[(c0 + q) * 32 + lane];
                            apply_chain<8>(w, s_chain);
                            if (row_ok) {
                                if (args.mc) {
#pragma unroll
                                    for (int q = 0; q < 8; q++)
                                        if (c0 + q < ncols) multimem_st(args.mc + row + (colp + c0 + q) * ldc, w[q]);
                                } else {
                                    for (int d = 0; d <= args.n_peers; d++) {
                                        float* dst = (d == 0 ? Cb : s_peers[d - 1]) + row + (colp + c0) * ldc;
#pragma unroll
                                        for (int q = 0; q < 8; q++) {
                                            if (c0 + q < ncols) *dst = w[q];
                                            dst += ldc;
                                        }
                                    }
                                }
                            }
                        }
                        continue;
                    }
                    // a warp writes 32 consecutive floats (128 B) per column.  Multi-GPU (fused all-gather): either ONE
                    // multimem.st per element to the multicast image (NVSwitch replicates it into every GPU's C,
                    // this one included), or the local C plus one P2P store per peer image.
                    if (row_ok) {
                        if (args.mc) {
                            float* dst = args.mc + row + colp * ldc;
#pragma unroll
                            for (int c = 0; c < 32; c++) {
                                if (c < ncols) multimem_st(dst, v[c]);
                                dst += ldc;
                            }
                        } else {
                            for (int d = 0; d <= args.n_peers; d++) {
                                float* dst = (d == 0 ? Cb : s_peers[d - 1]) + row + colp * ldc;
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
        } else {
            // ---- k-split unit: park the raw partial tile, then finish a 1/splits column slice of the tile
            const int S = args.splits;
            const unsigned want = unsigned(S) * CG;
            float* const ws_tile = args.ws + size_t(split_tile) * size_t(S) * (CG * TILE_ELEMS);   // [split][rank][col][row]
            {
                float* dst = ws_tile + (size_t(split) * CG + rank) * TILE_ELEMS + size_t(half * HALF_N) * TILE_M + quarter * 32 + lane;
#pragma unroll
                for (int c = 0; c < HALF_N; c++) __stcg(dst + c * TILE_M, acc[c]);
            }
#ifdef JZ_GEMM_PROFILE
            const long long pf_s1 = clock64();
#endif
            __threadfence();
            epi_bar_sync();
            unsigned* const tk = args.tickets + split_tile;
#ifdef JZ_GEMM_PROFILE
            const long long pf_s2 = clock64();
#endif
            if (te == 0) {
                atomicAdd(tk, 1u);
                while (ld_acquire_gpu(tk) < want) __nanosleep(20);
            }
            epi_bar_sync();
            __threadfence();
#ifdef JZ_GEMM_PROFILE
            if (blockIdx.x == 5 && blockIdx.z == 0 && te == 0) {
                args.prof[29] = pf_s1 - pf_e0; args.prof[30] = pf_s2 - pf_s1; args.prof[31] = clock64() - pf_s2;
            }
#endif
            // slice: columns [c_begin, c_end) of this CTA's 128 rows; a thread owns 4 consecutive rows (128-bit loads of
            // the partials, which sit [column][128 rows]) and every 8th column; the S partials are added in split
            // order, UB columns (UB x S independent loads) at a time
            const int c_begin = split * TILE_N / S, c_end = (split + 1) * TILE_N / S;
            const int r4 = (te & 31) * 4, cl = te >> 5;
            const size_t row0 = size_t(m0) + r4;
            const bool vec_ok = (ldc & 3) == 0 && aligned16(Cb) && row0 + 3 < args.m;
            const float* const src0 = ws_tile + size_t(rank) * TILE_ELEMS + r4;
            const size_t sstride = size_t(CG) * TILE_ELEMS;
            constexpr int UB = 4;
            for (int cb = c_begin + cl; cb < c_end; cb += 8 * UB) {
                float4 a4[UB];
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    const int c = cb + 8 * u;
                    a4[u] = c < c_end ? __ldcg(reinterpret_cast<const float4*>(src0 + size_t(c) * TILE_M)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll 2
                for (int s = 1; s < S; s++) {
                    float4 t4[UB];
#pragma unroll
                    for (int u = 0; u < UB; u++) {
                        const int c = cb + 8 * u;
                        t4[u] = c < c_end ? __ldcg(reinterpret_cast<const float4*>(src0 + size_t(s) * sstride + size_t(c) * TILE_M))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < UB; u++) {
                        a4[u].x = __fadd_rn(a4[u].x, t4[u].x); a4[u].y = __fadd_rn(a4[u].y, t4[u].y);
                        a4[u].z = __fadd_rn(a4[u].z, t4[u].z); a4[u].w = __fadd_rn(a4[u].w, t4[u].w);
                    }
                }
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    const int c = cb + 8 * u;
                    const size_t col = size_t(n0) + c;
                    if (c >= c_end || col >= args.n || row0 >= args.m) continue;
                    float v[4] = {args.alpha * a4[u].x, args.alpha * a4[u].y, args.alpha * a4[u].z, args.alpha * a4[u].w};
                    float* const dstc = Cb + row0 + col * ldc;
                    if (args.beta != 0.0f) {
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            if (row0 + q < args.m) v[q] += args.beta * dstc[q];
                    }
                    if (s_chain.bias) {
#pragma unroll
                        for (int q = 0; q < 4; q++) v[q] = apply_bias(v[q], s_chain, row0 + q < args.m ? row0 + q : row0, col);
                    }
                    if (s_chain.n) apply_chain<4>(v, s_chain);
                    if (args.mc) {
                        float* const dm = args.mc + row0 + col * ldc;
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            if (row0 + q < args.m) multimem_st(dm + q, v[q]);
                    } else {
                        for (int d = 0; d <= args.n_peers; d++) {
                            float* const dd = (d == 0 ? Cb : s_peers[d - 1]) + row0 + col * ldc;
                            if (vec_ok) {
                                *reinterpret_cast<float4*>(dd) = make_float4(v[0], v[1], v[2], v[3]);
                            } else {
#pragma unroll
                                for (int q = 0; q < 4; q++)
                                    if (row0 + q < args.m) dd[q] = v[q];
                            }
                        }
                    }
                }
            }
            // departure: the last CTA of the tile to get here leaves both counters at zero for the next launch
            epi_bar_sync();
            if (te == 0) {
                unsigned* const dn = args.tickets + (kTicketSlots / 2) + split_tile;
                if (atomicAdd(dn, 1u) == want - 1) {
                    *tk = 0;
                    *dn = 0;
                }
            }
        }
#ifdef JZ_GEMM_PROFILE
        if (blockIdx.x == 5 && blockIdx.z == 0 && lane == 0 && e == 0)
            { args.prof[24] = clock64() - pf_e0; args.prof[25] = split; args.prof[26] = args.splits; }
#endif
        if (!walk || (u += ustride) >= walk) break;
        pl = place(u);              // next unit of the walk; its first k-blocks are usually in shared memory already
        chunk_base += num_chunks;
        }
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc<CG>(tmem_base, TMEM_COLS);
}

// ======================================================================= persistent TF32 kernel
// Single-pass TF32, CTA pair per 256 x 256 tile, PERSISTENT: a pair walks the units pair, pair + P, pair + 2P ... (P pairs
// resident) and the accumulator of unit i + 1 fills the other TMEM half while the epilogue warps drain, finish and
// store unit i.  ncu on the one-tile-per-pair kernel shows why: at 4096^3 the tensor pipe is active only 69 % of the
// time in TF32 mode, the rest is the per-tile prologue (TMEM allocation, barrier init, first TMA round trip) and the
// epilogue (accumulator drain + 256 KB of stores), ~12 us per wave that nothing hides (profiles/r02c_ncu_gemm_summary.txt).
// Here the prologue is paid once per launch and the epilogue runs under the next unit's mainloop.
//   * the whole k range of a unit is chained in TMEM (no register-level promotion: TF32 mode is input-rounding
//     dominated, 7e-4; the chain's truncation adds < 1.2e-4 at k = 16384);
//   * epilogue: 4 x tcgen05.ld (128 columns per warp) into registers, TMEM half released at once, then alpha / beta /
//     bias / program / stores exactly as in gemm_tcgen05_kernel (program through a per-warp scratch that this kernel
//     owns, the operand stages being busy with the next unit);
//   * units are the same list as in the one-shot kernel (whole tiles, then k-splits of the last partial wave); the
//     units of a split tile meet through workspace and a ticket WITHOUT waiting (the last one to arrive finishes the
//     tile); pairs stay in step because every pair runs at the tensor pipe's rate, so the tiles of a "wave" still
//     share their A / B panels in L2.
constexpr int P_STAGES = 6;                                   // 6 x 32 KB operand stages
constexpr int P_STAGE_BYTES = A_BYTES + (256 / 2) * BK * 4;   // 32 KB: 128 rows of A + 128 rows of B per CTA
constexpr int P_SCRATCH_BYTES = NUM_EPI_WARPS * 4096;         // per-warp 32 x 32 scratch for the elementwise program
constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + P_SCRATCH_BYTES + 1024 + 256;

struct UnitInfo {
    int m0, n0, nb0, kb0, num_kb, split;
    unsigned split_tile;
};
__device__ __forceinline__ UnitInfo decode_unit(const GemmArgs& args, unsigned unit, uint32_t rank, int tile_n) {
    UnitInfo u;
    const int num_kb_total = int((args.k + BK - 1) / BK);
    unsigned tile = unit;
    u.kb0 = 0; u.num_kb = num_kb_total; u.split = -1; u.split_tile = 0;
    if (unit >= args.full_tiles) {
        const unsigned r = unit - args.full_tiles;
        u.split_tile = r / unsigned(args.splits);
        u.split = int(r % unsigned(args.splits));
        tile = args.full_tiles + u.split_tile;
        u.kb0 = u.split * args.kb_per_split;
        u.num_kb = num_kb_total - u.kb0 < args.kb_per_split ? num_kb_total - u.kb0 : args.kb_per_split;
    }
    constexpr unsigned GROUP = 8;
    const unsigned per_group = GROUP * args.tiles_n;
    const unsigned group_id = tile / per_group;
    const unsigned first_m = group_id * GROUP;
    const unsigned group_m = args.tiles_m - first_m < GROUP ? args.tiles_m - first_m : GROUP;
    const unsigned tm = first_m + (tile % per_group) % group_m;
    const unsigned tn = (tile % per_group) / group_m;
    u.m0 = int(tm) * (2 * TILE_M) + int(rank) * TILE_M;
    u.n0 = int(tn) * tile_n;
    u.nb0 = u.n0 + int(rank) * (tile_n / 2);
    return u;
}

template <bool AMN, bool BMN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tf32_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs args,
                            const unsigned total_units) {
    constexpr int CG = 2, TILE_N = 256, HALF_N = 128, STAGES = P_STAGES, STAGE_BYTES = P_STAGE_BYTES;
    constexpr int B_ROWS = 128;
    constexpr uint32_t IDESC = make_idesc_tf32(256, TILE_N, AMN, BMN);
    constexpr uint32_t KA = kstep_bytes<AMN>(), KB = kstep_bytes<BMN>();
    constexpr int TILE_ELEMS = TILE_M * TILE_N;

    extern __shared__ uint8_t smem_raw[];
    __shared__ ChainParams s_chain;
    __shared__ float* s_peers[JZ_MAX_PEERS];
    __shared__ bool s_last;   // this CTA was the last of a split tile's units to arrive
    stage_chain(&s_chain, args.chain, threadIdx.x);
    if (threadIdx.x == 32) {
#pragma unroll
        for (int q = 0; q < JZ_MAX_PEERS; q++) s_peers[q] = args.peers[q];
    }
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES + P_SCRATCH_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tmem_full_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
    auto tmem_empty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + STAGES * STAGE_BYTES + P_SCRATCH_BYTES + 8 * (2 * STAGES + 4));

    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const unsigned pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    if (warp == 0 && elect_one()) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1) {
        if (elect_one()) {
            for (int s = 0; s < STAGES; s++) {
                mbar_init(full_bar(s), 1);
                mbar_init(empty_bar(s), 1);
            }
            for (int b = 0; b < 2; b++) {
                mbar_init(tmem_full_bar(b), 1);
                mbar_init(tmem_empty_bar(b), NUM_EPI_WARPS * CG);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<CG>(tmem_slot, 2 * TILE_N);
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 0) {
        // ===================== TMA producer: one continuous k-block stream over this pair's units =====================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (unsigned unit = pair; unit < total_units; unit += num_pairs) {
                const UnitInfo u = decode_unit(args, unit, rank, TILE_N);
                for (int kb = 0; kb < u.num_kb; kb++) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    if (leader) mbar_arrive_expect_tx(full_bar(stage), uint32_t(STAGE_BYTES) * CG);
                    const uint32_t sa = smem_base + stage * STAGE_BYTES;
                    const int kc = (u.kb0 + kb) * BK;
                    tma_load_tile<AMN, TILE_M, true>(sa, &tmA, full_bar(stage), kc, u.m0, 0);
                    tma_load_tile<BMN, B_ROWS, true>(sa + A_BYTES, &tmB, full_bar(stage), kc, u.nb0, 0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA): unit i accumulates in TMEM half i & 1 =====================
        if (leader) {
            uint32_t stage = 0, phase = 0, it = 0;
            for (unsigned unit = pair; unit < total_units; unit += num_pairs, it++) {
                const UnitInfo u = decode_unit(args, unit, rank, TILE_N);
                const uint32_t buf = it & 1u;
                mbar_wait(tmem_empty_bar(buf), ((it >> 1) & 1u) ^ 1u);   // the epilogue warps have read this half out
                tc_fence_after();
                for (int kb = 0; kb < u.num_kb; kb++) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t d = tmem_base + buf * TILE_N;
                        const uint32_t sa = smem_base + stage * STAGE_BYTES;
#pragma unroll
                        for (int ks = 0; ks < BK / UMMA_K; ks++)
                            umma_tf32<CG>(d, make_smem_desc<AMN>(sa + ks * KA), make_smem_desc<BMN>(sa + A_BYTES + ks * KB), IDESC,
                                          (kb | ks) ? 1u : 0u);
                        umma_commit<CG>(empty_bar(stage));
                        if (kb == u.num_kb - 1) umma_commit<CG>(tmem_full_bar(buf));
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue warps: drain a finished unit while the next one is being multiplied =====================
        const int e = warp - FIRST_EPI_WARP;
        const int quarter = warp & 3;
        const int half = e >> 2;
        const int lane = threadIdx.x & 31;
        const int te = threadIdx.x - 32 * FIRST_EPI_WARP;
        const uint32_t lane_addr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(half * HALF_N);
        const uint32_t empty0 = tmem_empty_bar(0) & 0xFEFFFFFFu, empty1 = tmem_empty_bar(1) & 0xFEFFFFFFu;
        float* const sw = reinterpret_cast<float*>(gen_base + STAGES * STAGE_BYTES) + e * 1024;
        const size_t ldc = args.ldc;
        float* const Cb = args.C;
        uint32_t it = 0;
        for (unsigned unit = pair; unit < total_units; unit += num_pairs, it++) {
            const UnitInfo u = decode_unit(args, unit, rank, TILE_N);
            const uint32_t buf = it & 1u;
            mbar_wait(tmem_full_bar(buf), (it >> 1) & 1u);
            tc_fence_after();
            float acc[HALF_N];
#pragma unroll
            for (int p = 0; p < HALF_N / 32; p++) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(lane_addr + buf * TILE_N + p * 32, r);
#pragma unroll
                for (int c = 0; c < 32; c++) acc[p * 32 + c] = __uint_as_float(r[c]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(buf ? empty1 : empty0);   // this TMEM half may be overwritten now
            if (u.split < 0) {
                const size_t row = size_t(u.m0) + quarter * 32 + lane;
                const bool row_ok = row < args.m;
                const size_t ncol0 = size_t(u.n0) + half * HALF_N;
                // compact path for C = alpha * acc into a full block of 32 columns (see gemm_tcgen05_kernel: the general
                // block is thousands of instructions executed once per unit, and instruction fetch shows in the samples)
                const bool plain = args.beta == 0.0f && !s_chain.bias && s_chain.n == 0 && !args.mc && args.n_peers == 0;
#pragma unroll
                for (int p = 0; p < HALF_N / 32; p++) {
                    const size_t colp = ncol0 + p * 32;
                    if (plain && colp + 32 <= args.n) {   // warp-uniform
                        if (row_ok) {
                            float* dst = Cb + row + colp * ldc;
#pragma unroll
                            for (int c = 0; c < 32; c++) {
                                *dst = args.alpha * acc[p * 32 + c];
                                dst += ldc;
                            }
                        }
                        continue;
                    }
                    if (colp < args.n) {  // warp-uniform
                        const int ncols = args.n - colp < 32 ? int(args.n - colp) : 32;
                        float v[32];
#pragma unroll
                        for (int c = 0; c < 32; c++) v[c] = args.alpha * acc[p * 32 + c];
                        if (args.beta != 0.0f && row_ok) {
                            const float* src = Cb + row + colp * ldc;
#pragma unroll
                            for (int c = 0; c < 32; c++) {
                                if (c < ncols) v[c] += args.beta * *src;
                                src += ldc;
                            }
                        }
                        if (s_chain.bias) {
                            const float br = s_chain.bias_dim == 1 ? s_chain.bias[row_ok ? row : 0] : 0.0f;
#pragma unroll
                            for (int c = 0; c < 32; c++) {
                                const float bv = s_chain.bias_dim == 1 ? br : s_chain.bias[c < ncols ? colp + c : colp];
                                v[c] = __fadd_rn(__fmul_rn(s_chain.bias_s1, v[c]), __fmul_rn(s_chain.bias_s2, bv));
                            }
                        }
                        if (s_chain.n) {
                            __syncwarp();
#pragma unroll
                            for (int c = 0; c < 32; c++) sw[c * 32 + lane] = v[c];
                            __syncwarp();
#pragma unroll 1
                            for (int c0 = 0; c0 < ncols; c0 += 8) {
                                float w[8];
#pragma unroll
                                for (int q = 0; q < 8; q++) w[q] = sw[(c0 + q) * 32 + lane];
                                apply_chain<8>(w, s_chain);
                                if (row_ok) {
                                    if (args.mc) {
#pragma unroll
                                        for (int q = 0; q < 8; q++)
                                            if (c0 + q < ncols) multimem_st(args.mc + row + (colp + c0 + q) * ldc, w[q]);
                                    } else {
                                        for (int d = 0; d <= args.n_peers; d++) {
                                            float* dst = (d == 0 ? Cb : s_peers[d - 1]) + row + (colp + c0) * ldc;
#pragma unroll
                                            for (int q = 0; q < 8; q++) {
                                                if (c0 + q < ncols) *dst = w[q];
                                                dst += ldc;
                                            }
                                        }
                                    }
                                }
                            }
                            continue;
                        }
                        if (row_ok) {
                            if (args.mc) {
                                float* dst = args.mc + row + colp * ldc;
#pragma unroll
                                for (int c = 0; c < 32; c++) {
                                    if (c < ncols) multimem_st(dst, v[c]);
                                    dst += ldc;
                                }
                            } else {
                                for (int d = 0; d <= args.n_peers; d++) {
                                    float* dst = (d == 0 ? Cb : s_peers[d - 1]) + row + colp * ldc;
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
            } else {
                // ---- k-split unit: park the partial and take a ticket; the LAST of the S units of this tile (per CTA rank)
                // to arrive adds the S partials in split order and finishes the tile's rows of its rank.  Nobody waits:
                // a persistent pair owns units that may not be resident yet when another kernel shares the GPU, so
                // the waiting protocol of the one-shot kernel could stall here; the longer fix-up of the last arriver
                // runs under the next unit's mainloop like any other epilogue.
                const int S = args.splits;
                float* const ws_tile = args.ws + size_t(u.split_tile) * size_t(S) * (CG * TILE_ELEMS);
                {
                    float* dst = ws_tile + (size_t(u.split) * CG + rank) * TILE_ELEMS + size_t(half * HALF_N) * TILE_M + quarter * 32 + lane;
#pragma unroll
                    for (int c = 0; c < HALF_N; c++) __stcg(dst + c * TILE_M, acc[c]);
                }
                __threadfence();
                epi_bar_sync();
                unsigned* const tk = args.tickets + u.split_tile * CG + rank;
                if (te == 0) s_last = atomicAdd(tk, 1u) == unsigned(S) - 1u;
                epi_bar_sync();
                if (s_last) {
                    __threadfence();
                    const int r4 = (te & 31) * 4, cl = te >> 5;
                    const size_t row0 = size_t(u.m0) + r4;
                    const bool vec_ok = (ldc & 3) == 0 && aligned16(Cb) && row0 + 3 < args.m;
                    const float* const src0 = ws_tile + size_t(rank) * TILE_ELEMS + r4;
                    const size_t sstride = size_t(CG) * TILE_ELEMS;
                    constexpr int UB = 4;
                    for (int cb = cl; cb < TILE_N; cb += 8 * UB) {
                        float4 a4[UB];
#pragma unroll
                        for (int q = 0; q < UB; q++) a4[q] = __ldcg(reinterpret_cast<const float4*>(src0 + size_t(cb + 8 * q) * TILE_M));
#pragma unroll 4
                        for (int s = 1; s < S; s++) {
                            float4 t4[UB];
#pragma unroll
                            for (int q = 0; q < UB; q++)
                                t4[q] = __ldcg(reinterpret_cast<const float4*>(src0 + size_t(s) * sstride + size_t(cb + 8 * q) * TILE_M));
#pragma unroll
                            for (int q = 0; q < UB; q++) {
                                a4[q].x = __fadd_rn(a4[q].x, t4[q].x); a4[q].y = __fadd_rn(a4[q].y, t4[q].y);
                                a4[q].z = __fadd_rn(a4[q].z, t4[q].z); a4[q].w = __fadd_rn(a4[q].w, t4[q].w);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < UB; q++) {
                            const size_t col = size_t(u.n0) + cb + 8 * q;
                            if (col >= args.n || row0 >= args.m) continue;
                            float v[4] = {args.alpha * a4[q].x, args.alpha * a4[q].y, args.alpha * a4[q].z, args.alpha * a4[q].w};
                            float* const dstc = Cb + row0 + col * ldc;
                            if (args.beta != 0.0f) {
#pragma unroll
                                for (int t = 0; t < 4; t++)
                                    if (row0 + t < args.m) v[t] += args.beta * dstc[t];
                            }
                            if (s_chain.bias) {
#pragma unroll
                                for (int t = 0; t < 4; t++) v[t] = apply_bias(v[t], s_chain, row0 + t < args.m ? row0 + t : row0, col);
                            }
                            if (s_chain.n) apply_chain<4>(v, s_chain);
                            if (args.mc) {
#pragma unroll
                                for (int t = 0; t < 4; t++)
                                    if (row0 + t < args.m) multimem_st(args.mc + row0 + t + col * ldc, v[t]);
                            } else {
                                for (int d = 0; d <= args.n_peers; d++) {
                                    float* const dd = (d == 0 ? Cb : s_peers[d - 1]) + row0 + col * ldc;
                                    if (vec_ok) {
                                        *reinterpret_cast<float4*>(dd) = make_float4(v[0], v[1], v[2], v[3]);
                                    } else {
#pragma unroll
                                        for (int t = 0; t < 4; t++)
                                            if (row0 + t < args.m) dd[t] = v[t];
                                    }
                                }
                            }
                        }
                    }
                    if (te == 0) *tk = 0;   // ready for the next launch on this stream
                }
                epi_bar_sync();   // s_last is rewritten by the next split unit
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    if (warp == 1) tmem_dealloc<CG>(tmem_base, 2 * TILE_N);
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

// Tensor map of one operand over its source storage, zero OOB fill; third dimension = batch member.
//   K-major ([rows][k], row stride `stride_elems`): dims {k, rows, batch}, box {32, box_rows, 1}, 128B swizzle.
//   MN-major ([k][rows], k stride `stride_elems`): dims {rows, k, batch}, box {32 rows, 32 k, 1}, 128B swizzle / 32B atoms.
// A training loop multiplies the same buffers step after step (the pool hands a temporary its old block back), and the
// driver call costs about a microsecond of host time per operand on a launch-bound step: the encoded maps are kept in a
// small per-thread direct-mapped cache keyed on everything the encoding depends on (the map holds no device state).
struct MapKey {
    const float* ptr; size_t rows, k, stride, batch_stride; int box_rows; unsigned batch; bool mn;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && rows == o.rows && k == o.k && stride == o.stride && batch_stride == o.batch_stride &&
               box_rows == o.box_rows && batch == o.batch && mn == o.mn;
    }
};
struct MapSlot { MapKey key; bool valid; alignas(64) CUtensorMap map; };
static int make_map_uncached(CUtensorMap* map, const Operand& op, size_t rows, size_t k, int box_rows, unsigned batch);
static int make_map(CUtensorMap* map, const Operand& op, size_t rows, size_t k, int box_rows, unsigned batch) {
    constexpr size_t SLOTS = 64;
    static thread_local MapSlot cache[SLOTS];
    const MapKey key{op.ptr, rows, k, op.stride, op.batch_stride, box_rows, batch, op.mn};
    size_t h = reinterpret_cast<uintptr_t>(op.ptr) >> 8;
    h ^= (rows * 0x9E3779B97F4A7C15ull) ^ (k * 0xC2B2AE3D27D4EB4Full) ^ (size_t(box_rows) << 20) ^ (size_t(op.mn) << 40);
    MapSlot& slot = cache[(h ^ (h >> 17) ^ (h >> 31)) % SLOTS];
    if (slot.valid && slot.key == key) {
        std::memcpy(map, &slot.map, sizeof(CUtensorMap));
        return JZ_OK;
    }
    const int rc = make_map_uncached(map, op, rows, k, box_rows, batch);
    if (rc == JZ_OK) {
        slot.key = key;
        std::memcpy(&slot.map, map, sizeof(CUtensorMap));
        slot.valid = true;
    }
    return rc;
}
static int make_map_uncached(CUtensorMap* map, const Operand& op, size_t rows, size_t k, int box_rows, unsigned batch) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(JZ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const bool mn = op.mn;
    cuuint64_t dims[3] = {cuuint64_t(mn ? rows : k), cuuint64_t(mn ? k : rows), cuuint64_t(batch)};
    // a single product has no batch stride: use the natural one (dim1 * stride0), it is never stepped
    size_t bstride = batch > 1 ? op.batch_stride : size_t(dims[1]) * op.stride;
    if (bstride * sizeof(float) >= (size_t(1) << 40)) bstride = 4;
    cuuint64_t strides[2] = {cuuint64_t(op.stride) * sizeof(float), cuuint64_t(bstride) * sizeof(float)};
    cuuint32_t box[3] = {cuuint32_t(BK), cuuint32_t(mn ? BK : box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(op.ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(JZ_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
    return JZ_OK;
}

static bool pdl_enabled() { return pdl_on(); }

// args.tiles_m / tiles_n / full_tiles / splits / kb_per_split / ws / tickets are filled by the caller (plan_units)
template <int CG, int MODE, int TN, bool AMN, bool BMN, bool WALK>
static int launch_tc(const Operand& a, const Operand& b, const GemmArgs& args, unsigned batch, cudaStream_t s) {
    alignas(64) CUtensorMap ma, mb;
    int rc;
    if ((rc = make_map(&ma, a, args.m, args.k, TILE_M, batch)) != JZ_OK) return rc;
    if ((rc = make_map(&mb, b, args.n, args.k, b_rows<CG, TN>(), batch)) != JZ_OK) return rc;
    auto kern = gemm_tcgen05_kernel<CG, MODE, TN, AMN, BMN, WALK>;
    constexpr int SMEM = smem_bytes<CG, MODE, TN>();
    static bool attr_done = false;
    if (!attr_done) {
        JZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_done = true;
    }
    const unsigned tiles = args.tiles_m * args.tiles_n;
    const unsigned units = args.full_tiles + (tiles - args.full_tiles) * unsigned(args.splits);
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(units * CG, 1, batch);
    if (args.walk_units) {   // one CTA group per SM (pair) walks the (member, tile) units: see gemm_tc
        const unsigned resident = unsigned(ctx().sm_count) / CG;
        cfg.gridDim = dim3((args.walk_units < resident ? args.walk_units : resident) * CG, 1, 1);
    }
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    GemmArgs local = args;
    if (args.cluster_split) {
        // the units of a tile as ONE cluster of CG * splits CTAs -- only while every tile's cluster is resident at once (a
        // cluster of 8 full-SM CTAs needs 8 free SMs of one GPC); otherwise the workspace form, which the caller prepared
        const unsigned csize = unsigned(CG * args.splits);
        static int max_clusters[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // per cluster size; -1 = cannot launch
        if (csize <= 8 && max_clusters[csize] == 0) {
            attr[0].val.clusterDim.x = csize;
            cfg.numAttrs = 1;
            cfg.gridDim = dim3(csize, 1, 1);
            int nc = 0;
            if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess) { cudaGetLastError(); nc = 0; }
            max_clusters[csize] = nc > 0 ? nc : -1;
            if (std::getenv("JZ_GEMM_DEBUG")) std::fprintf(stderr, "[jz_gemm] CG=%d TN=%d MODE=%d: at most %d resident clusters of %u CTAs\n", CG, TN, MODE, nc, csize);
            cfg.gridDim = dim3(units * CG, 1, batch);
        }
        if (csize <= 8 && args.full_tiles == 0 && int(tiles) <= max_clusters[csize]) attr[0].val.clusterDim.x = csize;
        else { attr[0].val.clusterDim.x = CG; local.cluster_split = 0; }
    }
    ctx().gemm_last_cluster_split = local.cluster_split;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ma, mb, local);
    ctx().launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return cuda_fail(e, "gemm_tcgen05_kernel launch");
    return JZ_OK;
}

// persistent TF32 kernel: grid = min(units, resident pairs) CTA pairs
template <bool AMN, bool BMN>
static int launch_tf32_persistent(const Operand& a, const Operand& b, const GemmArgs& args, cudaStream_t s) {
    alignas(64) CUtensorMap ma, mb;
    int rc;
    if ((rc = make_map(&ma, a, args.m, args.k, TILE_M, 1)) != JZ_OK) return rc;
    if ((rc = make_map(&mb, b, args.n, args.k, 128, 1)) != JZ_OK) return rc;
    auto kern = gemm_tf32_persistent_kernel<AMN, BMN>;
    static bool attr_done = false;
    if (!attr_done) {
        JZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES));
        attr_done = true;
    }
    const unsigned tiles = args.tiles_m * args.tiles_n;
    const unsigned units = args.full_tiles + (tiles - args.full_tiles) * unsigned(args.splits);
    const unsigned pairs = unsigned(ctx().sm_count) / 2;
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((units < pairs ? units : pairs) * 2, 1, 1);
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = P_SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ma, mb, args, units);
    ctx().launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return cuda_fail(e, "gemm_tf32_persistent_kernel launch");
    return JZ_OK;
}
#ifdef JZ_GEMM_TC_PERSIST_IMPL   // only jz_gemm_tc_tf32_persist.cu instantiates the persistent kernels
static int launch_tf32_persistent_major(const Operand& a, const Operand& b, const GemmArgs& args, cudaStream_t s) {
    if (a.mn) return b.mn ? launch_tf32_persistent<true, true>(a, b, args, s) : launch_tf32_persistent<true, false>(a, b, args, s);
    return b.mn ? launch_tf32_persistent<false, true>(a, b, args, s) : launch_tf32_persistent<false, false>(a, b, args, s);
}
#endif

// operand majors are compile-time (they select TMA box shapes and descriptor layouts)
template <int CG, int MODE, int TN, bool WALK>
static int launch_tc_major_w(const Operand& a, const Operand& b, const GemmArgs& args, unsigned batch, cudaStream_t s) {
    if (a.mn) return b.mn ? launch_tc<CG, MODE, TN, true, true, WALK>(a, b, args, batch, s) : launch_tc<CG, MODE, TN, true, false, WALK>(a, b, args, batch, s);
    return b.mn ? launch_tc<CG, MODE, TN, false, true, WALK>(a, b, args, batch, s) : launch_tc<CG, MODE, TN, false, false, WALK>(a, b, args, batch, s);
}
// WALKABLE: this (CG, MODE, TN) also exists as the unit-walking variant used for strided batches (gemm_tc asks for a walk
// only on the variants instantiated with it, see walkable())
template <int CG, int MODE, int TN, bool WALKABLE = true>
static int launch_tc_major(const Operand& a, const Operand& b, const GemmArgs& args, unsigned batch, cudaStream_t s) {
    if constexpr (WALKABLE) {
        if (args.walk_units) return launch_tc_major_w<CG, MODE, TN, true>(a, b, args, batch, s);
    } else {
        if (args.walk_units) return fail(JZ_ERR_ARG, "gemm: this kernel variant has no unit-walking form");
    }
    return launch_tc_major_w<CG, MODE, TN, false>(a, b, args, batch, s);
}

#endif  // JZ_GEMM_TC_IMPL

}  // namespace tc
}  // namespace jz
