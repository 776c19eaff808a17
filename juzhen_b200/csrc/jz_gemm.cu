// jz_gemm.cu -- dispatch of the GEMM behind Matrix<CUDAfloat>::dot / operator* (SURVEY 8a row a17):
//     C(m x n, ldc) = alpha * op(A)(m x k) * op(B)(k x n) + beta * C        (column-major)
// replacing cublasSgemm (cpp/cumatrix.cu:177-197) and cublasSgemmStridedBatched (ml/layer.hpp:2896-2926).
// No cuBLAS anywhere.
//
// Main path (sm_100a): the TMA + tcgen05.mma.kind::tf32 + TMEM kernel of jz_gemm_tc.cuh (3xTF32 for fp32 accuracy by
// default, single-pass TF32 with NVIDIA_TF32=1).  This file picks the tile shape (CTA pair 256 x {256,128} or single
// CTA 128 x {256,128,64}), plans the units of the launch (whole tiles, then k-splits of the partial last wave so the
// tail fills every SM; small and skinny products become all-split launches), and hands strided batches to the same
// kernel with the member on blockIdx.z.
//
// Other paths: shapes the tensor path does not take (m or n < 64, k < 32, tiny) go to the latency-oriented
// warp-per-tile fp32 kernel of jz_gemm_small.cu when the product is at most 2^26 multiply-adds (a training step at
// batch 32), else to a bounds-checked shared-memory fp32 FMA (SIMT) kernel, which also serves mode JZ_GEMM_FP32_SIMT;
// rank-1 products (k == 1: the reference's broadcast idiom) are a streaming outer-product kernel.
#include <cuda.h>

#include <cmath>
#include <cstring>
#include <mutex>

#include "jz_common.cuh"
#include "jz_gemm_tc.cuh"
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
                                                         size_t ldc, ChainParams chain_p, unsigned gx) {
    pdl_enter();
    __shared__ ChainParams chain;
    stage_chain(&chain, chain_p, threadIdx.x);
    __syncthreads();
    __shared__ __align__(16) float As[SBK][SBM + 4];  // read back as float4
    __shared__ __align__(16) float Bs[SBK][SBN + 4];
    const int t = threadIdx.x;
    // linear block index = row block + gx * column block: neither dimension is bound by the 65535 limit of grid.y
    const size_t i0 = size_t(blockIdx.x % gx) * SBM, j0 = size_t(blockIdx.x / gx) * SBN;
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
            if (chain.bias) x = apply_bias(x, chain, gi < m ? gi : 0, gj);
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
    if (gx * gy >= (size_t(1) << 31)) return fail(JZ_ERR_UNSUPPORTED, "simt gemm: more than 2^31 output tiles");
    const unsigned grid = unsigned(gx * gy), gxu = unsigned(gx);
    if (!ta && !tb) JZ_LAUNCH((sgemm_simt_kernel<false, false>), grid, 256, 0, s, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain, gxu);
    else if (ta && !tb) JZ_LAUNCH((sgemm_simt_kernel<true, false>), grid, 256, 0, s, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain, gxu);
    else if (!ta && tb) JZ_LAUNCH((sgemm_simt_kernel<false, true>), grid, 256, 0, s, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain, gxu);
    else JZ_LAUNCH((sgemm_simt_kernel<true, true>), grid, 256, 0, s, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain, gxu);
    ctx().gemm_last_path = 2;
    return JZ_OK;
}

// rank-1: C = alpha*u*v^T (+beta*C) ; u = op(A)(:,0) (stride su), v = op(B)(0,:) (stride sv).
// u == nullptr means the product term is absent (k == 0).
__global__ void __launch_bounds__(256) rank1_kernel(size_t m, size_t n, float alpha, const float* u, size_t su,
                                                    const float* v, size_t sv, float beta, float* C, size_t ldc,
                                                    ChainParams chain_p) {
    pdl_enter();
    __shared__ ChainParams chain;
    stage_chain(&chain, chain_p, threadIdx.x);
    __syncthreads();
    for (size_t j = blockIdx.y; j < n; j += gridDim.y) {
        const float vj = u ? alpha * v[j * sv] : 0.0f;
        for (size_t i = size_t(blockIdx.x) * 256 + threadIdx.x; i < m; i += size_t(gridDim.x) * 256) {
            float x[1] = {u ? u[i * su] * vj : 0.0f};
            if (beta != 0.0f) x[0] += beta * C[j * ldc + i];
            if (chain.bias) x[0] = apply_bias(x[0], chain, i, j);
            if (chain.n) apply_chain<1>(x, chain);
            C[j * ldc + i] = x[0];
        }
    }
}

// ======================================================================= tensor-core path: host side
namespace tc {

// K-major copy of an operand TMA cannot address in place (base or row pitch not 16-byte aligned).
// source already K-major: src(r, kk) at r*ld + kk ; dst[r*kp + kk], one thread per quad of k, grid-stride.
__global__ void __launch_bounds__(256) prep_kmajor_kernel(float* dst, size_t kp, const float* src, size_t ld, size_t rows, size_t k) {
    pdl_enter();
    const size_t kq = kp >> 2;  // quads per row (kp is a multiple of 4)
    const size_t total = rows * kq;
    for (size_t idx = size_t(blockIdx.x) * 256 + threadIdx.x; idx < total; idx += size_t(gridDim.x) * 256) {
        const size_t r = idx / kq, kk = (idx - r * kq) << 2;
        float x[4];
#pragma unroll
        for (int i = 0; i < 4; i++) x[i] = kk + i < k ? src[r * ld + kk + i] : 0.0f;
        *reinterpret_cast<float4*>(dst + r * kp + kk) = make_float4(x[0], x[1], x[2], x[3]);
    }
}

// source MN-major: src(r, kk) at kk*ld + r ; dst[r*kp + kk]   (tiled transpose, 64 x 64 tiles through shared memory)
__global__ void __launch_bounds__(256) prep_transpose_kernel(float* dst, size_t kp, const float* src, size_t ld, size_t rows,
                                                             size_t k, size_t tiles_r, size_t tiles_k) {
    pdl_enter();
    __shared__ float tile[64][65];  // tile[kk][r]
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const size_t ntiles = tiles_r * tiles_k;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const size_t tr = t % tiles_r, tk = t / tiles_r;
        const size_t r0 = tr * 64, k0 = tk * 64;
        for (int e = threadIdx.x; e < 64 * 64; e += 256) {
            const int kl = e >> 6, rl = e & 63;
            tile[kl][rl] = (k0 + kl < k && r0 + rl < rows) ? src[(k0 + kl) * ld + r0 + rl] : 0.0f;
        }
        __syncthreads();
        // write: row r of the image, 4 consecutive k per thread (k0 and kp are multiples of 4: always in bounds)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int rl = ty + 16 * q;
            const size_t r = r0 + rl, kk = k0 + 4 * tx;
            if (r < rows && kk < kp)
                *reinterpret_cast<float4*>(dst + r * kp + kk) =
                    make_float4(tile[4 * tx + 0][rl], tile[4 * tx + 1][rl], tile[4 * tx + 2][rl], tile[4 * tx + 3][rl]);
        }
        __syncthreads();
    }
}

static bool env_flag(const char* name) {
    const char* e = std::getenv(name);
    return e && e[0] && e[0] != '0';
}
static int env_int(const char* name) {
    const char* e = std::getenv(name);
    return e && *e ? std::atoi(e) : 0;
}

// Describe an operand for the kernel.  kmajor_src: src(r,kk) at r*ld+kk, else at kk*ld+r.
// In place (no copy, either major) whenever TMA can address the source: 16-byte aligned base, row pitch and batch
// stride.  Otherwise (single products only) a pre-pass writes a K-major copy to workspace.
static int prepare_operand(Operand& op, const float* src, size_t ld, bool kmajor_src, size_t rows, size_t k, unsigned batch,
                           size_t batch_stride, cudaStream_t s) {
    static const bool no_mn = env_flag("JZ_GEMM_NO_MN");   // debugging: never consume MN-major sources in place
    const bool tma_ok = ld % 4 == 0 && aligned16(src) && (batch <= 1 || batch_stride % 4 == 0);
    if (tma_ok && (kmajor_src || !no_mn || batch > 1)) {
        op.ptr = src;
        op.stride = ld;
        op.batch_stride = batch_stride;
        op.mn = !kmajor_src;
        return JZ_OK;
    }
    if (batch > 1) return fail(JZ_ERR_UNSUPPORTED, "batched tensor-core gemm needs 16-byte aligned operands");
    const size_t kp = (k + 3) & ~size_t(3);
    int rc = ws_alloc(&op.owned, rows * kp * sizeof(float), s);
    if (rc != JZ_OK) return rc;
    float* dst = static_cast<float*>(op.owned);
    op.ptr = dst;
    op.stride = kp;
    op.batch_stride = 0;
    op.mn = false;
    const size_t cap = size_t(ctx().sm_count) * 8;
    if (kmajor_src) {
        const size_t blocks = ceil_div(rows * (kp >> 2), size_t(256));
        const unsigned grid = unsigned(blocks < cap * 2 ? (blocks ? blocks : 1) : cap * 2);
        JZ_LAUNCH(prep_kmajor_kernel, grid, 256, 0, s, dst, kp, src, ld, rows, k);
    } else {
        const size_t tiles_r = ceil_div(rows, 64), tiles_k = ceil_div(k, 64);
        const size_t nt = tiles_r * tiles_k;
        const unsigned grid = unsigned(nt < cap ? (nt ? nt : 1) : cap);
        JZ_LAUNCH(prep_transpose_kernel, grid, 256, 0, s, dst, kp, src, ld, rows, k, tiles_r, tiles_k);
    }
    return JZ_OK;
}

// Tile shape.  A CTA pair owns a 256 x TN tile, a single CTA a 128 x TN one.  256 x 256 pair tiles are preferred:
// per k-block the A tile is loaded (and, in 3xTF32, transformed) once per tile whatever its width, and narrower
// tiles are shared-memory-bandwidth bound in 3xTF32 (profiles/r01o_gemm_sweep.log: 4096^2 0.52 ms with 256-wide
// tiles against 0.62 ms with 128-wide ones).  Narrow or short outputs take the shape that wastes the least of the
// MMA: n <= 64 -> 128 x 64 single-CTA tiles, m <= 128 -> single-CTA tiles, n <= 128 -> 256 x 128 pairs.  Small tile
// counts are no reason to narrow the tile any more: split-K units fill the SMs instead (plan_units).
// JZ_GEMM_CG=1|2 and JZ_GEMM_TN=64|128|256 force the choice.
static void pick_tile(size_t m, size_t n, int& cg, int& tn) {
    static const int f_cg = env_int("JZ_GEMM_CG"), f_tn = env_int("JZ_GEMM_TN");
    if (n <= 64) { cg = 1; tn = 64; }
    else if (m <= 128) { cg = 1; tn = n <= 128 ? 128 : 256; }
    else if (n <= 128) { cg = 2; tn = 128; }
    else { cg = 2; tn = 256; }
    if (f_cg == 1 || f_cg == 2) cg = f_cg;
    if (f_tn == 64 || f_tn == 128 || f_tn == 256) tn = f_tn;
    if (cg == 2 && tn == 64) tn = 128;
}

// k-blocks (of 32) chained inside TMEM before RN promotion: 3xTF32 keeps the chain short for
// fp32-grade accuracy; TF32 mode is input-rounding dominated so long chains are harmless.
static int chunk_kb(bool split) {
    static const int forced = env_int("JZ_GEMM_CHUNK_KB");
    if (forced > 0) return forced;
    return split ? 4 : 32;
}

// Units of the launch: whole waves of tiles run whole; the tiles of the partial last wave (all tiles when there is
// less than one wave) are split along k so that the tail occupies every SM (pair).  A split must keep at least
// MIN_KB k-blocks so the fix-up (one partial tile out and back through L2) stays a fraction of its mainloop.
// JZ_GEMM_SPLITK=0 disables, =N forces N splits of every tile of the tail.
static void plan_units(GemmArgs& a, int cg, bool split3x, unsigned batch) {
    static const char* env = std::getenv("JZ_GEMM_SPLITK");
    static const int forced = env && *env ? std::atoi(env) : -1;
    const int num_kb = int(ceil_div(a.k, size_t(BK)));
    const unsigned tiles = a.tiles_m * a.tiles_n;
    a.full_tiles = tiles;
    a.splits = 1;
    a.kb_per_split = num_kb;
    if (batch > 1 || forced == 0 || forced == 1) return;
    const unsigned slots = unsigned(ctx().sm_count) / unsigned(cg);
    const unsigned rem = tiles % slots;
    if (rem == 0 || rem >= kTicketSlots / 2) return;
    constexpr int MIN_KB = 8;
    int S = forced > 1 ? forced : int(slots / rem);
    if (S > MAX_SPLITS) S = MAX_SPLITS;
    if (S > num_kb / MIN_KB) S = num_kb / MIN_KB;
    if (S < 2) return;
    int kbs = (num_kb + S - 1) / S;
    const int kbc = a.kb_per_chunk;
    if (kbs > kbc) kbs = ((kbs + kbc - 1) / kbc) * kbc;   // whole TMEM chunks per split
    S = (num_kb + kbs - 1) / kbs;
    if (S < 2) return;
    // the fix-up (partial tile out through L2, ticket, slice back in) costs about as much as this many k-blocks of
    // mainloop (a k-block is ~0.8 us in 3xTF32, ~0.27 us in TF32): split only when the tail gets shorter by more
    const int fixup_kb = split3x ? 8 : 24;
    if (forced <= 1 && num_kb - kbs <= fixup_kb) return;
    a.full_tiles = tiles - rem;
    a.splits = S;
    a.kb_per_split = kbs;
}

#ifdef JZ_GEMM_PROFILE
long long* prof_buf() {
    static long long* p = [] { void* q = nullptr; cudaMalloc(&q, 32 * sizeof(long long)); cudaMemset(q, 0, 32 * sizeof(long long)); return static_cast<long long*>(q); }();
    return p;
}
#endif

static int gemm_tc(int ta, int tb, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda,
                   const float* B, size_t ldb, float beta, float* C, size_t ldc, bool split3x, const ChainParams& chain,
                   float* const* peers, int n_peers, float* mc, unsigned batch, size_t strideA, size_t strideB,
                   size_t strideC, cudaStream_t s) {
    Operand a, b;
    int rc = prepare_operand(a, A, lda, /*kmajor_src=*/ta != 0, m, k, batch, strideA, s);
    if (rc == JZ_OK) rc = prepare_operand(b, B, ldb, /*kmajor_src=*/tb == 0, n, k, batch, strideB, s);
    void* ws = nullptr;
    if (rc == JZ_OK) {
        GemmArgs args;
        args.m = m; args.n = n; args.k = k;
        args.alpha = alpha; args.beta = beta;
        args.C = C; args.ldc = ldc; args.strideC = strideC;
        args.kb_per_chunk = chunk_kb(split3x);
        args.n_peers = n_peers;
        for (int q = 0; q < JZ_MAX_PEERS; q++) args.peers[q] = q < n_peers ? peers[q] : nullptr;
        args.mc = mc;
        args.chain = chain;
        args.ws = nullptr;
        args.tickets = nullptr;
        args.walk_units = 0;
        args.cluster_split = 0;
#ifdef JZ_GEMM_PROFILE
        args.prof = prof_buf();
#endif
        int cg, tn;
        pick_tile(m, n, cg, tn);
        // 3xTF32 on single-CTA tiles (n <= 64 or m <= 128): A's hi / lo parts go to tensor memory instead of a second
        // shared-memory copy -- these shapes are bound by shared-memory bandwidth and by the bytes in flight per SM, not
        // by the tensor pipe.  JZ_GEMM_TS=0 keeps the shared-memory form, =2 forces 128 x 128 TMEM-A tiles everywhere.
        static const int f_ts = [] {
            const char* e = std::getenv("JZ_GEMM_TS");
            return e && *e ? std::atoi(e) : -1;
        }();
        bool ts = split3x && f_ts != 0 && cg == 1;
        if (split3x && f_ts == 2) { ts = true; cg = 1; tn = (n <= 64 || env_int("JZ_GEMM_TN") == 64) ? 64 : 128; }
        if (ts && tn == 256) tn = 128;
        args.tiles_m = unsigned(ceil_div(m, size_t(cg * TILE_M)));
        args.tiles_n = unsigned(ceil_div(n, size_t(tn)));
        plan_units(args, cg, split3x, batch);
        if (split3x && f_ts < 0 && !ts && cg == 2 && tn == 256 && batch == 1 &&
            size_t(args.tiles_m) * args.tiles_n <= size_t(ctx().sm_count) / 4) {
            // A 3xTF32 product of fewer 256 x 256 tiles than half the CTA pairs: it either splits every tile along k (and
            // pays ~18 kclk for the partials' trip through L2) or leaves SMs idle.  128 x 128 TMEM-A tiles on single CTAs
            // cover the chip without (or with fewer) splits; they cost more per k-block and per output element, so they win
            // only while k is short.  Cycle estimates fitted to profiles/r02f_gemm_tile_plans.log (1024^3 22.3 -> 20.7 us,
            // 1280^3 36.1 -> 31.5, 1536^3 40.3 -> 36.2; 1024 x 784 x 8192 stays on pairs: 73 against 99 us).
            GemmArgs alt = args;
            alt.tiles_m = unsigned(ceil_div(m, size_t(TILE_M)));
            alt.tiles_n = unsigned(ceil_div(n, size_t(128)));
            plan_units(alt, 1, split3x, batch);
            const double units_alt = double(alt.full_tiles) + double(alt.tiles_m * alt.tiles_n - alt.full_tiles) * alt.splits;
            const double waves_alt = std::ceil(units_alt / double(ctx().sm_count));
            // (k-splits that can meet through distributed shared memory -- one cluster per tile, all resident: at most 74 / 33 / 15
            // clusters of 2 / 4 / 8 full-SM CTAs on 148 SMs, checked again by launch_tc -- cost about half of the workspace
            // form's fix-up: 1024^3 on TMEM-A tiles 20.9 -> 17.6 us, 1280^3 on pairs 35.8 -> 31.5, profiles/r02h_cluster_split.log)
            auto cluster_ok = [](const GemmArgs& g, int cg_) {
                const int cs = cg_ * g.splits;
                const unsigned t = g.tiles_m * g.tiles_n;
                return g.full_tiles == 0 && (g.splits == 2 || g.splits == 4) && (cs == 2 ? t <= 74 : cs == 4 ? t <= 33 : cs == 8 && t <= 15);
            };
            const double t_pair = double(args.kb_per_split) * (args.kb_per_split >= 32 ? 1650.0 : 2100.0) +
                                  (args.splits > 1 ? (cluster_ok(args, 2) ? 16000.0 : 18000.0) : 6000.0);
            const double t_alt = waves_alt * double(alt.kb_per_split) * 1350.0 + (alt.splits > 1 ? (cluster_ok(alt, 1) ? 4000.0 : 9000.0) : 3000.0);
            if (t_alt < t_pair) {
                args = alt;
                ts = true; cg = 1; tn = 128;
            }
        }
        if (args.splits == 1 && cg == 2 && tn == 256 && batch == 1 &&
            size_t(args.tiles_m) * args.tiles_n * 2 <= size_t(ctx().sm_count) / 2) {
            // a small grid that is not worth splitting along k: narrow tiles at least double the number of busy SM pairs
            tn = 128;
            args.tiles_n = unsigned(ceil_div(n, size_t(tn)));
            plan_units(args, cg, split3x, batch);
        }
        const unsigned split_tiles = args.tiles_m * args.tiles_n - args.full_tiles;
        // Every tile split 2 or 4 ways (a product of fewer tiles than half the SMs): the units of a tile form a thread-block
        // cluster and exchange their partial tiles through distributed shared memory instead of workspace + tickets
        // (launch_tc falls back to the workspace form when the clusters cannot all be resident).  JZ_GEMM_CLUSTER_SPLIT=0 disables.
        static const int f_cs = [] {
            const char* e = std::getenv("JZ_GEMM_CLUSTER_SPLIT");
            return e && *e ? std::atoi(e) : -1;
        }();
        args.cluster_split = (f_cs != 0 && split_tiles && args.full_tiles == 0 && (args.splits == 2 || args.splits == 4) &&
                              tn / args.splits >= 32 && cg * args.splits <= 8 && batch == 1) ? 1 : 0;
        if (split_tiles) {
            // arrival / departure counters: [0, tiles) and [kTicketSlots/2, ...) of the per-stream ticket block
            // (self-resetting, so no memset sits between two launches and programmatic dependent launch still applies)
            args.tickets = tickets_for(s);
            if (!args.tickets) rc = fail(JZ_ERR_CUDA, "gemm: could not allocate the ticket counters");
            if (rc == JZ_OK) rc = ws_alloc(&ws, size_t(split_tiles) * size_t(args.splits) * size_t(cg * TILE_M * tn) * sizeof(float), s);
            if (rc == JZ_OK) args.ws = static_cast<float*>(ws);
        }
        // single-pass TF32 with more units than SM pairs: the persistent kernel (epilogue of unit i under the mainloop of
        // unit i + 1).  JZ_GEMM_PERSIST=0 disables it, =1 forces it whenever the shape allows.
        static const int f_persist = [] {
            const char* e = std::getenv("JZ_GEMM_PERSIST");
            return e && *e ? std::atoi(e) : -1;
        }();
        const unsigned n_units = args.full_tiles + split_tiles * unsigned(args.splits);
        const bool persist_ok = !split3x && cg == 2 && tn == 256 && batch == 1;
        // ... while the operands of a wave of tiles fit in L2 (up to 8192^2 x 2): beyond that the tiles of a wave must read
        // their shared A / B panels at the same moment, and persistent pairs drift apart over a long launch (measured at
        // 16384^3: 602 -> 545 TFLOP/s, A.T*B 552 -> 432; at 4096^3 689 -> 729, profiles/r02d_gemm_tf32_persistent_ab.log)
        const bool fits_l2_wave = (double(m) + double(n)) * double(k) * 4.0 <= 540e6;
        const bool persist = persist_ok && (f_persist == 1 || (f_persist != 0 && fits_l2_wave && n_units > unsigned(ctx().sm_count) / 2));
        // A strided batch with more (member, tile) units than SMs (attention: hundreds to thousands of members of a few
        // k-blocks each): one CTA group per SM (pair) walks the units, so barrier / tensor-memory set-up is paid once per
        // SM and the producer loads the next unit's operands while the current one is stored.  No program in the epilogue
        // then (its scratch would overlay live operand stages).  JZ_GEMM_WALK=0 disables.
        static const bool no_walk = [] { const char* e = std::getenv("JZ_GEMM_WALK"); return e && e[0] == '0'; }();
        const unsigned long long units_all = (unsigned long long)(args.tiles_m) * args.tiles_n * batch;
        const bool walkable = ts || !split3x || cg == 2;   // the variants instantiated with the walk (not the shared-memory 3xTF32 form on single CTAs)
        const bool walk = !no_walk && walkable && batch > 1 && !persist && chain.n == 0 && !chain.bias && n_peers == 0 && !mc &&
                          units_all > (unsigned long long)(ctx().sm_count / cg) && units_all < (1ull << 31);
        if (walk) args.walk_units = unsigned(units_all);
        ctx().gemm_last_walk = walk ? 1 : 0;
        if (rc == JZ_OK) {
            const unsigned zcap = walk ? 0xFFFFFFFFu : 65535u;   // grid.z limit (a walk has no grid.z)
            for (unsigned long long b0l = 0; b0l < batch && rc == JZ_OK; b0l += zcap) {
                const unsigned b0 = unsigned(b0l);
                const unsigned nb = batch - b0 < zcap ? batch - b0 : zcap;
                Operand ab = a, bb = b;
                ab.ptr += size_t(b0) * a.batch_stride;
                bb.ptr += size_t(b0) * b.batch_stride;
                GemmArgs ar = args;
                ar.C += size_t(b0) * strideC;
                ctx().gemm_last_cluster_split = 0;
                if (persist) { ar.cluster_split = 0; rc = launch_tc_tf32_persistent(ab, bb, ar, s); continue; }
                if (ts) { rc = launch_tc_ts(tn, ab, bb, ar, nb, s); continue; }
                if (split3x) rc = cg == 2 ? launch_tc_cg<MODE_XFORM, 2>(tn, ab, bb, ar, nb, s) : launch_tc_cg<MODE_XFORM, 1>(tn, ab, bb, ar, nb, s);
                else rc = cg == 2 ? launch_tc_cg<MODE_TF32, 2>(tn, ab, bb, ar, nb, s) : launch_tc_cg<MODE_TF32, 1>(tn, ab, bb, ar, nb, s);
            }
        }
        ctx().gemm_last_splits = args.splits;
    }
    if (ws) ws_free(ws, s);
    if (a.owned) ws_free(a.owned, s);
    if (b.owned) ws_free(b.owned, s);
    if (rc == JZ_OK) ctx().gemm_last_path = 1;
    return rc;
}

}  // namespace tc

// replicate a finished m x n block (ldc) into every GPU's image through the multicast address (paths without a fused
// multicast epilogue: SIMT / rank-1 / small products)
__global__ void __launch_bounds__(256) mc_copy_kernel(float* mc, const float* src, size_t ldc, size_t m, size_t n) {
    pdl_enter();
    for (size_t j = blockIdx.y; j < n; j += gridDim.y)
        for (size_t i = size_t(blockIdx.x) * 256 + threadIdx.x; i < m; i += size_t(gridDim.x) * 256)
            asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc + j * ldc + i), "f"(src[j * ldc + i]) : "memory");
}

struct GemmExtra {   // what the plain single product does not need
    float* const* peers = nullptr;
    int n_peers = 0;
    float* mc = nullptr;          // multicast image of C
    size_t batch = 1, strideA = 0, strideB = 0, strideC = 0;
};

static int gemm_local(int ta, int tb, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda,
                      const float* B, size_t ldb, float beta, float* C, size_t ldc, const ChainParams& chain, int mode,
                      const GemmExtra& x, bool* fused_done, cudaStream_t s) {
    if (m == 0 || n == 0) return JZ_OK;
    if (!C) return fail(JZ_ERR_ARG, "jz_gemm: null C");
    if (ldc < m) return fail(JZ_ERR_SHAPE, "jz_gemm: ldc < m");
    if (mode < 0) mode = ctx().gemm_mode;
    if (mode > JZ_GEMM_FP32_SIMT) return fail(JZ_ERR_ARG, "jz_gemm: bad mode %d", mode);
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
    const bool want_tc = (mode == JZ_GEMM_3XTF32 || mode == JZ_GEMM_TF32) && ctx().cc_major == 10;
    // tensor path: a reasonably filled tile grid (m, n >= 64, k >= 32) from 2^22 multiply-adds up, and EVERY product
    // beyond the small-product kernel's range (2^26): skinny ones too -- a 4096 x 4096 x 48 product wastes most of its
    // tiles and is still an order of magnitude faster there than on the fp32 SIMT kernel
    const double macs = double(m) * double(n) * double(k);
    const bool big_enough = (m >= 64 && n >= 64 && k >= 32 && macs >= double(1 << 22)) || macs > double(1 << 26);
    const bool fits_i32 = m < (size_t(1) << 31) && n < (size_t(1) << 31) && k < (size_t(1) << 31);
    static const bool force_simt = std::getenv("JZ_GEMM_FORCE_SIMT") != nullptr;
    if (want_tc && big_enough && fits_i32 && !force_simt) {
        const bool split3x = mode == JZ_GEMM_3XTF32;
        *fused_done = true;
        // A program with transcendental steps costs ~45 instructions per element.  In the one-shot kernel the epilogue of
        // a tile is not hidden under another tile's mainloop, so on a large output that work sits in the tail of every
        // wave with only the epilogue warps busy (4096^2, log(exp(x/n)+1)/5: +60 us fused against +35 us for a separate
        // streaming pass at full HBM rate, profiles/r02e_bench.json config1).  Such programs are therefore run as a second,
        // streaming pass over C -- same steps, same roundings, bit-identical (fused_equals_separate_bitwise) -- unless the
        // epilogue also carries the gather to other GPUs, or the persistent TF32 kernel (overlapped epilogue) takes the product.
        static const bool keep_fused = std::getenv("JZ_GEMM_KEEP_FUSED") != nullptr;
        bool heavy = false;
        for (int i = 0; i < chain.n; i++) heavy |= chain.kind[i] == JZ_EXP || chain.kind[i] == JZ_LOG || chain.kind[i] == JZ_TANH || chain.kind[i] == JZ_DTANH;
        const bool persistent_tf32 = !split3x && m > 128 && n > 128 && (double(m) + double(n)) * double(k) * 4.0 <= 540e6 &&
                                     ceil_div(m, size_t(256)) * ceil_div(n, size_t(256)) > size_t(ctx().sm_count) / 2;
        if (heavy && !keep_fused && !persistent_tf32 && ldc == m && double(m) * double(n) >= double(1 << 22) && !x.n_peers && !x.mc) {
            ChainParams head = chain;
            head.n = 0;   // alpha / beta / bias stay in the product
            int rc = tc::gemm_tc(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, split3x, head, nullptr, 0, nullptr, 1, 0, 0, 0, s);
            if (rc != JZ_OK) return rc;
            jz_step steps[JZ_MAX_CHAIN];
            for (int i = 0; i < chain.n; i++) steps[i] = jz_step{chain.kind[i], chain.s1[i], chain.a[i]};
            return jz_chain(C, C, m * n, steps, chain.n, reinterpret_cast<jz_stream_t>(s));
        }
        return tc::gemm_tc(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, split3x, chain, x.peers, x.n_peers, x.mc, 1, 0, 0, 0, s);
    }
    if (gemm_small_wants(m, n, k))   // latency-bound shapes: the warp-per-tile fp32 kernel (jz_gemm_small.cu)
        return launch_gemm_small(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain, s);
    return launch_simt(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain, s);
}

static int gemm_entry(int ta, int tb, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda,
                      const float* B, size_t ldb, float beta, float* C, size_t ldc, const ChainParams& chain, int mode,
                      cudaStream_t s, const GemmExtra& x = GemmExtra()) {
    if (x.n_peers < 0 || x.n_peers > JZ_MAX_PEERS || (x.n_peers && !x.peers)) return fail(JZ_ERR_ARG, "jz_gemm: bad peer list");
    bool fused_done = false;
    int rc = gemm_local(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain, mode, x, &fused_done, s);
    if (rc != JZ_OK || fused_done || m == 0 || n == 0) return rc;
    // paths without a fused store (SIMT / rank-1 / small): replicate the finished block afterwards
    if (x.mc) {
        const size_t cap = size_t(ctx().sm_count) * 8;
        size_t gx = ceil_div(m, size_t(256));
        if (gx > cap) gx = cap;
        size_t gy = ceil_div(cap, gx);
        if (gy > n) gy = n;
        if (gy > 65535) gy = 65535;
        JZ_LAUNCH(mc_copy_kernel, dim3((unsigned)gx, (unsigned)gy, 1), 256, 0, s, x.mc, C, ldc, m, n);
        return JZ_OK;
    }
    for (int q = 0; q < x.n_peers; q++) {
        rc = jz_copy2d(x.peers[q], ldc, C, ldc, m, n, 0, s);
        if (rc != JZ_OK) return rc;
    }
    return JZ_OK;
}

}  // namespace jz

using namespace jz;

extern "C" {

int jz_gemm(int transA, int transB, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda,
            const float* B, size_t ldb, float beta, float* C, size_t ldc, int mode, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    ChainParams chain;
    make_chain(chain, nullptr, 0);
    return gemm_entry(transA, transB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain, mode, as_stream(stream));
}

/* strided batch (cublasSgemmStridedBatched in TransformerLayer, ml/layer.hpp:2896-2926): member i uses
   A + i*strideA, B + i*strideB, C + i*strideC.  Tensor-shaped members (m, n >= 64, k >= 32, 16-byte aligned strides)
   run as ONE launch of the tcgen05 kernel with the member on blockIdx.z (3-D tensor maps); smaller members as one
   launch of the small-product kernel; anything else member by member. */
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
    if (mode < 0) mode = ctx().gemm_mode;
    if (mode > JZ_GEMM_FP32_SIMT) return fail(JZ_ERR_ARG, "jz_gemm_strided_batched: bad mode %d", mode);
    ChainParams chain;
    make_chain(chain, nullptr, 0);
    cudaStream_t s = as_stream(stream);
    const double macs = double(m) * double(n) * double(k);
    static const bool no_btc = std::getenv("JZ_GEMM_NO_BATCHED_TC") != nullptr;
    // members of at least 64 x 64 x 32 with 2^24 multiply-adds over the whole batch.  Small members pay a tile's fixed cost
    // (prologue, TMEM allocation, epilogue) per 0.5 M multiply-adds and still beat the fp32 small-product kernel 5x
    // (seq 64, d_h 128, batch 2048: 0.09 ms against 0.53 ms; cuBLAS fp32 0.06 ms; profiles/r02a_attention.log, r02e_attention.log)
    const bool tc_member = m >= 64 && n >= 64 && k >= 32 && (macs >= double(1 << 21) || macs * double(batch) >= double(1 << 24)) && batch > 1 &&
                           batch < (size_t(1) << 31) && (mode == JZ_GEMM_3XTF32 || mode == JZ_GEMM_TF32) && ctx().cc_major == 10 &&
                           m < (size_t(1) << 31) && n < (size_t(1) << 31) && k < (size_t(1) << 31) && !no_btc;
    const bool tma_ok = lda % 4 == 0 && ldb % 4 == 0 && strideA % 4 == 0 && strideB % 4 == 0 && aligned16(A) && aligned16(B);
    if (tc_member && tma_ok)
        return tc::gemm_tc(transA, transB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mode == JZ_GEMM_3XTF32, chain, nullptr, 0,
                           nullptr, unsigned(batch), strideA, strideB, strideC, s);
    if (k >= 2 && macs < double(1 << 24) && gemm_small_wants(m, n, k)) {
        for (size_t b0 = 0; b0 < batch; b0 += 65535) {   // grid.z limit
            const size_t nb = batch - b0 < 65535 ? batch - b0 : 65535;
            int rc = launch_gemm_small(transA, transB, m, n, k, alpha, A + b0 * strideA, lda, B + b0 * strideB, ldb, beta,
                                       C + b0 * strideC, ldc, chain, s, nb, strideA, strideB, strideC);
            if (rc != JZ_OK) return rc;
        }
        return JZ_OK;
    }
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

int jz_gemm_bias_chain(int transA, int transB, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda,
                       const float* B, size_t ldb, float* C, size_t ldc, const float* bias, int bias_dim, float s1, float s2,
                       const jz_step* steps, int nsteps, int mode, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    ChainParams chain;
    if (make_chain(chain, steps, nsteps) != JZ_OK) return fail(JZ_ERR_ARG, "jz_gemm_bias_chain: bad step list");
    if (bias) {
        if (bias_dim != 0 && bias_dim != 1) return fail(JZ_ERR_ARG, "jz_gemm_bias_chain: bias_dim must be 0 or 1");
        chain.bias = bias;
        chain.bias_dim = bias_dim;
        chain.bias_s1 = s1;
        chain.bias_s2 = s2;
    }
    return gemm_entry(transA, transB, m, n, k, alpha, A, lda, B, ldb, 0.0f, C, ldc, chain, mode, as_stream(stream));
}

int jz_gemm_chain_bcast(int transA, int transB, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda,
                        const float* B, size_t ldb, float* C, size_t ldc, float* const* peer_C, int n_peers,
                        const jz_step* steps, int nsteps, int mode, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    ChainParams chain;
    if (make_chain(chain, steps, nsteps) != JZ_OK) return fail(JZ_ERR_ARG, "jz_gemm_chain_bcast: bad step list");
    GemmExtra x;
    x.peers = peer_C;
    x.n_peers = n_peers;
    return gemm_entry(transA, transB, m, n, k, alpha, A, lda, B, ldb, 0.0f, C, ldc, chain, mode, as_stream(stream), x);
}

int jz_gemm_chain_mcast(int transA, int transB, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda,
                        const float* B, size_t ldb, float* C, size_t ldc, float* mc_C, const jz_step* steps, int nsteps,
                        int mode, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    if (!mc_C) return fail(JZ_ERR_ARG, "jz_gemm_chain_mcast: null multicast address");
    ChainParams chain;
    if (make_chain(chain, steps, nsteps) != JZ_OK) return fail(JZ_ERR_ARG, "jz_gemm_chain_mcast: bad step list");
    GemmExtra x;
    x.mc = mc_C;
    return gemm_entry(transA, transB, m, n, k, alpha, A, lda, B, ldb, 0.0f, C, ldc, chain, mode, as_stream(stream), x);
}

int jz_gemm_last_splits(void) { return ctx().gemm_last_splits; }
int jz_gemm_last_cluster_split(void) { return ctx().gemm_last_cluster_split; }
int jz_gemm_last_walk(void) { return ctx().gemm_last_walk; }

}  // extern "C"

#ifdef JZ_GEMM_PROFILE
// instrumented builds only (not declared in include/jz_b200.h): the cycle counters of the last tensor-core GEMM launch
extern "C" __attribute__((visibility("default"))) int jz_debug_gemm_prof(long long* out32) {
    cudaDeviceSynchronize();
    const bool ok = cudaMemcpy(out32, jz::tc::prof_buf(), sizeof(long long) * 32, cudaMemcpyDeviceToHost) == cudaSuccess;
    cudaMemset(jz::tc::prof_buf(), 0, sizeof(long long) * 32);   // roles that do not exist in the next launch's CTA 5 read as 0
    return ok ? 0 : 1;
}
#endif
