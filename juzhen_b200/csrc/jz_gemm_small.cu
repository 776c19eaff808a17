// jz_gemm_small.cu -- fp32 FMA GEMM for SMALL products (at most a few 10^7 multiply-adds): the shapes of a
// training step at batch 32 ((1024 x 784)(784 x 32), (128 x 1024)(1024 x 32), (1024 x 32)(32 x 784),
// (10 x 128)(128 x 32) ...), replacing cublasSgemm (cpp/cumatrix.cu:177-197) where a tensor-core tile would be
// mostly padding and its pipeline prologue longer than the product.
//
// These products are latency problems, not throughput problems: the whole of A fits in L2 and the math is a few
// microseconds of one SM row.  A classic shared-memory-tiled kernel serialises global-load latency once per k-tile
// (measured on B200: 70-90 us for the two forward products above, 57 % of the demo_mnist step).  Here instead:
//   * a WARP owns 32 rows x 8 columns of C for a slice of k, its lanes own the rows: A is read with one
//     coalesced 128-byte load per k (plain A) or one 128-bit load per lane per 4 k (flagged-transpose A); the
//     4 k x 8 columns of B an iteration needs are one element per lane, loaded once and shared by shuffle;
//     no shared memory and no barrier inside the k loop, several iterations' loads in flight;
//   * the 8 / 16 / 32 warps of a CTA split k (every iteration of a warp is one memory round trip, so deep
//     products want many short slices); the partial tiles meet in shared memory once, at the end, and are
//     summed in a fixed order -> deterministic results;
//   * narrow column blocks keep the grid at >= one CTA per SM for the shapes that matter; A is re-read from L2
//     by the CTAs that share a row block.
// Measured on B200 inside a launch loop (scripts/small_gemm_bench.py, warm L2): (1024 x 784)(784 x 32) 51 -> 9 us,
// (128 x 1024)(1024 x 32) 71 -> 10 us, (1024 x 784)^T(1024 x 32) 72 -> 14 us against the shared-memory-tiled kernel.
// ncu with a warm cache: 10 us, 1.96 instructions per cycle per SM, a third of the stalls on the shuffle (MIO) queue --
// what is left is instruction issue (32 shuffles per 32 FMAs), not memory latency.
// Products with m, n >= 64 and k >= 32 stay on the tensor-core path (8 us at (1024 x 32)(32 x 784)).
// alpha/beta and the fused elementwise chain are applied exactly as in the other GEMM kernels.
#include "jz_common.cuh"
#include "jz_math.cuh"

namespace jz {


// op(A)(i, kk .. kk+3) for this lane's row
template <bool TA, bool VEC>
__device__ __forceinline__ void sk_load_a(float (&a)[4], const float* __restrict__ A, size_t lda, size_t i, size_t kk) {
    if constexpr (!TA) {
#pragma unroll
        for (int q = 0; q < 4; q++) a[q] = A[(kk + q) * lda + i];
    } else if constexpr (VEC) {
        const float4 v = *reinterpret_cast<const float4*>(A + i * lda + kk);
        a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
    } else {
#pragma unroll
        for (int q = 0; q < 4; q++) a[q] = A[i * lda + kk + q];
    }
}

// The 4 k x 8 columns of op(B) one iteration needs are exactly one float per lane: each lane loads ONE element
// (plain B: lane = 4*j + q reads B[(j0+j)*ldb + kk + q], eight 16-byte runs; flagged B: lane = 8*q + j reads
// B[(kk+q)*ldb + j0 + j], four 32-byte runs), parks it in a 128-byte warp-private shared-memory line, and every lane
// reads the line back as eight broadcast 128-bit words {B(kk..kk+3, j)}.  One global load instruction per iteration
// instead of eight warp-uniform ones, one live register instead of 32 (several iterations' loads fit in flight under
// the 64-register budget of a 1024-thread CTA), and 9 shared-memory instructions instead of the 32 shuffles of the
// register-only form, which left the kernel issue-bound on the shuffle queue (ncu: a third of all stalls).
template <bool TB, int NT>
__device__ __forceinline__ float sk_load_b(const float* __restrict__ B, size_t ldb, size_t kk, size_t j0, size_t n, int lane) {
    static_assert(NT == 8, "one B element per lane: 4 k x 8 columns");
    const int q = TB ? lane >> 3 : lane & 3, j = TB ? lane & 7 : lane >> 2;
    const size_t jj = j0 + j < n ? j0 + j : n - 1;   // clamped: surplus columns are computed, never stored
    return TB ? B[(kk + q) * ldb + jj] : B[jj * ldb + kk + q];
}
// line: this warp's 32-float buffer for this iteration (double-buffered by the caller); slot j*4 + q holds B(kk+q, j)
template <bool TB, int NT>
__device__ __forceinline__ void sk_fma_b(float (&acc)[NT], const float (&a)[4], float bval, float* line, int lane) {
    line[TB ? ((lane & 7) << 2) + (lane >> 3) : lane] = bval;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NT; j++) {
        const float4 b = reinterpret_cast<const float4*>(line)[j];
        acc[j] = fmaf(a[0], b.x, acc[j]);
        acc[j] = fmaf(a[1], b.y, acc[j]);
        acc[j] = fmaf(a[2], b.z, acc[j]);
        acc[j] = fmaf(a[3], b.w, acc[j]);
    }
}

// KSPLIT: the CTA's warps split k (true) or take different column blocks (false).
template <bool TA, bool TB, int NT, bool KSPLIT, bool VEC, int SK_WARPS>
__global__ void __launch_bounds__(32 * SK_WARPS) gemm_small_kernel(size_t m, size_t n, size_t k, float alpha,
                                                                   const float* __restrict__ A, size_t lda,
                                                                   const float* __restrict__ B, size_t ldb, float beta,
                                                                   float* C, size_t ldc, ChainParams chain_p,
                                                                   size_t strideA, size_t strideB, size_t strideC, unsigned gx) {
    pdl_enter();
    __shared__ ChainParams chain;
    __shared__ float red[KSPLIT ? SK_WARPS : 1][KSPLIT ? NT : 1][32];
    __shared__ __align__(16) float bline[SK_WARPS][2][32];   // per warp, two iterations deep
    stage_chain(&chain, chain_p, threadIdx.x);
    A += size_t(blockIdx.z) * strideA;   // strided batch (blockIdx.z = 0 for a single product)
    B += size_t(blockIdx.z) * strideB;
    C += size_t(blockIdx.z) * strideC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // linear block index = row block + gx * column block (grid.y would cap n at 65535 column blocks)
    const unsigned bx = blockIdx.x % gx, by = blockIdx.x / gx;
    const size_t i = size_t(bx) * 32 + lane;
    const size_t il = i < m ? i : m - 1;   // clamped row for loads
    const size_t j0 = (size_t(by) * (KSPLIT ? 1 : SK_WARPS) + (KSPLIT ? 0 : warp)) * NT;
    size_t kbeg = 0, kend = k;
    if (KSPLIT) {
        const size_t per = ((k + 4 * SK_WARPS - 1) / (4 * SK_WARPS)) * 4;   // multiple of 4: slices keep 16-byte phase
        kbeg = size_t(warp) * per < k ? size_t(warp) * per : k;
        kend = kbeg + per < k ? kbeg + per : k;
    }
    float acc[NT];
#pragma unroll
    for (int j = 0; j < NT; j++) acc[j] = 0.0f;
    if (KSPLIT || j0 < n) {
        size_t kk = kbeg;
        unsigned par = 0;
#pragma unroll 4
        for (; kk + 4 <= kend; kk += 4) {
            float a[4];
            sk_load_a<TA, VEC>(a, A, lda, il, kk);
            const float bval = sk_load_b<TB, NT>(B, ldb, kk, j0, n, lane);
            sk_fma_b<TB, NT>(acc, a, bval, bline[warp][par], lane);
            par ^= 1u;
        }
        for (; kk < kend; kk++) {   // k tail, one at a time
            const float a = TA ? A[il * lda + kk] : A[kk * lda + il];
#pragma unroll
            for (int j = 0; j < NT; j++) {
                const size_t jj = j0 + j < n ? j0 + j : n - 1;
                acc[j] = fmaf(a, TB ? B[kk * ldb + jj] : B[jj * ldb + kk], acc[j]);
            }
        }
    }
    __syncthreads();   // chain staged (and, below, partial tiles complete)
    if constexpr (KSPLIT) {
#pragma unroll
        for (int j = 0; j < NT; j++) red[warp][j][lane] = acc[j];
        __syncthreads();
        // 32 x NT outputs: warp w sums the partials of columns w, w + SK_WARPS, ... (lane = row), fixed order
#pragma unroll
        for (int r = 0; r < (NT + SK_WARPS - 1) / SK_WARPS; r++) {
            const int j = warp + SK_WARPS * r;
            if (j >= NT) break;
            float s = 0.0f;
#pragma unroll
            for (int w = 0; w < SK_WARPS; w++) s += red[w][j][lane];
            if (i < m && j0 + j < n) {
                float x[1] = {alpha * s};
                float* dst = C + (j0 + j) * ldc + i;
                if (beta != 0.0f) x[0] += beta * *dst;
                if (chain.bias) x[0] = apply_bias(x[0], chain, i, j0 + j);
                if (chain.n) apply_chain<1>(x, chain);
                *dst = x[0];
            }
        }
    } else {
        if (i < m) {
#pragma unroll
            for (int j = 0; j < NT; j++) {
                if (j0 + j < n) {
                    float x[1] = {alpha * acc[j]};
                    float* dst = C + (j0 + j) * ldc + i;
                    if (beta != 0.0f) x[0] += beta * *dst;
                    if (chain.bias) x[0] = apply_bias(x[0], chain, i, j0 + j);
                    if (chain.n) apply_chain<1>(x, chain);
                    *dst = x[0];
                }
            }
        }
    }
}

struct SmallBatch { size_t count = 1, sA = 0, sB = 0, sC = 0; };

template <bool TA, bool TB, int NT, bool KSPLIT, int W>
static int launch_small_v(bool vec, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda, const float* B,
                          size_t ldb, float beta, float* C, size_t ldc, const ChainParams& chain, const SmallBatch& bt,
                          cudaStream_t s) {
    const size_t gx = ceil_div(m, size_t(32));
    const size_t gy = ceil_div(n, size_t(NT) * (KSPLIT ? 1 : W));
    if (gx * gy >= (size_t(1) << 31) || bt.count > 65535) return fail(JZ_ERR_UNSUPPORTED, "small gemm: output or batch too large");
    const dim3 grid((unsigned)(gx * gy), 1, (unsigned)bt.count);
    const unsigned gxu = unsigned(gx);
    if (vec) JZ_LAUNCH((gemm_small_kernel<TA, TB, NT, KSPLIT, true, W>), grid, 32 * W, 0, s, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain, bt.sA, bt.sB, bt.sC, gxu);
    else JZ_LAUNCH((gemm_small_kernel<TA, TB, NT, KSPLIT, false, W>), grid, 32 * W, 0, s, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain, bt.sA, bt.sB, bt.sC, gxu);
    return JZ_OK;
}
// variants: k split over 8 / 16 / 32 warps, 8-column blocks
template <bool TA, bool TB>
static int launch_small_cfg(bool ksplit, int warps, int nt, bool vec, size_t m, size_t n, size_t k, float alpha,
                            const float* A, size_t lda, const float* B, size_t ldb, float beta, float* C, size_t ldc,
                            const ChainParams& chain, const SmallBatch& bt, cudaStream_t s) {
#define JZ_SMALL_CASE(KS, W, NT_) \
    if (ksplit == KS && warps == W && nt == NT_) \
        return launch_small_v<TA, TB, NT_, KS, W>(vec, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain, bt, s);
    JZ_SMALL_CASE(true, 8, 8) JZ_SMALL_CASE(true, 16, 8) JZ_SMALL_CASE(true, 32, 8)
#undef JZ_SMALL_CASE
    return fail(JZ_ERR_ARG, "small gemm: no kernel variant for ksplit=%d warps=%d nt=%d", int(ksplit), warps, nt);
}

// is this product one for the small kernel?  (m, n, k >= 1 checked by the caller)
bool gemm_small_wants(size_t m, size_t n, size_t k) {
    static const bool off = [] { const char* e = std::getenv("JZ_GEMM_NO_SMALL"); return e && e[0] && e[0] != '0'; }();
    if (off || k < 2) return false;
    return double(m) * double(n) * double(k) <= double(1 << 26);
}

static int env_int(const char* name) {
    const char* e = std::getenv(name);
    return e && *e ? std::atoi(e) : -1;
}

int launch_gemm_small(int ta, int tb, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda,
                      const float* B, size_t ldb, float beta, float* C, size_t ldc, const ChainParams& chain,
                      cudaStream_t s, size_t batch, size_t strideA, size_t strideB, size_t strideC) {
    SmallBatch bt;
    bt.count = batch; bt.sA = strideA; bt.sB = strideB; bt.sC = strideC;
    // the 128-bit loads of a flagged-transpose A need 16-byte phase on its k runs (every batch member included)
    const bool vec = !ta || (lda % 4 == 0 && aligned16(A) && (batch <= 1 || strideA % 4 == 0));
    static const int f_warps = env_int("JZ_SMALL_WARPS");
    const bool ksplit = true;
    const int nt = 8;
    // more warps = shorter serial k per warp (measured: k = 784: 16.4 / 14.4 / 10.3 us with 8 / 16 / 32 warps)
    int warps = k >= 512 ? 32 : (k >= 128 ? 16 : 8);
    if (f_warps == 8 || f_warps == 16 || f_warps == 32) warps = f_warps;
    int rc;
    if (!ta && !tb) rc = launch_small_cfg<false, false>(ksplit, warps, nt, vec, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain, bt, s);
    else if (ta && !tb) rc = launch_small_cfg<true, false>(ksplit, warps, nt, vec, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain, bt, s);
    else if (!ta && tb) rc = launch_small_cfg<false, true>(ksplit, warps, nt, vec, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain, bt, s);
    else rc = launch_small_cfg<true, true>(ksplit, warps, nt, vec, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, chain, bt, s);
    if (rc == JZ_OK) ctx().gemm_last_path = 4;
    return rc;
}

}  // namespace jz
