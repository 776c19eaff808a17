// jz_layout.cu -- transpose-aware 2-D data movement and binary ops
// (SURVEY 8a rows a6, a10 mixed-flag cases, a14 T() materialisation, a15 slice, a16 stack,
// and the broadcast idioms of 8f-1).
//
// The reference funnels every one of these through cublasSgeam or an int-indexed scalar
// gather (cpp/cumatrix.cu:227-260, cpp/cukernels.cu:72-90,258-323).  Here:
//   * same-orientation 2-D maps: threads run down the contiguous (row) dimension with
//     128-bit accesses when rows/ld/pointers allow, several columns per thread in flight;
//   * anything with a transposed operand: 32x32 shared-memory tiles (+1 padding, no bank
//     conflicts); both the global read and the global write are coalesced 128 B rows;
//   * all index math in size_t (the reference's int math overflows past 2^31 elements).
// Data movement is bit-exact by construction (plain copies, no arithmetic).
#include "jz_common.cuh"
#include "jz_math.cuh"

namespace jz {

constexpr int kTile = 32;
constexpr int kTileRows = 8;  // block = 32 x 8 threads, 4 tile rows per thread
constexpr int kColsPerThread = 4;

// ------------------------------------------------------------------ same-orientation 2-D map
struct Map2dGeom {
    unsigned tx, ty;        // block shape
    dim3 grid;
};

static Map2dGeom geom2d(size_t row_units, size_t cols) {
    unsigned tx = 1;
    while (tx < 256 && tx < row_units) tx <<= 1;
    unsigned ty = 256 / tx;
    Map2dGeom g;
    g.tx = tx;
    g.ty = ty;
    size_t gx = ceil_div(row_units, tx);
    size_t gy = ceil_div(cols, size_t(ty) * kColsPerThread);
    if (gy > 32768) gy = 32768;
    if (gx > 2147483647u) gx = 2147483647u;
    g.grid = dim3(unsigned(gx ? gx : 1), unsigned(gy ? gy : 1));
    return g;
}

// Op::run<VEC>(i, j): processes VEC consecutive rows starting at row i of column j.
template <int VEC, class Op>
__global__ void __launch_bounds__(256) map2d_kernel(Op op, size_t rows, size_t cols) {
    pdl_enter();
    const size_t row_units = rows / VEC;
    for (size_t iu = size_t(blockIdx.x) * blockDim.x + threadIdx.x; iu < row_units;
         iu += size_t(gridDim.x) * blockDim.x) {
        for (size_t j0 = (size_t(blockIdx.y) * blockDim.y + threadIdx.y) * kColsPerThread; j0 < cols;
             j0 += size_t(gridDim.y) * blockDim.y * kColsPerThread) {
#pragma unroll
            for (int c = 0; c < kColsPerThread; c++)
                if (j0 + c < cols) op.template run<VEC>(iu * VEC, j0 + c);
        }
    }
}

struct Copy2dOp {
    float* dst; size_t ldd; const float* src; size_t lds;
    template <int VEC>
    __device__ __forceinline__ void run(size_t i, size_t j) const {
        if constexpr (VEC == 4)
            *reinterpret_cast<float4*>(dst + j * ldd + i) = *reinterpret_cast<const float4*>(src + j * lds + i);
        else
            dst[j * ldd + i] = src[j * lds + i];
    }
};

template <class F>
struct Bin2dOp {
    float* out; size_t ldo; const float* a; size_t lda; const float* b; size_t ldb; F f;
    template <int VEC>
    __device__ __forceinline__ void run(size_t i, size_t j) const {
        if constexpr (VEC == 4) {
            const float4 x = *reinterpret_cast<const float4*>(a + j * lda + i);
            const float4 y = *reinterpret_cast<const float4*>(b + j * ldb + i);
            float4 r;
            r.x = f(x.x, y.x); r.y = f(x.y, y.y); r.z = f(x.z, y.z); r.w = f(x.w, y.w);
            *reinterpret_cast<float4*>(out + j * ldo + i) = r;
        } else {
            out[j * ldo + i] = f(a[j * lda + i], b[j * ldb + i]);
        }
    }
};

struct AxpbyF2 {
    float s1, s2;
    __device__ __forceinline__ float operator()(float x, float y) const {
        return __fadd_rn(__fmul_rn(s1, x), __fmul_rn(s2, y));
    }
};
struct MulF2 {
    __device__ __forceinline__ float operator()(float x, float y) const { return __fmul_rn(x, y); }
};

// out(i,j) = s1*a(i,j) + s2*v[dim==1 ? i : j]   (a, out contiguous rows x cols)
struct BcastOp {
    float* out; const float* a; const float* v; size_t rows; int dim; float s1, s2;
    template <int VEC>
    __device__ __forceinline__ void run(size_t i, size_t j) const {
        if constexpr (VEC == 4) {
            const float4 x = *reinterpret_cast<const float4*>(a + j * rows + i);
            float4 w;
            if (dim == 1) w = *reinterpret_cast<const float4*>(v + i);
            else { const float t = v[j]; w = make_float4(t, t, t, t); }
            float4 r;
            r.x = __fadd_rn(__fmul_rn(s1, x.x), __fmul_rn(s2, w.x));
            r.y = __fadd_rn(__fmul_rn(s1, x.y), __fmul_rn(s2, w.y));
            r.z = __fadd_rn(__fmul_rn(s1, x.z), __fmul_rn(s2, w.z));
            r.w = __fadd_rn(__fmul_rn(s1, x.w), __fmul_rn(s2, w.w));
            *reinterpret_cast<float4*>(out + j * rows + i) = r;
        } else {
            const float w = dim == 1 ? v[i] : v[j];
            out[j * rows + i] = __fadd_rn(__fmul_rn(s1, a[j * rows + i]), __fmul_rn(s2, w));
        }
    }
};

// out(i,j) = u[i] * v[j]
struct OuterOp {
    float* out; size_t ldo; const float* u; const float* v;
    template <int VEC>
    __device__ __forceinline__ void run(size_t i, size_t j) const {
        const float t = v[j];
        if constexpr (VEC == 4) {
            const float4 x = *reinterpret_cast<const float4*>(u + i);
            *reinterpret_cast<float4*>(out + j * ldo + i) =
                make_float4(__fmul_rn(x.x, t), __fmul_rn(x.y, t), __fmul_rn(x.z, t), __fmul_rn(x.w, t));
        } else {
            out[j * ldo + i] = __fmul_rn(u[i], t);
        }
    }
};

template <class Op>
static int launch_map2d(const Op& op, size_t rows, size_t cols, bool vec_ok, cudaStream_t s) {
    if (rows == 0 || cols == 0) return JZ_OK;
    if (vec_ok && (rows % 4 == 0)) {
        Map2dGeom g = geom2d(rows / 4, cols);
        JZ_LAUNCH((map2d_kernel<4, Op>), g.grid, dim3(g.tx, g.ty), 0, s, op, rows, cols);
    } else {
        Map2dGeom g = geom2d(rows, cols);
        JZ_LAUNCH((map2d_kernel<1, Op>), g.grid, dim3(g.tx, g.ty), 0, s, op, rows, cols);
    }
    return JZ_OK;
}

// ------------------------------------------------------------------ tiled transposing kernels
// dst(i,j) = src(j,i); dst rows x cols (ldd), src cols x rows (lds)
__global__ void __launch_bounds__(kTile* kTileRows) transpose_kernel(float* dst, size_t ldd, const float* src,
                                                                     size_t lds, size_t rows, size_t cols,
                                                                     size_t tiles_i, size_t tiles_j) {
    pdl_enter();
    __shared__ float tile[kTile][kTile + 1];
    const size_t ntiles = tiles_i * tiles_j;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        // consecutive CTAs walk down dst rows (i) first: their src reads are adjacent columns
        const size_t ti = t % tiles_i, tj = t / tiles_i;
        const size_t i0 = ti * kTile, j0 = tj * kTile;
        // read: src element (j, i) at i*lds + j ; x runs along j (contiguous)
#pragma unroll
        for (int k = 0; k < kTile; k += kTileRows) {
            const size_t i = i0 + threadIdx.y + k, j = j0 + threadIdx.x;
            if (i < rows && j < cols) tile[threadIdx.y + k][threadIdx.x] = src[i * lds + j];
        }
        __syncthreads();
        // write: dst element (i, j) at j*ldd + i ; x runs along i (contiguous)
#pragma unroll
        for (int k = 0; k < kTile; k += kTileRows) {
            const size_t j = j0 + threadIdx.y + k, i = i0 + threadIdx.x;
            if (i < rows && j < cols) dst[j * ldd + i] = tile[threadIdx.x][threadIdx.y + k];
        }
        __syncthreads();
    }
}

// 64 x 64 tiles, 128-bit global accesses on BOTH sides (requires 16-byte aligned pointers and
// ld % 4 == 0; interior tiles only -- ragged edges take the scalar path inside the same kernel).
// Shared-memory pitch 65 keeps the transposed reads at most 2-way conflicted.
__global__ void __launch_bounds__(256) transpose64_kernel(float* dst, size_t ldd, const float* src, size_t lds,
                                                          size_t rows, size_t cols, size_t tiles_i, size_t tiles_j) {
    pdl_enter();
    __shared__ float tile[64][65];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const size_t ntiles = tiles_i * tiles_j;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const size_t ti = t % tiles_i, tj = t / tiles_i;
        const size_t i0 = ti * 64, j0 = tj * 64;
        const bool full = i0 + 64 <= rows && j0 + 64 <= cols;
        if (full) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int il = ty + 16 * r;  // row of dst == column of src
                const float4 v = *reinterpret_cast<const float4*>(src + (i0 + il) * lds + j0 + 4 * tx);
                tile[il][4 * tx + 0] = v.x; tile[il][4 * tx + 1] = v.y;
                tile[il][4 * tx + 2] = v.z; tile[il][4 * tx + 3] = v.w;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int jl = ty + 16 * r;
                float4 v;
                v.x = tile[4 * tx + 0][jl]; v.y = tile[4 * tx + 1][jl];
                v.z = tile[4 * tx + 2][jl]; v.w = tile[4 * tx + 3][jl];
                *reinterpret_cast<float4*>(dst + (j0 + jl) * ldd + i0 + 4 * tx) = v;
            }
        } else {
            for (int e = threadIdx.x; e < 64 * 64; e += 256) {
                const int il = e >> 6, jl = e & 63;
                if (i0 + il < rows && j0 + jl < cols) tile[il][jl] = src[(i0 + il) * lds + j0 + jl];
            }
            __syncthreads();
            for (int e = threadIdx.x; e < 64 * 64; e += 256) {
                const int jl = e >> 6, il = e & 63;
                if (i0 + il < rows && j0 + jl < cols) dst[(j0 + jl) * ldd + i0 + il] = tile[il][jl];
            }
        }
        __syncthreads();
    }
}

// out(i,j) = f(opA(i,j), opB(i,j)) with at least one transposed operand
template <bool TA, bool TB, class F>
__global__ void __launch_bounds__(kTile* kTileRows) bin2d_t_kernel(float* out, size_t ldo, size_t rows, size_t cols,
                                                                   const float* a, size_t lda, const float* b,
                                                                   size_t ldb, F f, size_t tiles_i, size_t tiles_j) {
    pdl_enter();
    __shared__ float ta[TA ? kTile : 1][kTile + 1];
    __shared__ float tb[TB ? kTile : 1][kTile + 1];
    const size_t ntiles = tiles_i * tiles_j;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const size_t ti = t % tiles_i, tj = t / tiles_i;
        const size_t i0 = ti * kTile, j0 = tj * kTile;
        if (TA || TB) {
#pragma unroll
            for (int k = 0; k < kTile; k += kTileRows) {
                const size_t i = i0 + threadIdx.y + k, j = j0 + threadIdx.x;
                if (i < rows && j < cols) {
                    if (TA) ta[threadIdx.y + k][threadIdx.x] = a[i * lda + j];
                    if (TB) tb[threadIdx.y + k][threadIdx.x] = b[i * ldb + j];
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < kTile; k += kTileRows) {
            const size_t j = j0 + threadIdx.y + k, i = i0 + threadIdx.x;
            if (i < rows && j < cols) {
                const float x = TA ? ta[threadIdx.x][threadIdx.y + k] : a[j * lda + i];
                const float y = TB ? tb[threadIdx.x][threadIdx.y + k] : b[j * ldb + i];
                out[j * ldo + i] = f(x, y);
            }
        }
        if (TA || TB) __syncthreads();
    }
}

static unsigned tile_grid(size_t ntiles) {
    // 4 resident CTAs per SM measured best for the 64 x 64 transposing tiles (6.05 vs 5.40 TB/s at 8)
    const size_t cap = size_t(ctx().sm_count) * 4;
    return unsigned(ntiles < cap ? (ntiles ? ntiles : 1) : cap);
}

template <class F>
static int launch_bin2d(float* out, size_t ldo, size_t rows, size_t cols, const float* a, size_t lda, int ta,
                        const float* b, size_t ldb, int tb, F f, cudaStream_t s) {
    if (rows == 0 || cols == 0) return JZ_OK;
    if (!out || !a || !b) return fail(JZ_ERR_ARG, "null pointer");
    if (ldo < rows || lda < (ta ? cols : rows) || ldb < (tb ? cols : rows))
        return fail(JZ_ERR_SHAPE, "leading dimension smaller than the matrix");
    if (!ta && !tb) {
        const bool vec = aligned16(out) && aligned16(a) && aligned16(b) && ldo % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0;
        return launch_map2d(Bin2dOp<F>{out, ldo, a, lda, b, ldb, f}, rows, cols, vec, s);
    }
    const size_t tiles_i = ceil_div(rows, kTile), tiles_j = ceil_div(cols, kTile);
    const dim3 block(kTile, kTileRows);
    const unsigned grid = tile_grid(tiles_i * tiles_j);
    if (ta && tb) JZ_LAUNCH((bin2d_t_kernel<true, true, F>), grid, block, 0, s, out, ldo, rows, cols, a, lda, b, ldb, f, tiles_i, tiles_j);
    else if (ta) JZ_LAUNCH((bin2d_t_kernel<true, false, F>), grid, block, 0, s, out, ldo, rows, cols, a, lda, b, ldb, f, tiles_i, tiles_j);
    else JZ_LAUNCH((bin2d_t_kernel<false, true, F>), grid, block, 0, s, out, ldo, rows, cols, a, lda, b, ldb, f, tiles_i, tiles_j);
    return JZ_OK;
}

}  // namespace jz

using namespace jz;

extern "C" {

int jz_copy2d(float* dst, size_t ldd, const float* src, size_t lds, size_t rows, size_t cols, int trans,
              jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    if (rows == 0 || cols == 0) return JZ_OK;
    if (!dst || !src) return fail(JZ_ERR_ARG, "jz_copy2d: null pointer");
    if (ldd < rows || lds < (trans ? cols : rows)) return fail(JZ_ERR_SHAPE, "jz_copy2d: leading dimension too small");
    cudaStream_t s = as_stream(stream);
    if (!trans) {
        if (ldd == rows && lds == rows) return jz_copy(dst, src, rows * cols, stream);
        const bool vec = aligned16(dst) && aligned16(src) && ldd % 4 == 0 && lds % 4 == 0;
        return launch_map2d(Copy2dOp{dst, ldd, src, lds}, rows, cols, vec, s);
    }
    if (aligned16(dst) && aligned16(src) && ldd % 4 == 0 && lds % 4 == 0 && rows >= 64 && cols >= 64) {
        const size_t t64_i = ceil_div(rows, size_t(64)), t64_j = ceil_div(cols, size_t(64));
        // one tile per CTA (the hardware scheduler keeps every SM's queue of loads full): measured 6.51 TB/s against 6.06
        // for the best persistent grid (profiles/r01h_tune_transpose.log)
        const size_t nt = t64_i * t64_j;
        const unsigned grid = unsigned(nt < 0x7fffffffull ? nt : 0x7fffffffull);
        JZ_LAUNCH(transpose64_kernel, grid, 256, 0, s, dst, ldd, src, lds, rows, cols, t64_i, t64_j);
        return JZ_OK;
    }
    const size_t tiles_i = ceil_div(rows, kTile), tiles_j = ceil_div(cols, kTile);
    JZ_LAUNCH(transpose_kernel, tile_grid(tiles_i * tiles_j), dim3(kTile, kTileRows), 0, s, dst, ldd, src, lds, rows,
              cols, tiles_i, tiles_j);
    return JZ_OK;
}

int jz_axpby2d(float* out, size_t ldo, size_t rows, size_t cols, const float* a, size_t lda, int a_trans,
               const float* b, size_t ldb, int b_trans, float s1, float s2, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    return launch_bin2d(out, ldo, rows, cols, a, lda, a_trans, b, ldb, b_trans, AxpbyF2{s1, s2}, as_stream(stream));
}

int jz_hadamard2d(float* out, size_t ldo, size_t rows, size_t cols, const float* a, size_t lda, int a_trans,
                  const float* b, size_t ldb, int b_trans, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    return launch_bin2d(out, ldo, rows, cols, a, lda, a_trans, b, ldb, b_trans, MulF2{}, as_stream(stream));
}

int jz_add_bcast(float* out, const float* a, size_t rows, size_t cols, const float* v, int dim, float s1, float s2,
                 jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    if (rows == 0 || cols == 0) return JZ_OK;
    if (!out || !a || !v) return fail(JZ_ERR_ARG, "jz_add_bcast: null pointer");
    if (dim != 0 && dim != 1) return fail(JZ_ERR_ARG, "jz_add_bcast: dim must be 0 or 1");
    const bool vec = aligned16(out) && aligned16(a) && (dim == 0 || aligned16(v));
    return launch_map2d(BcastOp{out, a, v, rows, dim, s1, s2}, rows, cols, vec, as_stream(stream));
}

int jz_outer(float* out, size_t ldo, const float* u, size_t rows, const float* v, size_t cols, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    if (rows == 0 || cols == 0) return JZ_OK;
    if (!out || !u || !v) return fail(JZ_ERR_ARG, "jz_outer: null pointer");
    if (ldo < rows) return fail(JZ_ERR_SHAPE, "jz_outer: ldo < rows");
    const bool vec = aligned16(out) && aligned16(u) && ldo % 4 == 0;
    return launch_map2d(OuterOp{out, ldo, u, v}, rows, cols, vec, as_stream(stream));
}

}  // extern "C"
