// jz_reduce.cu -- row / column reductions (SURVEY 8a rows a11 sum, a12 max-reduce,
// a13 softmax composite, a18 norm).  Algorithmic traffic: 4 B/elem read (+4 B per output),
// column softmax 8 B/elem.
//
// The reference reaches these through cublasSgemv with a freshly filled ones vector (3 fill
// launches + gemv, cpp/cumatrix.cu:312-337), a one-thread-per-column serial functor kernel
// with stride-`rows` (uncoalesced) accesses (cpp/cumatrix.cuh:334-384) and ~10 kernels + 2
// rank-1 GEMMs for the softmax (ml/layer.hpp:252-283).  Here every variant is one or two
// coalesced streaming passes, chosen by shape:
//
//  dim 0 (reduce down each contiguous column):
//    rows <= 32      : a CTA stages 256 columns (one contiguous span) in shared memory with
//                      128-bit loads, then one thread reduces one column from a skewed,
//                      bank-conflict-free layout;
//    otherwise       : one warp per column (128-bit loads, 4 accumulators per lane, shuffle
//                      tree), or, for few very long columns, (chunk x column) CTAs writing
//                      partials that a second tiny pass folds -- no atomics, so results are
//                      deterministic for a given shape.
//  dim 1 (reduce across columns, i.e. add the columns into a vector):
//    each thread owns 4 consecutive rows (float4) and walks columns; the columns are split
//    across threadIdx.y and across CTAs (grid.y).  Every global read is a full coalesced row
//    segment.  The column chunks meet in ONE launch: the 8 CTAs of a thread-block cluster
//    (cluster dims 1 x 8) add their partial vectors through distributed shared memory, each
//    cluster writes one partial vector, and the last CTA of a row block to finish (a ticket
//    counter after __threadfence) folds the few cluster partials in fixed order -- no float
//    atomics, deterministic per shape, no second kernel (which cost 7 us: 40 % of a 4096^2
//    row sum).  Shapes with fewer than 8 column chunks keep the plain two-pass form.
#include <cooperative_groups.h>

#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <unordered_map>

#include "jz_common.cuh"
#include "jz_math.cuh"

namespace jz {

struct SumOp {
    static __device__ __forceinline__ float init() { return 0.0f; }
    static __device__ __forceinline__ float apply(float a, float b) { return __fadd_rn(a, b); }
};
struct MaxOp {  // the LogisticLayer column-max functor: m = -1e30f; m = m > v ? m : v
    static __device__ __forceinline__ float init() { return -1e30f; }
    static __device__ __forceinline__ float apply(float a, float b) { return a > b ? a : b; }
};

template <class Op>
__device__ __forceinline__ float warp_reduce(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = Op::apply(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------ dim 0, short columns
constexpr int kShortCols = 256;  // columns per CTA
constexpr int kShortMaxRows = 32;

__device__ __forceinline__ int skew(int e) { return e + (e >> 5); }

// stage `ncols` columns of `rows` floats (column c at a + c*ld) into skewed smem
__device__ __forceinline__ void stage_columns(float* sm, const float* a, size_t ld, int rows, int ncols,
                                              bool contiguous_vec) {
    const int total = rows * ncols;
    if (contiguous_vec) {  // ld == rows, base 16B aligned, span is one contiguous run
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const int total4 = total >> 2;
        for (int e4 = threadIdx.x; e4 < total4; e4 += blockDim.x) {
            const float4 v = a4[e4];
            const int e = e4 << 2;
            sm[skew(e)] = v.x; sm[skew(e + 1)] = v.y; sm[skew(e + 2)] = v.z; sm[skew(e + 3)] = v.w;
        }
        for (int e = (total4 << 2) + threadIdx.x; e < total; e += blockDim.x) sm[skew(e)] = a[e];
    } else {
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            const int c = e / rows, r = e - c * rows;
            sm[skew(e)] = a[size_t(c) * ld + r];
        }
    }
}

template <class Op>
__global__ void __launch_bounds__(kShortCols) colreduce_short_kernel(float* out, const float* a, int rows,
                                                                     size_t cols, size_t ld, bool vec) {
    pdl_enter();
    extern __shared__ float sm[];
    for (size_t c0 = size_t(blockIdx.x) * kShortCols; c0 < cols; c0 += size_t(gridDim.x) * kShortCols) {
        const int ncols = int(cols - c0 < size_t(kShortCols) ? cols - c0 : kShortCols);
        stage_columns(sm, a + c0 * ld, ld, rows, ncols, vec);
        __syncthreads();
        if (int(threadIdx.x) < ncols) {
            float acc = Op::init();
            const int base = threadIdx.x * rows;
            for (int r = 0; r < rows; r++) acc = Op::apply(acc, sm[skew(base + r)]);  // serial, reference order
            out[c0 + threadIdx.x] = acc;
        }
        __syncthreads();
    }
}

// column softmax for short columns: exact reference order inside a column
// mode 0: out = softmax ; mode 1: out = -(y - softmax)/nb evaluated as the operators do
__global__ void __launch_bounds__(kShortCols) softmax_short_kernel(float* out, const float* a, const float* y,
                                                                   int rows, size_t cols, size_t ld, bool vec,
                                                                   int mode, float rnb) {
    pdl_enter();
    extern __shared__ float sm[];
    for (size_t c0 = size_t(blockIdx.x) * kShortCols; c0 < cols; c0 += size_t(gridDim.x) * kShortCols) {
        const int ncols = int(cols - c0 < size_t(kShortCols) ? cols - c0 : kShortCols);
        stage_columns(sm, a + c0 * ld, ld, rows, ncols, vec);
        __syncthreads();
        if (int(threadIdx.x) < ncols) {
            const int base = threadIdx.x * rows;
            float m = -1e30f;
            for (int r = 0; r < rows; r++) { const float v = sm[skew(base + r)]; m = m > v ? m : v; }
            float z = 0.0f;
            for (int r = 0; r < rows; r++) {
                const float e = expf(__fadd_rn(-m, sm[skew(base + r)]));
                sm[skew(base + r)] = e;
                z = __fadd_rn(z, e);
            }
            const float inv = __fdiv_rn(1.0f, z);
            for (int r = 0; r < rows; r++) sm[skew(base + r)] = __fmul_rn(sm[skew(base + r)], inv);
        }
        __syncthreads();
        // coalesced write-back (out is contiguous rows x cols, ld == rows)
        const int total = rows * ncols;
        float* o = out + c0 * size_t(rows);
        const float* yy = y ? y + c0 * size_t(rows) : nullptr;
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            float s = sm[skew(e)];
            if (mode == 1) {
                const float d = __fadd_rn(-s, yy[e]);          // rM.add(lM,-1,1): -1*S + 1*Y
                const float neg = __fadd_rn(-d, 0.0f);         // unary minus: -1*d + 0
                s = __fadd_rn(__fmul_rn(rnb, neg), 0.0f);      // /nb: (float)(1/nb)*x + 0
            }
            o[e] = s;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ dim 0, warp per column
template <class Op, bool VEC>
__device__ __forceinline__ float lane_reduce_span(const float* col, size_t len, int lane) {
    float a0 = Op::init(), a1 = Op::init(), a2 = Op::init(), a3 = Op::init();
    if (VEC) {
        const float4* c4 = reinterpret_cast<const float4*>(col);
        const size_t n4 = len >> 2;
        size_t i = lane;
        for (; i + 96 < n4; i += 128) {
            const float4 v0 = c4[i], v1 = c4[i + 32], v2 = c4[i + 64], v3 = c4[i + 96];
            a0 = Op::apply(a0, Op::apply(Op::apply(v0.x, v0.y), Op::apply(v0.z, v0.w)));
            a1 = Op::apply(a1, Op::apply(Op::apply(v1.x, v1.y), Op::apply(v1.z, v1.w)));
            a2 = Op::apply(a2, Op::apply(Op::apply(v2.x, v2.y), Op::apply(v2.z, v2.w)));
            a3 = Op::apply(a3, Op::apply(Op::apply(v3.x, v3.y), Op::apply(v3.z, v3.w)));
        }
        for (; i < n4; i += 32) {
            const float4 v0 = c4[i];
            a0 = Op::apply(a0, Op::apply(Op::apply(v0.x, v0.y), Op::apply(v0.z, v0.w)));
        }
        for (size_t t = (n4 << 2) + lane; t < len; t += 32) a1 = Op::apply(a1, col[t]);
    } else {
        size_t i = lane;
        for (; i + 96 < len; i += 128) {
            a0 = Op::apply(a0, col[i]);
            a1 = Op::apply(a1, col[i + 32]);
            a2 = Op::apply(a2, col[i + 64]);
            a3 = Op::apply(a3, col[i + 96]);
        }
        for (; i < len; i += 32) a0 = Op::apply(a0, col[i]);
    }
    return Op::apply(Op::apply(a0, a1), Op::apply(a2, a3));
}

template <class Op, bool VEC>
__global__ void __launch_bounds__(256) colreduce_warp_kernel(float* out, const float* a, size_t rows, size_t cols,
                                                             size_t ld) {
    pdl_enter();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (size_t c = size_t(blockIdx.x) * 8 + warp; c < cols; c += size_t(gridDim.x) * 8) {
        float v = lane_reduce_span<Op, VEC>(a + c * ld, rows, lane);
        v = warp_reduce<Op>(v);
        if (lane == 0) out[c] = v;
    }
}

// few, very long columns: CTA (chunk, column) -> partial[column * nchunks + chunk]
template <class Op, bool VEC>
__global__ void __launch_bounds__(256) colreduce_chunk_kernel(float* partial, const float* a, size_t rows,
                                                              size_t ld, size_t chunk_len, unsigned nchunks) {
    pdl_enter();
    __shared__ float ws[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t c = blockIdx.y;
    const size_t begin = size_t(blockIdx.x) * chunk_len;
    const size_t end = begin + chunk_len < rows ? begin + chunk_len : rows;
    // each warp takes a contiguous eighth of the chunk (multiple of 4 elements keeps float4 alignment)
    const size_t len = end > begin ? end - begin : 0;
    size_t per = ((len + 7) / 8 + 3) & ~size_t(3);
    size_t wb = begin + size_t(warp) * per;
    size_t we = wb + per < end ? wb + per : end;
    float v = Op::init();
    if (wb < we) v = lane_reduce_span<Op, VEC>(a + c * ld + wb, we - wb, lane);
    v = warp_reduce<Op>(v);
    if (lane == 0) ws[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < 8 ? ws[lane] : Op::init();
        t = warp_reduce<Op>(t);
        if (lane == 0) partial[c * nchunks + blockIdx.x] = t;
    }
}

// ------------------------------------------------------------------ dim 1 (across columns)
// block (tx, ty); thread owns VEC rows; columns [j_begin, j_end) of this CTA's chunk are strided over ty.
constexpr int kRowCluster = 8;   // CTAs per cluster along the column-chunk axis (portable maximum)

// CLUSTER = false: writes one partial vector per column chunk (out: nchunks x rows).
// CLUSTER = true : gridDim.y is a multiple of kRowCluster; out: (nchunks / kRowCluster) x rows partial vectors, and the
//                  last CTA of each row block folds them into `final_out` (see the file header).
template <class Op, int VEC, bool CLUSTER>
__global__ void __launch_bounds__(256) rowreduce_kernel(float* out, const float* a, size_t rows, size_t cols, size_t ld,
                                                        size_t cols_per_chunk, float* final_out, unsigned* tickets) {
    pdl_enter();
    extern __shared__ float sm[];  // ty x (tx*VEC)
    const size_t row_units = rows / VEC;
    const size_t iu = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t j_begin = size_t(blockIdx.y) * cols_per_chunk;
    const size_t j_end = j_begin + cols_per_chunk < cols ? j_begin + cols_per_chunk : cols;
    float acc[VEC];
#pragma unroll
    for (int q = 0; q < VEC; q++) acc[q] = Op::init();
    if (iu < row_units) {
        const float* base = a + iu * VEC;
        size_t j = j_begin + threadIdx.y;
        const size_t step = blockDim.y;
        if constexpr (VEC == 4) {
            float4 b0 = make_float4(Op::init(), Op::init(), Op::init(), Op::init()), b1 = b0, b2 = b0, b3 = b0;
            for (; j + 3 * step < j_end; j += 4 * step) {
                const float4 v0 = *reinterpret_cast<const float4*>(base + j * ld);
                const float4 v1 = *reinterpret_cast<const float4*>(base + (j + step) * ld);
                const float4 v2 = *reinterpret_cast<const float4*>(base + (j + 2 * step) * ld);
                const float4 v3 = *reinterpret_cast<const float4*>(base + (j + 3 * step) * ld);
                b0.x = Op::apply(b0.x, v0.x); b0.y = Op::apply(b0.y, v0.y); b0.z = Op::apply(b0.z, v0.z); b0.w = Op::apply(b0.w, v0.w);
                b1.x = Op::apply(b1.x, v1.x); b1.y = Op::apply(b1.y, v1.y); b1.z = Op::apply(b1.z, v1.z); b1.w = Op::apply(b1.w, v1.w);
                b2.x = Op::apply(b2.x, v2.x); b2.y = Op::apply(b2.y, v2.y); b2.z = Op::apply(b2.z, v2.z); b2.w = Op::apply(b2.w, v2.w);
                b3.x = Op::apply(b3.x, v3.x); b3.y = Op::apply(b3.y, v3.y); b3.z = Op::apply(b3.z, v3.z); b3.w = Op::apply(b3.w, v3.w);
            }
            for (; j < j_end; j += step) {
                const float4 v0 = *reinterpret_cast<const float4*>(base + j * ld);
                b0.x = Op::apply(b0.x, v0.x); b0.y = Op::apply(b0.y, v0.y); b0.z = Op::apply(b0.z, v0.z); b0.w = Op::apply(b0.w, v0.w);
            }
            acc[0] = Op::apply(Op::apply(b0.x, b1.x), Op::apply(b2.x, b3.x));
            acc[1 % VEC] = Op::apply(Op::apply(b0.y, b1.y), Op::apply(b2.y, b3.y));
            acc[2 % VEC] = Op::apply(Op::apply(b0.z, b1.z), Op::apply(b2.z, b3.z));
            acc[3 % VEC] = Op::apply(Op::apply(b0.w, b1.w), Op::apply(b2.w, b3.w));
        } else {
            float b0 = Op::init(), b1 = b0, b2 = b0, b3 = b0;
            for (; j + 3 * step < j_end; j += 4 * step) {
                b0 = Op::apply(b0, base[j * ld]);
                b1 = Op::apply(b1, base[(j + step) * ld]);
                b2 = Op::apply(b2, base[(j + 2 * step) * ld]);
                b3 = Op::apply(b3, base[(j + 3 * step) * ld]);
            }
            for (; j < j_end; j += step) b0 = Op::apply(b0, base[j * ld]);
            acc[0] = Op::apply(Op::apply(b0, b1), Op::apply(b2, b3));
        }
    }
    // fold threadIdx.y in shared memory
    const int width = blockDim.x * VEC;
#pragma unroll
    for (int q = 0; q < VEC; q++) sm[threadIdx.y * width + threadIdx.x * VEC + q] = acc[q];
    __syncthreads();
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthreads = blockDim.x * blockDim.y;
    if constexpr (!CLUSTER) {
        for (int e = tid; e < width; e += nthreads) {
            float t = Op::init();
            for (int yy = 0; yy < int(blockDim.y); yy++) t = Op::apply(t, sm[yy * width + e]);
            const size_t row = size_t(blockIdx.x) * width + e;
            if (row < rows) out[size_t(blockIdx.y) * rows + row] = t;
        }
    } else {
        namespace cg = cooperative_groups;
        cg::cluster_group cluster = cg::this_cluster();
        // this CTA's vector, folded over threadIdx.y, left in sm[0 .. width) (column e is touched by one thread only)
        for (int e = tid; e < width; e += nthreads) {
            float t = Op::init();
            for (int yy = 0; yy < int(blockDim.y); yy++) t = Op::apply(t, sm[yy * width + e]);
            sm[e] = t;
        }
        cluster.sync();
        // CTA r of the cluster adds slice r of the 8 vectors through distributed shared memory, fixed order
        const unsigned r = cluster.block_rank();
        const int per = (width + kRowCluster - 1) / kRowCluster;
        const size_t cchunk = blockIdx.y / kRowCluster;
        for (int e = int(r) * per + tid; e < int(r + 1) * per && e < width; e += nthreads) {
            float t = Op::init();
#pragma unroll
            for (int q = 0; q < kRowCluster; q++) t = Op::apply(t, cluster.map_shared_rank(sm, q)[e]);
            const size_t row = size_t(blockIdx.x) * width + e;
            if (row < rows) out[cchunk * rows + row] = t;
        }
        cluster.sync();   // nobody exits while its shared memory may still be read
        // ticket: the last CTA of this row block folds the cluster partials
        __shared__ bool last;
        __threadfence();
        __syncthreads();
        if (tid == 0) last = atomicAdd(&tickets[blockIdx.x], 1u) == gridDim.y - 1;
        __syncthreads();
        if (last) {
            __threadfence();
            const size_t nclusters = gridDim.y / kRowCluster;
            for (int e = tid; e < width; e += nthreads) {
                const size_t row = size_t(blockIdx.x) * width + e;
                if (row >= rows) continue;
                float t0 = Op::init(), t1 = t0, t2 = t0, t3 = t0;
                size_t c = 0;
                for (; c + 4 <= nclusters; c += 4) {
                    t0 = Op::apply(t0, __ldcg(out + c * rows + row));
                    t1 = Op::apply(t1, __ldcg(out + (c + 1) * rows + row));
                    t2 = Op::apply(t2, __ldcg(out + (c + 2) * rows + row));
                    t3 = Op::apply(t3, __ldcg(out + (c + 3) * rows + row));
                }
                for (; c < nclusters; c++) t0 = Op::apply(t0, __ldcg(out + c * rows + row));
                final_out[row] = Op::apply(Op::apply(t0, t1), Op::apply(t2, t3));
            }
            if (tid == 0) tickets[blockIdx.x] = 0;   // ready for the next launch on this stream
        }
    }
}

// ------------------------------------------------------------------ long-column softmax (warp per column)
template <bool VEC>
__global__ void __launch_bounds__(256) softmax_warp_kernel(float* out, const float* a, size_t rows, size_t cols,
                                                           size_t ld) {
    pdl_enter();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (size_t c = size_t(blockIdx.x) * 8 + warp; c < cols; c += size_t(gridDim.x) * 8) {
        const float* col = a + c * ld;
        float* o = out + c * rows;
        float m = warp_reduce<MaxOp>(lane_reduce_span<MaxOp, VEC>(col, rows, lane));
        float z = 0.0f;
        for (size_t i = lane; i < rows; i += 32) {
            const float e = expf(__fadd_rn(-m, col[i]));
            o[i] = e;
            z += e;
        }
        z = warp_reduce<SumOp>(z);
        const float inv = __fdiv_rn(1.0f, z);
        for (size_t i = lane; i < rows; i += 32) o[i] = __fmul_rn(o[i], inv);  // same lane re-reads its own writes
    }
}

// 16-byte asynchronous global -> shared copy (LDGSTS, L2 only) and the wait for this thread's own copies
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

// ---- register-cached column softmax: the column is read ONCE (128-bit loads), kept in registers
// through max / exp / sum / scale, and written ONCE: 8 B/elem of HBM traffic, the algorithmic
// minimum.  G threads cooperate on one column: a warp (G = 32, 8 columns per CTA) for rows up to
// 2048, a whole 512-thread CTA (G = 512) for rows up to 32768.  mode 1 fuses the softmax-CE
// gradient -(y - s)/nb (ml/layer.hpp:263) into the same pass.
// PREF (long columns, one CTA per SM): the phases of a column serialise inside the CTA (load everything, reduce,
// exponentiate, reduce, store), so HBM idles while the CTA computes.  With PREF every thread copies ITS elements of
// the NEXT column into a private slice of shared memory with cp.async while it works on the current one from
// registers, and starts the next column from shared memory: loads of column c+1 overlap the math and the stores of
// column c (no extra barrier: a thread only ever reads back what it copied itself).
template <int NV, bool BLOCK, bool PREF = false>
__global__ void __launch_bounds__(BLOCK ? 512 : 256)
softmax_reg_kernel(float* out, const float* a, const float* y, size_t rows, size_t cols, size_t ld, int mode, float rnb) {
    pdl_enter();
    constexpr int G = BLOCK ? 512 : 32;
    static_assert(!PREF || BLOCK, "prefetch staging is for the CTA-per-column form");
    extern __shared__ float4 sm_stage[];   // PREF: [NV][512]
    __shared__ float red[16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = BLOCK ? threadIdx.x : lane;
    const size_t n4 = rows >> 2;
    const size_t col_step = BLOCK ? gridDim.x : size_t(gridDim.x) * 8;
    auto prefetch = [&](size_t c) {
        const float4* col = reinterpret_cast<const float4*>(a + c * ld);
#pragma unroll
        for (int q = 0; q < NV; q++) {
            const size_t i = size_t(g) + size_t(q) * G;
            if (i < n4) cp_async16(&sm_stage[q * G + g], col + i);
        }
    };
    if (PREF && blockIdx.x < cols) prefetch(blockIdx.x);
    for (size_t c = BLOCK ? blockIdx.x : size_t(blockIdx.x) * 8 + warp; c < cols; c += col_step) {
        const float4* col = reinterpret_cast<const float4*>(a + c * ld);
        float4 v[NV];
        float m = -1e30f;
        if (PREF) cp_async_wait_all();
#pragma unroll
        for (int q = 0; q < NV; q++) {
            const size_t i = size_t(g) + size_t(q) * G;
            if (i < n4) {
                v[q] = PREF ? sm_stage[q * G + g] : col[i];
                m = fmaxf(m, fmaxf(fmaxf(v[q].x, v[q].y), fmaxf(v[q].z, v[q].w)));
            }
        }
        if (PREF && c + col_step < cols) prefetch(c + col_step);
        m = warp_reduce<MaxOp>(m);
        if (BLOCK) {
            __syncthreads();  // protects `red` from the previous column's readers
            if (lane == 0) red[warp] = m;
            __syncthreads();
            m = red[lane & 15];
            m = warp_reduce<MaxOp>(m);
        }
        float z = 0.0f;
#pragma unroll
        for (int q = 0; q < NV; q++) {
            const size_t i = size_t(g) + size_t(q) * G;
            if (i < n4) {
                v[q].x = expf(__fadd_rn(-m, v[q].x)); v[q].y = expf(__fadd_rn(-m, v[q].y));
                v[q].z = expf(__fadd_rn(-m, v[q].z)); v[q].w = expf(__fadd_rn(-m, v[q].w));
                z += (v[q].x + v[q].y) + (v[q].z + v[q].w);
            }
        }
        z = warp_reduce<SumOp>(z);
        if (BLOCK) {
            __syncthreads();
            if (lane == 0) red[warp] = z;
            __syncthreads();
            z = red[lane & 15];
            // 16 partials duplicated over 32 lanes: reduce over 16 lanes only
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
        }
        const float inv = __fdiv_rn(1.0f, z);
        float4* o4 = reinterpret_cast<float4*>(out + c * rows);
        const float4* y4 = reinterpret_cast<const float4*>(mode == 1 ? y + c * rows : nullptr);
#pragma unroll
        for (int q = 0; q < NV; q++) {
            const size_t i = size_t(g) + size_t(q) * G;
            if (i < n4) {
                float4 r;
                r.x = __fmul_rn(v[q].x, inv); r.y = __fmul_rn(v[q].y, inv);
                r.z = __fmul_rn(v[q].z, inv); r.w = __fmul_rn(v[q].w, inv);
                if (mode == 1) {
                    const float4 t = y4[i];
                    r.x = __fadd_rn(__fmul_rn(rnb, __fadd_rn(-__fadd_rn(-r.x, t.x), 0.0f)), 0.0f);
                    r.y = __fadd_rn(__fmul_rn(rnb, __fadd_rn(-__fadd_rn(-r.y, t.y), 0.0f)), 0.0f);
                    r.z = __fadd_rn(__fmul_rn(rnb, __fadd_rn(-__fadd_rn(-r.z, t.z), 0.0f)), 0.0f);
                    r.w = __fadd_rn(__fmul_rn(rnb, __fadd_rn(-__fadd_rn(-r.w, t.w), 0.0f)), 0.0f);
                }
                o4[i] = r;
            }
        }
    }
}

struct CeGradF {
    float rnb;
    __device__ __forceinline__ float operator()(float s, float y) const {
        const float d = __fadd_rn(-s, y);
        const float neg = __fadd_rn(-d, 0.0f);
        return __fadd_rn(__fmul_rn(rnb, neg), 0.0f);
    }
};
template <class F>
__global__ void __launch_bounds__(256) map2_inplace_kernel(float* out, const float* b, size_t n, F f) {
    pdl_enter();
    for (size_t i = size_t(blockIdx.x) * 256 + threadIdx.x; i < n; i += size_t(gridDim.x) * 256) out[i] = f(out[i], b[i]);
}

// ------------------------------------------------------------------ nrm2
__global__ void __launch_bounds__(256) sumsq_partial_kernel(double* partial, const float* x, size_t n, bool vec) {
    pdl_enter();
    __shared__ double ws[8];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    double acc = 0.0;
    const size_t stride = size_t(gridDim.x) * 256;
    if (vec) {
        const float4* x4 = reinterpret_cast<const float4*>(x);
        const size_t n4 = n >> 2;
        int flush = 0;
        for (size_t i = size_t(blockIdx.x) * 256 + threadIdx.x; i < n4; i += stride) {
            const float4 v = x4[i];
            a0 = fmaf(v.x, v.x, a0); a1 = fmaf(v.y, v.y, a1); a2 = fmaf(v.z, v.z, a2); a3 = fmaf(v.w, v.w, a3);
            if (++flush == 64) { acc += double(a0) + double(a1) + double(a2) + double(a3); a0 = a1 = a2 = a3 = 0.f; flush = 0; }
        }
        if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float t = x[(n4 << 2) + threadIdx.x]; a0 = fmaf(t, t, a0); }
    } else {
        int flush = 0;
        for (size_t i = size_t(blockIdx.x) * 256 + threadIdx.x; i < n; i += stride) {
            const float t = x[i];
            a0 = fmaf(t, t, a0);
            if (++flush == 256) { acc += double(a0); a0 = 0.f; flush = 0; }
        }
    }
    acc += double(a0) + double(a1) + double(a2) + double(a3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += ws[w];
        partial[blockIdx.x] = t;
    }
}

__global__ void nrm2_final_kernel(float* out, const double* partial, int n) {
    pdl_enter();
    __shared__ double ws[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < int(blockDim.x >> 5); w++) t += ws[w];
        out[0] = float(sqrt(t));
    }
}

// ------------------------------------------------------------------ host dispatch
template <class Op>
static int reduce_dim1(float* out, const float* a, size_t rows, size_t cols, size_t ld, cudaStream_t s);

// grid policy knobs (JZ_REDUCE_CAP / JZ_REDUCE_WAVES override them for tuning runs): the streaming kernels
// measured fastest with one work item per CTA rather than a persistent grid (profiles/r01g_tune_stream.log)
static size_t stream_cap() {
    static const size_t v = [] {
        const char* e = std::getenv("JZ_REDUCE_CAP");
        return e ? size_t(std::atoll(e)) * size_t(ctx().sm_count) : size_t(0x7fffffff);
    }();
    return v;
}
// CTAs per SM-slot for the dim-1 column split.  Measured on B200 (profiles/r01w_rowreduce_tuning.log): one wave is
// best up to 2^26 elements (fewer, longer chunks: 0.68 / 0.83 of the copy peak at 2^24 / 2^26 against 0.51 / 0.72
// with two), two waves at 2^28 (0.97 against 0.93).  JZ_REDUCE_WAVES overrides.
static size_t row_waves(size_t elems) {
    static const size_t forced = [] {
        const char* e = std::getenv("JZ_REDUCE_WAVES");
        return e ? size_t(std::atoll(e)) : size_t(0);
    }();
    if (forced) return forced;
    return elems >= (size_t(1) << 27) ? 2 : 1;
}

template <class Op>
static int reduce_dim0(float* out, const float* a, size_t rows, size_t cols, size_t ld, cudaStream_t s) {
    const size_t cap = size_t(ctx().sm_count) * 8;
    if (rows <= size_t(kShortMaxRows)) {
        const bool vec = (ld == rows) && aligned16(a);
        const size_t blocks = ceil_div(cols, size_t(kShortCols));
        const size_t bytes = (size_t(kShortCols) * rows + (size_t(kShortCols) * rows) / 32 + 2) * sizeof(float);
        JZ_LAUNCH((colreduce_short_kernel<Op>), unsigned(blocks < cap ? blocks : cap), kShortCols, bytes, s, out, a,
                  int(rows), cols, ld, vec);
        return JZ_OK;
    }
    const bool vec = aligned16(a) && ld % 4 == 0;
    const size_t warps_wanted = cap * 8;  // one full wave of warps
    if (cols * 4 >= warps_wanted || rows < 32768) {
        const size_t blocks = ceil_div(cols, size_t(8));
        const size_t lim = stream_cap();
        const unsigned grid = unsigned(blocks < lim ? blocks : lim);
        if (vec) JZ_LAUNCH((colreduce_warp_kernel<Op, true>), grid, 256, 0, s, out, a, rows, cols, ld);
        else JZ_LAUNCH((colreduce_warp_kernel<Op, false>), grid, 256, 0, s, out, a, rows, cols, ld);
        return JZ_OK;
    }
    // few long columns: split rows into chunks so that ~2 waves of CTAs exist
    size_t nchunks = ceil_div(2 * cap, cols);
    size_t chunk_len = ceil_div(rows, nchunks);
    if (chunk_len < 8192) chunk_len = 8192;
    chunk_len = (chunk_len + 31) & ~size_t(31);
    nchunks = ceil_div(rows, chunk_len);
    if (cols > 65535) return fail(JZ_ERR_UNSUPPORTED, "reduce: too many long columns");
    WsGuard wg(s);
    int rc = ws_alloc(&wg.p, nchunks * cols * sizeof(float), s);
    if (rc != JZ_OK) return rc;
    float* pf = static_cast<float*>(wg.p);
    const dim3 grid((unsigned)nchunks, (unsigned)cols, 1);
    if (vec) JZ_LAUNCH((colreduce_chunk_kernel<Op, true>), grid, 256, 0, s, pf, a, rows, ld, chunk_len, unsigned(nchunks));
    else JZ_LAUNCH((colreduce_chunk_kernel<Op, false>), grid, 256, 0, s, pf, a, rows, ld, chunk_len, unsigned(nchunks));
    // fold: partial is an (nchunks x cols) col-major matrix -> reduce down its columns
    return reduce_dim0<Op>(out, pf, nchunks, cols, nchunks, s);
}

template <class Op>
static int reduce_dim1(float* out, const float* a, size_t rows, size_t cols, size_t ld, cudaStream_t s) {
    const size_t cap = size_t(ctx().sm_count) * 8;
    const bool vec = aligned16(a) && ld % 4 == 0 && rows % 4 == 0;
    const int V = vec ? 4 : 1;
    const size_t row_units = rows / V;
    unsigned tx = 1;
    while (tx < 256 && tx < row_units) tx <<= 1;
    if (tx > 64 && vec) tx = 64;  // keep >= 4 column phases per CTA for MLP
    const unsigned ty = 256 / tx;
    const size_t gx = ceil_div(row_units, tx);
    // split columns into chunks so the grid has ~2 waves, but keep >= 8*ty columns per chunk
    size_t nchunks = ceil_div(row_waves(rows * cols) * cap, gx);
    const size_t min_cols = size_t(ty) * 8;
    if (nchunks * min_cols > cols) nchunks = cols / min_cols;
    if (nchunks < 1) nchunks = 1;
    if (nchunks > 65535) nchunks = 65535;
    size_t cpc = ceil_div(cols, nchunks);
    nchunks = ceil_div(cols, cpc);
    const size_t smem = size_t(ty) * tx * V * sizeof(float);
    const dim3 block(tx, ty, 1);
    static const bool no_cluster = std::getenv("JZ_REDUCE_NO_CLUSTER") != nullptr;
    if (nchunks >= kRowCluster && gx <= kTicketSlots && !no_cluster) {
        // one launch: clusters of 8 column chunks + last-CTA fold.  Round the chunk count up to a multiple of 8
        // (trailing chunks past the last column contribute Op::init()).
        const size_t nch = ceil_div(nchunks, size_t(kRowCluster)) * kRowCluster;
        cpc = ceil_div(cols, nch);
        unsigned* tickets = tickets_for(s);
        if (!tickets) return fail(JZ_ERR_CUDA, "reduce: could not allocate the ticket counters");
        WsGuard wg(s);
        int rc = ws_alloc(&wg.p, (nch / kRowCluster) * rows * sizeof(float), s);
        if (rc != JZ_OK) return rc;
        void* partial = wg.p;
        cudaLaunchConfig_t cfg;
        std::memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)gx, (unsigned)nch, 1);
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1;
        attr[0].val.clusterDim.y = kRowCluster;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        float* part = static_cast<float*>(partial);
        cudaError_t e = vec ? cudaLaunchKernelEx(&cfg, rowreduce_kernel<Op, 4, true>, part, a, rows, cols, ld, cpc, out, tickets)
                            : cudaLaunchKernelEx(&cfg, rowreduce_kernel<Op, 1, true>, part, a, rows, cols, ld, cpc, out, tickets);
        ctx().launches.fetch_add(1, std::memory_order_relaxed);
        if (e != cudaSuccess) return cuda_fail(e, "rowreduce_kernel (cluster) launch");
        return JZ_OK;
    }
    const dim3 grid((unsigned)gx, (unsigned)nchunks, 1);
    float* dst = out;
    WsGuard wg(s);
    if (nchunks > 1) {
        int rc = ws_alloc(&wg.p, nchunks * rows * sizeof(float), s);
        if (rc != JZ_OK) return rc;
        dst = static_cast<float*>(wg.p);
    }
    if (vec) JZ_LAUNCH((rowreduce_kernel<Op, 4, false>), grid, block, smem, s, dst, a, rows, cols, ld, cpc, nullptr, nullptr);
    else JZ_LAUNCH((rowreduce_kernel<Op, 1, false>), grid, block, smem, s, dst, a, rows, cols, ld, cpc, nullptr, nullptr);
    if (nchunks > 1) {
        // partial is rows x nchunks (col-major, ld = rows): fold across its columns
        return reduce_dim1<Op>(out, dst, rows, nchunks, rows, s);
    }
    return JZ_OK;
}

template <class Op>
static int reduce_entry(float* out, const float* a, size_t rows, size_t cols, size_t ld, int dim, cudaStream_t s) {
    if (dim != 0 && dim != 1) return fail(JZ_ERR_ARG, "reduce: dim must be 0 or 1");
    if (ld < rows) return fail(JZ_ERR_SHAPE, "reduce: ld < rows");
    const size_t nout = dim == 0 ? cols : rows;
    if (nout == 0) return JZ_OK;
    if (!out || (!a && rows * cols)) return fail(JZ_ERR_ARG, "reduce: null pointer");
    if ((dim == 0 ? rows : cols) == 0) {  // empty reduction: identity
        float init = 0.0f;
        if (std::is_same<Op, MaxOp>::value) init = -1e30f;
        return jz_fill(out, nout, init, s);
    }
    return dim == 0 ? reduce_dim0<Op>(out, a, rows, cols, ld, s) : reduce_dim1<Op>(out, a, rows, cols, ld, s);
}

// Columns too long for one CTA's registers: the CL CTAs of a thread-block cluster hold one column between them (512
// threads x NV float4 each), exchange their partial max and partial sum through distributed shared memory, and the
// column is still read once and written once (rows > 32768 used to fall back to the three-pass warp kernel, 20 B per
// element).  Up to 32768 rows one 512-thread CTA per column stays faster: two cluster barriers per column cost more
// than they save there (measured at 32768^2: 0.75 against 0.80 of the copy peak).
template <int NV, int CL>
__global__ void __launch_bounds__(512)
softmax_cluster_kernel(float* out, const float* a, const float* y, size_t rows, size_t cols, size_t ld, int mode, float rnb) {
    pdl_enter();
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float red[16];
    __shared__ float xchg[2];   // this CTA's partial max / partial sum, read by the other CTAs of the cluster
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned r = cluster.block_rank();
    const size_t n4 = rows >> 2;
    const size_t seg = (n4 + CL - 1) / CL;                  // float4 words per CTA
    const size_t lo = size_t(r) * seg, hi = lo + seg < n4 ? lo + seg : n4;
    const size_t ncl = gridDim.x / CL;
    for (size_t c = blockIdx.x / CL; c < cols; c += ncl) {
        const float4* col = reinterpret_cast<const float4*>(a + c * ld);
        float4 v[NV];
        float m = -1e30f;
#pragma unroll
        for (int q = 0; q < NV; q++) {
            const size_t i = lo + threadIdx.x + size_t(q) * 512;
            if (i < hi) {
                v[q] = col[i];
                m = fmaxf(m, fmaxf(fmaxf(v[q].x, v[q].y), fmaxf(v[q].z, v[q].w)));
            }
        }
        m = warp_reduce<MaxOp>(m);
        if (lane == 0) red[warp] = m;
        __syncthreads();
        m = warp_reduce<MaxOp>(red[lane & 15]);
        if (threadIdx.x == 0) xchg[0] = m;
        cluster.sync();
#pragma unroll
        for (int q = 0; q < CL; q++) m = fmaxf(m, cluster.map_shared_rank(xchg, q)[0]);
        float z = 0.0f;
#pragma unroll
        for (int q = 0; q < NV; q++) {
            const size_t i = lo + threadIdx.x + size_t(q) * 512;
            if (i < hi) {
                v[q].x = expf(__fadd_rn(-m, v[q].x)); v[q].y = expf(__fadd_rn(-m, v[q].y));
                v[q].z = expf(__fadd_rn(-m, v[q].z)); v[q].w = expf(__fadd_rn(-m, v[q].w));
                z += (v[q].x + v[q].y) + (v[q].z + v[q].w);
            }
        }
        z = warp_reduce<SumOp>(z);
        __syncthreads();   // `red` readers of the max are done
        if (lane == 0) red[warp] = z;
        __syncthreads();
        z = red[lane & 15];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
        if (threadIdx.x == 0) xchg[1] = z;
        cluster.sync();
        z = 0.0f;
#pragma unroll
        for (int q = 0; q < CL; q++) z += cluster.map_shared_rank(xchg, q)[1];   // same order in every CTA
        const float inv = __fdiv_rn(1.0f, z);
        float4* o4 = reinterpret_cast<float4*>(out + c * rows);
        const float4* y4 = reinterpret_cast<const float4*>(mode == 1 ? y + c * rows : nullptr);
#pragma unroll
        for (int q = 0; q < NV; q++) {
            const size_t i = lo + threadIdx.x + size_t(q) * 512;
            if (i < hi) {
                float4 t;
                t.x = __fmul_rn(v[q].x, inv); t.y = __fmul_rn(v[q].y, inv);
                t.z = __fmul_rn(v[q].z, inv); t.w = __fmul_rn(v[q].w, inv);
                if (mode == 1) {
                    const float4 u = y4[i];
                    t.x = __fadd_rn(__fmul_rn(rnb, __fadd_rn(-__fadd_rn(-t.x, u.x), 0.0f)), 0.0f);
                    t.y = __fadd_rn(__fmul_rn(rnb, __fadd_rn(-__fadd_rn(-t.y, u.y), 0.0f)), 0.0f);
                    t.z = __fadd_rn(__fmul_rn(rnb, __fadd_rn(-__fadd_rn(-t.z, u.z), 0.0f)), 0.0f);
                    t.w = __fadd_rn(__fmul_rn(rnb, __fadd_rn(-__fadd_rn(-t.w, u.w), 0.0f)), 0.0f);
                }
                o4[i] = t;
            }
        }
        cluster.sync();   // nobody overwrites xchg for the next column while a peer may still read it
    }
}

// ---- long columns (rows > 16384; JZ_SOFTMAX_LONG_MIN moves the threshold): a column is split into chunks of kLongChunk
// elements, one CTA each, that meet
// through global memory.  Every CTA keeps its chunk in registers (read ONCE), publishes its partial (max m_c,
// z_c = sum exp(x - m_c)), takes a ticket on the column's counter and waits until all chunks of the column have
// published; then M = max m_c, Z = sum z_c * exp(m_c - M) (chunk order: identical in every CTA) and the chunk is
// written ONCE as exp(x - m_c) * (exp(m_c - M) / Z).  8 B/elem of HBM traffic for any column length, no cluster
// barriers, and the whole chip is busy whatever the column count.  The chunks of a column sit in consecutive blocks of
// one launch and far fewer of them exist than CTAs are resident, so under in-order block dispatch the wait cannot
// deadlock (same argument as the split-K units of jz_gemm_tc.cuh).
constexpr int kLongChunk = 8192;   // elements per CTA: 8 float4 per thread

__global__ void __launch_bounds__(256) softmax_chunks_kernel(float* out, const float* a, const float* y, float2* part, unsigned* tickets,
                                                             size_t rows, size_t ld, unsigned nchunks, int mode, float rnb) {
    pdl_enter();
    __shared__ float red[8];
    __shared__ float s_scale;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned chunk = blockIdx.x % nchunks;
    const size_t c = blockIdx.x / nchunks, r0 = size_t(chunk) * kLongChunk;
    const size_t n4 = (rows - r0 < size_t(kLongChunk) ? rows - r0 : size_t(kLongChunk)) >> 2;
    const float4* col = reinterpret_cast<const float4*>(a + c * ld + r0);
    float4 v[8];
    float m = -1e30f;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const size_t i = threadIdx.x + size_t(q) * 256;
        if (i < n4) {
            v[q] = col[i];
            m = fmaxf(m, fmaxf(fmaxf(v[q].x, v[q].y), fmaxf(v[q].z, v[q].w)));
        }
    }
    m = warp_reduce<MaxOp>(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = warp_reduce<MaxOp>(red[lane & 7]);
    float z = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const size_t i = threadIdx.x + size_t(q) * 256;
        if (i < n4) {
            v[q].x = expf(__fadd_rn(-m, v[q].x)); v[q].y = expf(__fadd_rn(-m, v[q].y));
            v[q].z = expf(__fadd_rn(-m, v[q].z)); v[q].w = expf(__fadd_rn(-m, v[q].w));
            z += (v[q].x + v[q].y) + (v[q].z + v[q].w);
        }
    }
    z = warp_reduce<SumOp>(z);
    __syncthreads();
    if (lane == 0) red[warp] = z;
    __syncthreads();
    float2* const pc = part + c * nchunks;
    if (threadIdx.x == 0) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; w++) t += red[w];
        __stcg(pc + chunk, make_float2(m, t));
        __threadfence();
        atomicAdd(tickets + c, 1u);
        unsigned seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(tickets + c) : "memory");
            if (seen < nchunks) __nanosleep(40);
        } while (seen < nchunks);
    }
    __syncthreads();
    // combine the column's partials with the whole CTA: thread q takes chunk q (nchunks <= 256), two block reductions in
    // a fixed tree, so every CTA of the column gets the same M and Z.  (One thread walking the list paid an L2 round
    // trip per chunk, twice: 32 chunks at 262144 rows = ~10 us per CTA with its registers parked; 0.65 of the copy peak.)
    {
        const bool have = threadIdx.x < nchunks;
        const float2 p = have ? __ldcg(pc + threadIdx.x) : make_float2(-1e30f, 0.0f);
        float M = warp_reduce<MaxOp>(p.x);
        if (lane == 0) red[warp] = M;
        __syncthreads();
        M = warp_reduce<MaxOp>(red[lane & 7]);
        float Z = have ? __fmul_rn(p.y, expf(__fadd_rn(-M, p.x))) : 0.0f;
        Z = warp_reduce<SumOp>(Z);
        __syncthreads();
        if (lane == 0) red[warp] = Z;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.0f;
#pragma unroll
            for (int w = 0; w < 8; w++) t += red[w];
            s_scale = __fmul_rn(expf(__fadd_rn(-M, m)), __fdiv_rn(1.0f, t));
        }
    }
    __syncthreads();
    const float scale = s_scale;
    float4* o4 = reinterpret_cast<float4*>(out + c * rows + r0);
    const float4* y4 = reinterpret_cast<const float4*>(mode == 1 ? y + c * rows + r0 : nullptr);
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const size_t i = threadIdx.x + size_t(q) * 256;
        if (i < n4) {
            float4 t;
            t.x = __fmul_rn(v[q].x, scale); t.y = __fmul_rn(v[q].y, scale);
            t.z = __fmul_rn(v[q].z, scale); t.w = __fmul_rn(v[q].w, scale);
            if (mode == 1) {
                const float4 u = y4[i];
                t.x = __fadd_rn(__fmul_rn(rnb, __fadd_rn(-__fadd_rn(-t.x, u.x), 0.0f)), 0.0f);
                t.y = __fadd_rn(__fmul_rn(rnb, __fadd_rn(-__fadd_rn(-t.y, u.y), 0.0f)), 0.0f);
                t.z = __fadd_rn(__fmul_rn(rnb, __fadd_rn(-__fadd_rn(-t.z, u.z), 0.0f)), 0.0f);
                t.w = __fadd_rn(__fmul_rn(rnb, __fadd_rn(-__fadd_rn(-t.w, u.w), 0.0f)), 0.0f);
            }
            o4[i] = t;
        }
    }
}

static bool nchunks_ok(size_t rows, size_t cols) {
    const size_t nchunks = ceil_div(rows, size_t(kLongChunk));
    return nchunks <= 256 && nchunks * cols < (size_t(1) << 31);   // a column's CTAs must all fit on the chip at once
}

static int launch_softmax_long(float* out, const float* a, const float* y, size_t rows, size_t cols, size_t ld, int mode, float rnb,
                               cudaStream_t s) {
    const size_t nchunks = ceil_div(rows, size_t(kLongChunk));
    const size_t ticket_bytes = (cols * sizeof(unsigned) + 511) & ~size_t(511);
    WsGuard wg(s);
    int rc = ws_alloc(&wg.p, ticket_bytes + cols * nchunks * sizeof(float2), s);
    if (rc != JZ_OK) return rc;
    unsigned* tickets = static_cast<unsigned*>(wg.p);
    float2* part = reinterpret_cast<float2*>(static_cast<char*>(wg.p) + ticket_bytes);
    JZ_CUDA(cudaMemsetAsync(tickets, 0, ticket_bytes, s));
    // (equal shares of the column instead of fixed 8192-element chunks with a short last one measured slower, 0.84 against
    // 0.94 of the copy peak at 65536 rows where both give the same chunks: the runtime chunk length costs the address arithmetic)
    JZ_LAUNCH(softmax_chunks_kernel, unsigned(nchunks * cols), 256, 0, s, out, a, y, part, tickets, rows, ld, unsigned(nchunks), mode, rnb);
    return JZ_OK;
}

template <int NV, int CL>
static int launch_softmax_cluster(float* out, const float* a, const float* y, size_t rows, size_t cols, size_t ld, int mode,
                                  float rnb, cudaStream_t s) {
    // (a cp.async next-column prefetch like softmax_reg_kernel's was tried here: the staging memory halves the resident
    // CTAs and the kernel got slower, 0.46 -> 0.32 of the copy peak at 65536 rows, profiles/r02b_long_column_softmax.log)
    const size_t slots = size_t(ctx().sm_count) * 2 / CL;   // ~2 CTAs per SM
    const size_t ncl = cols < slots ? cols : slots;
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(unsigned(ncl * CL), 1, 1);
    cfg.blockDim = dim3(512, 1, 1);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, softmax_cluster_kernel<NV, CL>, out, a, y, rows, cols, ld, mode, rnb);
    ctx().launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return cuda_fail(e, "softmax_cluster_kernel launch");
    return JZ_OK;
}

}  // namespace jz

using namespace jz;

extern "C" {

int jz_sum(float* out, const float* a, size_t rows, size_t cols, size_t ld, int dim, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    return reduce_entry<SumOp>(out, a, rows, cols, ld, dim, as_stream(stream));
}

int jz_max(float* out, const float* a, size_t rows, size_t cols, size_t ld, int dim, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    return reduce_entry<MaxOp>(out, a, rows, cols, ld, dim, as_stream(stream));
}

static int softmax_impl(float* out, const float* a, const float* y, size_t rows, size_t cols, size_t ld, int mode,
                        float rnb, cudaStream_t s) {
    if (rows == 0 || cols == 0) return JZ_OK;
    if (!out || !a || (mode == 1 && !y)) return fail(JZ_ERR_ARG, "softmax: null pointer");
    if (ld < rows) return fail(JZ_ERR_SHAPE, "softmax: ld < rows");
    const size_t cap = size_t(ctx().sm_count) * 8;
    if (rows <= size_t(kShortMaxRows)) {
        const bool vec = (ld == rows) && aligned16(a);
        const size_t blocks = ceil_div(cols, size_t(kShortCols));
        const size_t bytes = (size_t(kShortCols) * rows + (size_t(kShortCols) * rows) / 32 + 2) * sizeof(float);
        JZ_LAUNCH(softmax_short_kernel, unsigned(blocks < cap ? blocks : cap), kShortCols, bytes, s, out, a, y, int(rows),
                  cols, ld, vec, mode, rnb);
        return JZ_OK;
    }
    const bool vec = aligned16(a) && ld % 4 == 0;
    const bool regs_base = vec && rows % 4 == 0 && aligned16(out) && (mode == 0 || aligned16(y));
    static const bool no_cluster_sm = std::getenv("JZ_SOFTMAX_NO_CLUSTER") != nullptr;
    static const bool use_cluster_sm = std::getenv("JZ_SOFTMAX_CLUSTER") != nullptr;   // the older DSMEM form, for comparison
    static const size_t long_min = [] {
        const char* e = std::getenv("JZ_SOFTMAX_LONG_MIN");
        return e && *e ? size_t(std::atoll(e)) : size_t(16384);
    }();
    if (regs_base && rows > long_min && !use_cluster_sm && !no_cluster_sm && nchunks_ok(rows, cols))
        return launch_softmax_long(out, a, y, rows, cols, ld, mode, rnb, s);
    if (regs_base && rows > 32768 && rows <= 262144 && !no_cluster_sm) {   // column shared by a cluster (see the kernel)
        const size_t n4 = rows >> 2;
        if (n4 <= 8192) return launch_softmax_cluster<8, 2>(out, a, y, rows, cols, ld, mode, rnb, s);
        if (n4 <= 16384) return launch_softmax_cluster<8, 4>(out, a, y, rows, cols, ld, mode, rnb, s);
        if (n4 <= 32768) return launch_softmax_cluster<8, 8>(out, a, y, rows, cols, ld, mode, rnb, s);
        return launch_softmax_cluster<16, 8>(out, a, y, rows, cols, ld, mode, rnb, s);
    }
    const bool regs_ok = regs_base && rows <= 32768;
    if (regs_ok) {
        const size_t n4 = rows >> 2;
        if (n4 <= 512) {  // one warp per column
            const size_t blocks = ceil_div(cols, size_t(8));
            const unsigned grid = unsigned(blocks < cap ? blocks : cap);
#define JZ_SM_WARP(NV) JZ_LAUNCH((softmax_reg_kernel<NV, false>), grid, 256, 0, s, out, a, y, rows, cols, ld, mode, rnb)
            if (n4 <= 32) JZ_SM_WARP(1);
            else if (n4 <= 64) JZ_SM_WARP(2);
            else if (n4 <= 128) JZ_SM_WARP(4);
            else if (n4 <= 256) JZ_SM_WARP(8);
            else JZ_SM_WARP(16);
#undef JZ_SM_WARP
        } else {          // one 512-thread CTA per column
            const unsigned grid = unsigned(cols < cap * 4 ? cols : cap * 4);
#define JZ_SM_BLOCK(NV) JZ_LAUNCH((softmax_reg_kernel<NV, true>), grid, 512, 0, s, out, a, y, rows, cols, ld, mode, rnb)
            static const bool no_pref = std::getenv("JZ_SOFTMAX_NO_PREFETCH") != nullptr;
            if (n4 <= 1024) JZ_SM_BLOCK(2);
            else if (n4 <= 2048) JZ_SM_BLOCK(4);
            else if (n4 <= 4096) JZ_SM_BLOCK(8);       // two CTAs per SM already overlap each other's phases (1.03 of the copy peak)
            else if (no_pref || cols < 2 * size_t(ctx().sm_count)) JZ_SM_BLOCK(16);
            else {
                // 16384 < rows <= 32768, one CTA per SM: persistent CTAs that prefetch their next column through shared
                // memory (0.77 -> 0.84 of the copy peak at 24576 rows, 0.85 -> 0.92 at 32768, profiles/r02b_long_column_softmax.log)
                static bool attr_done = false;
                if (!attr_done) {
                    JZ_CUDA(cudaFuncSetAttribute(softmax_reg_kernel<16, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 512 * 16));
                    attr_done = true;
                }
                const size_t resident = size_t(ctx().sm_count);
                const unsigned pgrid = unsigned(cols < resident ? cols : resident);
                JZ_LAUNCH((softmax_reg_kernel<16, true, true>), pgrid, 512, 16 * 512 * 16, s, out, a, y, rows, cols, ld, mode, rnb);
            }
#undef JZ_SM_BLOCK
        }
        return JZ_OK;
    }
    const size_t blocks = ceil_div(cols, size_t(8));
    const unsigned grid = unsigned(blocks < cap ? blocks : cap);
    if (vec) JZ_LAUNCH((softmax_warp_kernel<true>), grid, 256, 0, s, out, a, rows, cols, ld);
    else JZ_LAUNCH((softmax_warp_kernel<false>), grid, 256, 0, s, out, a, rows, cols, ld);
    if (mode == 1) {
        const size_t n = rows * cols;
        const size_t b2 = ceil_div(n, size_t(1024));
        JZ_LAUNCH((map2_inplace_kernel<CeGradF>), unsigned(b2 < cap ? b2 : cap), 256, 0, s, out, y, n, CeGradF{rnb});
    }
    return JZ_OK;
}

int jz_softmax_cols(float* out, const float* a, size_t rows, size_t cols, size_t ld, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    return softmax_impl(out, a, nullptr, rows, cols, ld, 0, 0.0f, as_stream(stream));
}

int jz_softmax_ce_grad(float* out, const float* x, const float* y, size_t rows, size_t cols, float nb,
                       jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    const float rnb = float(1.0 / double(nb));  // operator/ : scale((float)(1.0/r)), cpp/operators.hpp:236-246
    return softmax_impl(out, x, y, rows, cols, rows, 1, rnb, as_stream(stream));
}

int jz_softmax_ce_grad_scaled(float* out, const float* x, const float* y, size_t rows, size_t cols, float rnb,
                              jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    return softmax_impl(out, x, y, rows, cols, rows, 1, rnb, as_stream(stream));
}

int jz_nrm2(const float* x, size_t n, float* result_host, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    if (!result_host) return fail(JZ_ERR_ARG, "jz_nrm2: null result pointer");
    if (n == 0) { *result_host = 0.0f; return JZ_OK; }
    if (!x) return fail(JZ_ERR_ARG, "jz_nrm2: null pointer");
    cudaStream_t s = as_stream(stream);
    const size_t cap = size_t(ctx().sm_count) * 4;
    const bool vec = aligned16(x);
    const size_t want = ceil_div(vec ? (n >> 2) + 1 : n, size_t(256) * 4);
    const unsigned grid = unsigned(want < cap ? (want ? want : 1) : cap);
    WsGuard wg(s);
    int rc = ws_alloc(&wg.p, grid * sizeof(double) + 16, s);
    if (rc != JZ_OK) return rc;
    double* partial = static_cast<double*>(wg.p);
    float* dres = reinterpret_cast<float*>(partial + grid);
    JZ_LAUNCH(sumsq_partial_kernel, grid, 256, 0, s, partial, x, n, vec);
    JZ_LAUNCH(nrm2_final_kernel, 1, 256, 0, s, dres, partial, int(grid));
    JZ_CUDA(cudaMemcpyAsync(result_host, dres, sizeof(float), cudaMemcpyDeviceToHost, s));
    JZ_CUDA(cudaStreamSynchronize(s));
    return JZ_OK;
}

}  // extern "C"
