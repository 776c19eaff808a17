// cumatrix.cuh -- Matrix<CUDAfloat> for B200 (sm_100a): the host-side C++ shell over libjz_b200.so.
//
// Drop-in for the reference's cpp/cumatrix.cuh: same class name, same member and free-function
// signatures (cpp/cumatrix.cuh:118-276,321-448), so cpp/operators.hpp, cpp/juzhen.hpp, ml/layer.hpp,
// ml/util.cuh and the examples/tests compile against it unchanged.  Nothing in here computes:
// every operation is one call into the C ABI of include/jz_b200.h (hand-written sm_100a kernels,
// no cuBLAS / cuRAND on the path).  What this file owns is the reference's *semantics*:
//   - column-major storage + lazy transpose flag; T() aliases the buffer (cpp/cumatrix.cu:305-310);
//   - lvalue overloads allocate, rvalue overloads reuse the operand's buffer and return the same
//     pointer (tests/testElementwiseReduceTorchDump.cu:45-48 checks the identity);
//   - the public (name, rows, cols) ctor is observably zero-filled (cpp/cumatrix.cu:50-63) while
//     internal temporaries are NOT (the reference pays 4 B/elem of zero-fill on every temporary);
//   - shape errors are std::invalid_argument("Matrix dimensions are not compatible") raised before
//     any launch (cpp/cumatrix.cu:181-184); device errors log and exit(1) (cpp/cumatrix.cuh:34-55);
//   - all work is issued on the legacy default stream so it stays ordered with the raw kernel
//     launches of unchanged callers;
//   - elementwise chains and GEMM epilogues are FUSED behind the eager-looking API by deferring work in the
//     storage object (jz_lazy.hpp): `log(exp(A*B)+1.0f)/5.0f` is one GEMM launch with a 4-step epilogue.
// Only the generic functor entry points elemwise<F> / reduce<F> keep kernels in this header: their
// functor is an nvcc extended-lambda type, so they cannot live behind a C ABI.
#ifndef JZ_B200_CUMATRIX_CUH
#define JZ_B200_CUMATRIX_CUH

#include <cublas_v2.h>  // type names only: GPU_handle / global_handle are part of the reference's surface
#include <cuda.h>
#include <cuda_runtime.h>

#include <jz_b200.h>

#include <stdexcept>
#include <type_traits>

#include "jz_lazy.hpp"

#include "core.hpp"
#include "matrix.hpp"
#include "operators.hpp"

// ---- error conventions of the reference (macro names are used by ml/layer.hpp, ml/util.cuh)
inline void jz_cuda_guard(cudaError_t code, const char* file, int line) {
    if (code != cudaSuccess) {
        std::fprintf(stderr, "CUDA ERROR: %s %s:%d\n", cudaGetErrorString(code), file, line);
        LOG_ERROR("CUDA ERROR: {} {}:{}", cudaGetErrorString(code), file, line);
        ERROR_OUT;
    }
}
#define CudaErrorCheck(ans) \
    { jz_cuda_guard((ans), __FILE__, __LINE__); }

inline void jz_cublas_guard(cublasStatus_t code, const char* func, const char* file, int line) {
    if (code != CUBLAS_STATUS_SUCCESS) {
        std::fprintf(stderr, "CUBLAS ERROR: status %d, %s, %s:%d\n", int(code), func, file, line);
        LOG_ERROR("CUBLAS ERROR: {}, {}, {}:{}", int(code), func, file, line);
        ERROR_OUT;
    }
}
#define CuBLASErrorCheck(ans) \
    { jz_cublas_guard((ans), __FUNCTION__, __FILE__, __LINE__); }

// status of a C-ABI call -> the reference's behaviour
inline void jz_guard(int rc, const char* what) {
    if (rc == JZ_OK) return;
    if (rc == JZ_ERR_SHAPE) throw std::invalid_argument("Matrix dimensions are not compatible");
    std::fprintf(stderr, "jz_b200 error %d in %s: %s\n", rc, what, jz_last_error());
    LOG_ERROR("jz_b200 error {} in {}: {}", rc, what, jz_last_error());
    ERROR_OUT;
}
#define JZ_DO(call) jz_guard((call), #call)

// launch geometry macros kept for callers that still launch their own one-thread-per-element kernels
#define threadsPerBlock 1024
#define cudaConfig(numElem) ((unsigned int)numElem + threadsPerBlock - 1) / threadsPerBlock, threadsPerBlock

typedef cublasHandle_t GPU_handle;

template <>
class Matrix<CUDAfloat> {
    // Field names and order of meaning are a contract: MatrixView<CUDAfloat> (cpp/core.hpp:60-64)
    // reads elements / numrow / numcol / transpose directly.
    size_t numcol;
    size_t numrow;
    bool transpose;
    std::string name;
    // shared handle on the device storage; .get() hands MatrixView real (materialised) bytes
    jzb200::LazyBuf<CUDAfloat> elements;

    struct Raw {};  // tag: internal temporary, storage left uninitialised
    Matrix(Raw, const char* name, size_t numrow, size_t numcol, bool trans);
    Matrix(const char* name, size_t numrow, size_t numcol, int trans);  // zero-filled
    Matrix(const char* name, size_t numrow, size_t numcol, int trans, jzb200::LazyBuf<CUDAfloat> storage)
        : numcol(numcol), numrow(numrow), transpose(trans != 0), name(name), elements(std::move(storage)) {}

    static jzb200::LazyBuf<CUDAfloat> new_storage(size_t count);
    jzb200::Storage& store() const { return *elements.storage(); }
    float* dev() const {  // bytes for READING: runs whatever was deferred on this storage
        store().materialize();
        return store().ptr;
    }
    float* wdev() {       // bytes for in-place MODIFICATION: deferred readers take their snapshot first
        store().before_write();
        return store().ptr;
    }
    size_t count() const { return numrow * numcol; }
    // in-place add against a not-yet-computed u*ones^T / ones*v^T product as one broadcast pass (cumatrix.cu)
    bool add_broadcast(const Matrix<CUDAfloat>& B, float s1, float s2);
    // deferred out-of-place elementwise result (same physical shape and flag as this matrix)
    Matrix<CUDAfloat> mapped(const char* name, const jz_step& step) const;

   public:
    // kept for source compatibility with code that talks to cuBLAS itself (TransformerLayer's batched
    // attention, ml/layer.hpp:2896-2926).  This backend never uses it; the launcher only creates it
    // when built with -DJZ_LEGACY_CUBLAS_HANDLE.
    static GPU_handle global_handle;

    explicit Matrix(const Matrix<float>& M);  // synchronous upload, keeps the transpose flag
    Matrix(const char* name, size_t numrow, size_t numcol) : Matrix(name, numrow, numcol, 0) {}
    Matrix() : Matrix("un_init", 2, 2, 0) {}

    Matrix(const Matrix<CUDAfloat>& M);
    Matrix(Matrix<CUDAfloat>&& M) noexcept;
    Matrix<CUDAfloat>& operator=(const Matrix<CUDAfloat>& M);
    Matrix<CUDAfloat>& operator=(Matrix<CUDAfloat>&& M) noexcept;

    inline size_t idx(size_t i, size_t j) const { return transpose ? i * numrow + j : j * numrow + i; }
    // element access dereferences DEVICE memory on the host, as in the reference: signature only
    inline CUDAfloat elem(size_t i, size_t j) const { return elements[idx(i, j)]; }
    inline CUDAfloat& elem(size_t i, size_t j) { return elements[idx(i, j)]; }
    inline CUDAfloat operator()(size_t i, size_t j) const { return elements[idx(i, j)]; }
    inline CUDAfloat& operator()(size_t i, size_t j) { return elements[idx(i, j)]; }

    inline size_t num_col() const { return transpose ? numrow : numcol; }
    inline size_t num_row() const { return transpose ? numcol : numrow; }
    inline size_t get_transpose() const { return transpose; }
    std::string get_name() const { return name; }
    // the raw device pointer leaves the class: materialise, and never defer on this storage again
    // (callers such as ml/layer.hpp launch their own kernels on it)
    const CUDAfloat* data() const { return elements ? reinterpret_cast<const CUDAfloat*>(store().escape()) : nullptr; }

    void ones();
    void zeros();
    static Matrix<CUDAfloat> randn(size_t m, size_t n);
    static Matrix<CUDAfloat> rand(size_t m, size_t n);
    static Matrix<CUDAfloat> ones(size_t m, size_t n);
    static Matrix<CUDAfloat> zeros(size_t m, size_t n);

    Matrix<CUDAfloat> dot(const Matrix<CUDAfloat>& B) const;

    Matrix<CUDAfloat> add(const Matrix<CUDAfloat>& B, float s1, float s2) const;  // s1*this + s2*B
    void add(const Matrix<CUDAfloat>& B, float s1, float s2);
    Matrix<CUDAfloat> add(float a, float s1) const;                               // s1*this + a
    void add(float a, float s1);
    Matrix<CUDAfloat> scale(float s1) const { return add(0, s1); }
    void scale(float s1);
    void eleminv(double l);                                                       // l / this
    Matrix<CUDAfloat> eleminv(double l) const;

    float norm() const;
    const Matrix<CUDAfloat> T() const;
    Matrix<float> to_host() const;

    Matrix<CUDAfloat> slice(size_t rstart, size_t rend, size_t cstart, size_t cend) const;
    void slice(size_t rstart, size_t rend, size_t cstart, size_t cend, const Matrix<CUDAfloat>& M);
    Matrix<CUDAfloat> rows(size_t rstart, size_t rend) const { return slice(rstart, rend, 0, num_col()); }
    void rows(size_t rstart, size_t rend, const Matrix<CUDAfloat>& M) { slice(rstart, rend, 0, num_col(), M); }
    Matrix<CUDAfloat> columns(size_t cstart, size_t cend) const { return slice(0, num_row(), cstart, cend); }
    void columns(size_t cstart, size_t cend, const Matrix<CUDAfloat>& M) { slice(0, num_row(), cstart, cend, M); }

    friend Matrix<CUDAfloat> sum(const Matrix<CUDAfloat>& M, int dim);
    friend Matrix<CUDAfloat> jz_unary_new(int op, const char* name, const Matrix<CUDAfloat>& M);
    friend Matrix<CUDAfloat> jz_unary_reuse(int op, Matrix<CUDAfloat>&& M);

    template <class Function>
    friend Matrix<CUDAfloat> reduce(Function func, const Matrix<CUDAfloat>& M, int dim, int k);
    template <class Function>
    friend Matrix<CUDAfloat> elemwise(Function func, const Matrix<CUDAfloat>& M);
    template <class Function>
    friend Matrix<CUDAfloat> elemwise(Function func, Matrix<CUDAfloat>&& M);

    friend void copy(Matrix<CUDAfloat>& dest, const Matrix<CUDAfloat>& src);
    friend Matrix<CUDAfloat>& fill(Matrix<CUDAfloat>& M, double a);
    friend class MatrixView<CUDAfloat>;
    friend Matrix<CUDAfloat> hstack(std::vector<MatrixView<CUDAfloat>> matrices);
    friend const Matrix<CUDAfloat> vstack(std::vector<MatrixView<CUDAfloat>> matrices);
    friend Matrix<CUDAfloat> hadmd(const Matrix<CUDAfloat>& M1, const Matrix<CUDAfloat>& M2);
    friend Matrix<CUDAfloat> hadmd(const Matrix<CUDAfloat>& M1, Matrix<CUDAfloat>&& M2);
    friend Matrix<CUDAfloat> hadmd(Matrix<CUDAfloat>&& M1, const Matrix<CUDAfloat>& M2);
    friend void read<CUDAfloat>(FILE* fp, Matrix<CUDAfloat>& M);
};

// Counter-based Philox4x32-10 stream behind Matrix<CUDAfloat>::randn / rand (replaces the cuRAND
// XORWOW generator, cpp/cumatrix.cuh:278-319).  Same usage: construct once with a seed in compute().
struct GPUSampler {
    static unsigned long long seed;
    static unsigned long long offset;  // elements drawn so far: successive calls never overlap
    explicit GPUSampler(int s) {
        seed = (unsigned long long)(unsigned int)s;
        offset = 0;
        LOG_INFO("GPU sampler is initialized with seed {}.", s);
    }
    void setseed(int s) { offset = (unsigned long long)(unsigned int)s; }  // the reference moves the stream offset
    ~GPUSampler() { LOG_INFO("GPU sampler is destroyed!"); }
};

// ---- free functions (cpp/cumatrix.cuh:321-332,432-448)
Matrix<CUDAfloat> sum(const Matrix<CUDAfloat>& M, int dim);
std::ostream& operator<<(std::ostream& os, const Matrix<CUDAfloat>& M);
Matrix<CUDAfloat> exp(const Matrix<CUDAfloat>& M);
Matrix<CUDAfloat> exp(Matrix<CUDAfloat>&& M);
Matrix<CUDAfloat> log(const Matrix<CUDAfloat>& M);
Matrix<CUDAfloat> tanh(const Matrix<CUDAfloat>& M);
Matrix<CUDAfloat> tanh(Matrix<CUDAfloat>&& M);
Matrix<CUDAfloat> d_tanh(const Matrix<CUDAfloat>& M);
Matrix<CUDAfloat> d_tanh(Matrix<CUDAfloat>&& M);
Matrix<CUDAfloat> square(const Matrix<CUDAfloat>& M);
Matrix<CUDAfloat> square(Matrix<CUDAfloat>&& M);
Matrix<CUDAfloat> hadmd(const Matrix<CUDAfloat>& M1, const Matrix<CUDAfloat>& M2);
Matrix<CUDAfloat> hadmd(const Matrix<CUDAfloat>& M1, Matrix<CUDAfloat>&& M2);
Matrix<CUDAfloat> hadmd(Matrix<CUDAfloat>&& M1, const Matrix<CUDAfloat>& M2);
Matrix<CUDAfloat> hadmd(Matrix<CUDAfloat>&& M1, Matrix<CUDAfloat>&& M2);
Matrix<CUDAfloat>& fill(Matrix<CUDAfloat>& M, double a);
void copy(Matrix<CUDAfloat>& dest, const Matrix<CUDAfloat>& src);
Matrix<CUDAfloat> hstack(std::vector<MatrixView<CUDAfloat>> matrices);
const Matrix<CUDAfloat> vstack(std::vector<MatrixView<CUDAfloat>> matrices);

template <>
void write(FILE* fp, const Matrix<CUDAfloat>& M);
template <>
void read(FILE* fp, Matrix<CUDAfloat>& M);

// ------------------------------------------------------------------ generic functor kernels
namespace jzb200 {

inline cudaStream_t stream() { return reinterpret_cast<cudaStream_t>(jz_cpp_stream()); }

inline unsigned stream_grid(size_t tiles) {
    const size_t cap = 0x7fffffffu;  // one 16 KB tile per CTA: measured faster than a persistent grid on B200
    return unsigned(tiles < cap ? (tiles ? tiles : 1) : cap);
}

// out[i] = f(in[i]); 128-bit accesses, 4 independent loads in flight per thread, grid-stride.
// `out` may alias `in` (the rvalue overload).
template <class Function>
__global__ void __launch_bounds__(256) functor_map_v4(Function f, float* out, const float* in, size_t n) {
    const size_t n4 = n >> 2;
    const float4* in4 = reinterpret_cast<const float4*>(in);
    float4* out4 = reinterpret_cast<float4*>(out);
    for (size_t base = size_t(blockIdx.x) * 1024; base < n4; base += size_t(gridDim.x) * 1024) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const size_t i = base + u * 256 + threadIdx.x;
            if (i < n4) v[u] = in4[i];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const size_t i = base + u * 256 + threadIdx.x;
            if (i < n4) out4[i] = make_float4(f(v[u].x), f(v[u].y), f(v[u].z), f(v[u].w));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const size_t i = (n4 << 2) + threadIdx.x;
        out[i] = f(in[i]);
    }
}

template <class Function>
__global__ void __launch_bounds__(256) functor_map_s(Function f, float* out, const float* in, size_t n) {
    for (size_t i = size_t(blockIdx.x) * 256 + threadIdx.x; i < n; i += size_t(gridDim.x) * 256) out[i] = f(in[i]);
}

template <class Function>
inline void launch_functor_map(Function f, float* out, const float* in, size_t n) {
    if (n == 0) return;
    const bool vec = ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(in)) & 15u) == 0;
    if (vec) functor_map_v4<<<stream_grid((n / 4 + 1023) / 1024), 256, 0, stream()>>>(f, out, in, n);
    else functor_map_s<<<stream_grid((n + 1023) / 1024), 256, 0, stream()>>>(f, out, in, n);
    CudaErrorCheck(cudaPeekAtLastError());
}

// The reduce functor is an opaque serial program over one vector:
//   func(float* v, float* vdes, int lenv, int lendes)   (cpp/cumatrix.cuh:334-342)
// It may overwrite v (examples/knn.cu:61-69) and capture device pointers, so it cannot be split or
// re-associated: one thread runs one vector.  128-thread CTAs keep more SMs busy than the
// reference's 1024 when there are few vectors.
template <class Function>
__global__ void __launch_bounds__(128) functor_reduce_kernel(Function func, float* vecdes, float* vec, size_t lenvec,
                                                             size_t lenvecdes, size_t numvecs) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < numvecs) func(&vec[i * lenvec], &vecdes[i * lenvecdes], lenvec, lenvecdes);
}

}  // namespace jzb200

namespace jzb200 {
// Is this reduce functor "the maximum of the vector, floored at -1e30" -- the one LogisticLayer::grad / eval use
// (ml/layer.hpp:254-259)?  The functor is opaque, but a __host__ __device__ extended lambda can be CALLED on the host:
// probe it with a few vectors (including the empty one for its initial value) and compare with that definition.  jz_max
// computes exactly this (same floor), so a functor that passes may be replaced by the deferred column-max stage of the
// softmax head; a functor that is device-only, or fails any probe, runs as the opaque kernel it always was.
template <class Function>
inline bool functor_is_floored_max(Function& func) {
#if defined(__CUDACC_EXTENDED_LAMBDA__)
    // captureless lambdas only: a functor with captures may hold DEVICE pointers (examples/knn.cu:84 does) and must
    // never be run on the host
    if constexpr (__nv_is_extended_host_device_lambda_closure_type(Function) && std::is_empty<Function>::value) {
        static const float probes[4][6] = {{-3.5f, 2.25f, 2.0f, -7.0f, 0.5f, 1.0f}, {-9.0f, -4.0f, -4.5f, -100.0f, -5.0f, -6.0f},
                                           {0.0f, -0.0f, 1e-30f, -1e-30f, 0.0f, 0.0f}, {7.0f, 7.0f, 7.0f, 7.0f, 7.0f, 8.0f}};
        static const float want[4] = {2.25f, -4.0f, 1e-30f, 8.0f};
        for (int t = 0; t < 4; t++) {
            float v[6], out[1] = {123.0f};
            for (int i = 0; i < 6; i++) v[i] = probes[t][i];
            func(v, out, 6, 1);
            if (out[0] != want[t]) return false;
            for (int i = 0; i < 6; i++)
                if (v[i] != probes[t][i]) return false;   // a functor that scribbles on its input is not a pure max
        }
        float none[1] = {0.0f}, out[1] = {123.0f};
        func(none, out, 0, 1);
        return out[0] == -1e30f;
    }
#endif
    (void)func;
    return false;
}
}  // namespace jzb200

template <class Function>
Matrix<CUDAfloat> reduce(Function func, const Matrix<CUDAfloat>& M, int dim, int k) {
    // the functor walks PHYSICAL columns; reducing along the other direction needs the transpose in memory
    const bool along_physical_columns = (dim == 0) != M.transpose;
    if (k == 1 && dim == 0 && !M.transpose && M.numrow > 0 && M.numcol > 0 && M.store().lazy_ok() && jzb200::functor_is_floored_max(func)) {
        // column maxima, DEFINED but not computed: first stage of the softmax head (jz_lazy.hpp, Producer::COLMAX)
        Matrix<CUDAfloat> result(Matrix<CUDAfloat>::Raw{}, "resM", 1, M.numcol, false);
        std::unique_ptr<jzb200::Producer> p(new jzb200::Producer());
        p->kind = jzb200::Producer::COLMAX;
        p->src = M.elements.storage();
        p->rows = M.numrow;
        p->cols = M.numcol;
        result.store().producer = std::move(p);
        jzb200::add_reader(M.elements.storage(), result.elements.storage());
        return result;
    }
    if (along_physical_columns) {
        Matrix<CUDAfloat> result(Matrix<CUDAfloat>::Raw{}, "resM", k, M.numcol, false);
        JZ_DO(jz_fill(result.dev(), result.count(), 0.0f, jz_cpp_stream()));  // functors may accumulate into vdes
        M.store().before_write();  // the functor may scribble on its input vector (examples/knn.cu:61-69)
        if (M.numcol)
            jzb200::functor_reduce_kernel<<<unsigned((M.numcol + 127) / 128), 128, 0, jzb200::stream()>>>(
                func, result.dev(), M.dev(), M.numrow, size_t(k), M.numcol);
        CudaErrorCheck(cudaPeekAtLastError());
        if (M.transpose) return result.T();
        return result;
    }
    // materialise the physical transpose once (bit-exact tiled copy; the reference zero-fills + geam-adds)
    Matrix<CUDAfloat> t(Matrix<CUDAfloat>::Raw{}, "tM", M.numcol, M.numrow, false);
    JZ_DO(jz_copy2d(t.dev(), t.numrow, M.dev(), M.numrow, t.numrow, t.numcol, 1, jz_cpp_stream()));
    Matrix<CUDAfloat> result(Matrix<CUDAfloat>::Raw{}, "resM", k, t.numcol, false);
    JZ_DO(jz_fill(result.dev(), result.count(), 0.0f, jz_cpp_stream()));
    if (t.numcol)
        jzb200::functor_reduce_kernel<<<unsigned((t.numcol + 127) / 128), 128, 0, jzb200::stream()>>>(
            func, result.dev(), t.dev(), t.numrow, size_t(k), t.numcol);
    CudaErrorCheck(cudaPeekAtLastError());
    if (M.transpose) return result;
    return result.T();
}

template <class Function>
Matrix<CUDAfloat> elemwise(Function func, const Matrix<CUDAfloat>& M) {
    Matrix<CUDAfloat> result(Matrix<CUDAfloat>::Raw{}, "resM", M.numrow, M.numcol, M.transpose);
    jzb200::launch_functor_map(func, result.dev(), M.dev(), M.count());
    return result;
}

template <class Function>
Matrix<CUDAfloat> elemwise(Function func, Matrix<CUDAfloat>&& M) {
    float* p = M.wdev();
    jzb200::launch_functor_map(func, p, p, M.count());
    return std::move(M);
}

#endif
