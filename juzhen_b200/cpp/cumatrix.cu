// cumatrix.cu -- member and free functions of Matrix<CUDAfloat> on top of the C ABI (include/jz_b200.h).
//
// Replaces the reference's cpp/cumatrix.cu + cpp/cukernels.cu.  Each function states the reference
// behaviour it reproduces (file:line) and makes exactly one jz_* call for the device work; shape
// checks happen here, before anything is launched.  Differences from the reference that are
// deliberate (and invisible through the API):
//   * internal temporaries are not zero-filled, and GEMM writes C with beta = 0 instead of
//     zero-fill + beta = 1 (cpp/cumatrix.cu:185-193);
//   * s1*M + a, sums, transposed adds and Hadamard products are single fused passes instead of
//     fill + axpy / fill x3 + gemv / geam + product;
//   * arithmetic follows the CPU oracle's rounding (separate multiply and add, cpp/core.hpp:413-482),
//     so elementwise results are bit-identical to Matrix<float>, not merely close.
#include "cumatrix.cuh"

#include <algorithm>

using std::string;

GPU_handle Matrix<CUDAfloat>::global_handle = nullptr;
unsigned long long GPUSampler::seed = 0;
unsigned long long GPUSampler::offset = 0;

namespace {
inline jz_stream_t S() { return jz_cpp_stream(); }
inline void require_same_shape(const Matrix<CUDAfloat>& a, const Matrix<CUDAfloat>& b) {
    if (a.num_row() != b.num_row() || a.num_col() != b.num_col())
        throw std::invalid_argument("Matrix dimensions are not compatible");
}
}  // namespace

// ------------------------------------------------------------------ storage and lifetime
std::shared_ptr<CUDAfloat[]> Matrix<CUDAfloat>::new_storage(size_t count) {
    return std::shared_ptr<CUDAfloat[]>(Memory<CUDAfloat>::allocate(count),
                                        [](CUDAfloat* p) { Memory<CUDAfloat>::free(p); });
}

Matrix<CUDAfloat>::Matrix(Raw, const char* name, size_t numrow, size_t numcol, bool trans)
    : numcol(numcol), numrow(numrow), transpose(trans), name(name), elements(new_storage(numrow * numcol)) {}

// public construction is observably zero-filled (cpp/cumatrix.cu:50-63; demo.cu relies on it)
Matrix<CUDAfloat>::Matrix(const char* name, size_t numrow, size_t numcol, int trans)
    : Matrix(Raw{}, name, numrow, numcol, trans != 0) {
    zeros();
}

// upload (cpp/cumatrix.cu:28-48): synchronous, physical buffer and flag carried over unchanged
Matrix<CUDAfloat>::Matrix(const Matrix<float>& M)
    : Matrix(Raw{}, ("cu_" + M.name).c_str(), M.numrow, M.numcol, M.transpose) {
    JZ_DO(jz_memcpy_h2d(dev(), M.elements.get(), count(), S()));
    JZ_DO(jz_sync(S()));
}

Matrix<CUDAfloat>::Matrix(const Matrix<CUDAfloat>& M)
    : Matrix(Raw{}, ("copy of" + M.name).c_str(), M.numrow, M.numcol, M.transpose) {
    JZ_DO(jz_memcpy_d2d(dev(), M.dev(), count(), S()));
}

Matrix<CUDAfloat>::Matrix(Matrix<CUDAfloat>&& M) noexcept
    : numcol(M.numcol), numrow(M.numrow), transpose(M.transpose), name(std::move(M.name)),
      elements(std::move(M.elements)) {
    M.elements = nullptr;
}

Matrix<CUDAfloat>& Matrix<CUDAfloat>::operator=(const Matrix<CUDAfloat>& M) {
    if (this == &M) return *this;
    name = "copy of " + M.name;
    if (count() != M.count() || !elements || elements.use_count() > 1) elements = new_storage(M.count());
    numrow = M.numrow;
    numcol = M.numcol;
    transpose = M.transpose;
    JZ_DO(jz_memcpy_d2d(dev(), M.dev(), count(), S()));
    return *this;
}

Matrix<CUDAfloat>& Matrix<CUDAfloat>::operator=(Matrix<CUDAfloat>&& M) noexcept {
    if (this == &M) return *this;
    name = std::move(M.name);
    numrow = M.numrow;
    numcol = M.numcol;
    transpose = M.transpose;
    elements = std::move(M.elements);
    M.elements = nullptr;
    return *this;
}

// download (cpp/cumatrix.cu:146-165): the sync point of the API
Matrix<float> Matrix<CUDAfloat>::to_host() const {
    Matrix<float> host((name + "->host").c_str(), numrow, numcol, transpose);
    JZ_DO(jz_memcpy_d2h(host.elements.get(), dev(), count(), S()));
    return host;
}

const Matrix<CUDAfloat> Matrix<CUDAfloat>::T() const {
    return Matrix<CUDAfloat>((name + "_T").c_str(), numrow, numcol, !transpose, elements);
}

// ------------------------------------------------------------------ fillers and RNG
void Matrix<CUDAfloat>::ones() { JZ_DO(jz_fill(dev(), count(), 1.0f, S())); }
void Matrix<CUDAfloat>::zeros() { JZ_DO(jz_fill(dev(), count(), 0.0f, S())); }

Matrix<CUDAfloat>& fill(Matrix<CUDAfloat>& M, double a) {
    JZ_DO(jz_fill(M.dev(), M.count(), float(a), S()));
    return M;
}

Matrix<CUDAfloat> Matrix<CUDAfloat>::ones(size_t m, size_t n) {
    Matrix<CUDAfloat> M(Raw{}, "ones", m, n, false);
    M.ones();
    return M;
}

Matrix<CUDAfloat> Matrix<CUDAfloat>::zeros(size_t m, size_t n) { return Matrix<CUDAfloat>("zeros", m, n); }

// cpp/cumatrix.cu:358-420 drew from cuRAND XORWOW (with a scratch cudaMalloc for odd counts); here a
// counter-based Philox kernel writes any count in place.  Not bit-compatible with cuRAND -- the
// reference's GPU stream was never reproducible against its CPU mt19937 stream either.
Matrix<CUDAfloat> Matrix<CUDAfloat>::randn(size_t m, size_t n) {
    Matrix<CUDAfloat> M(Raw{}, "randn", m, n, false);
    JZ_DO(jz_rand_normal(M.dev(), m * n, GPUSampler::seed, GPUSampler::offset, S()));
    GPUSampler::offset += m * n;
    return M;
}

Matrix<CUDAfloat> Matrix<CUDAfloat>::rand(size_t m, size_t n) {
    Matrix<CUDAfloat> M(Raw{}, "rand", m, n, false);
    JZ_DO(jz_rand_uniform(M.dev(), m * n, GPUSampler::seed, GPUSampler::offset, S()));
    GPUSampler::offset += m * n;
    return M;
}

// ------------------------------------------------------------------ GEMM (cpp/cumatrix.cu:177-197)
Matrix<CUDAfloat> Matrix<CUDAfloat>::dot(const Matrix<CUDAfloat>& B) const {
    if (num_col() != B.num_row()) throw std::invalid_argument("Matrix dimensions are not compatible");
    const size_t m = num_row(), n = B.num_col(), k = num_col();
    Matrix<CUDAfloat> C(Raw{}, "dot", m, n, false);
    JZ_DO(jz_gemm(transpose, B.transpose, m, n, k, 1.0f, dev(), numrow, B.dev(), B.numrow, 0.0f, C.dev(),
                  m ? m : 1, -1, S()));
    return C;
}

// ------------------------------------------------------------------ affine / axpby / reciprocal
// s1*M + a (cpp/cumatrix.cu:199-215); layout and flag preserved
Matrix<CUDAfloat> Matrix<CUDAfloat>::add(float a, float s1) const {
    Matrix<CUDAfloat> C(Raw{}, "add", numrow, numcol, transpose);
    JZ_DO(jz_affine(C.dev(), dev(), count(), s1, a, S()));
    return C;
}
void Matrix<CUDAfloat>::add(float a, float s1) { JZ_DO(jz_affine(dev(), dev(), count(), s1, a, S())); }

// M *= s1.  The CPU oracle's scale is add(0, s1) (cpp/core.hpp:148-149): s1*x + 0.0f.
void Matrix<CUDAfloat>::scale(float s1) { JZ_DO(jz_affine(dev(), dev(), count(), s1, 0.0f, S())); }

// s1*this + s2*B into a fresh, non-transposed matrix of the logical shape (cpp/cumatrix.cu:227-244)
Matrix<CUDAfloat> Matrix<CUDAfloat>::add(const Matrix<CUDAfloat>& B, float s1, float s2) const {
    require_same_shape(*this, B);
    const size_t r = num_row(), c = num_col();
    Matrix<CUDAfloat> C(Raw{}, "add", r, c, false);
    if (!transpose && !B.transpose) JZ_DO(jz_axpby(C.dev(), dev(), B.dev(), count(), s1, s2, S()));
    else JZ_DO(jz_axpby2d(C.dev(), r ? r : 1, r, c, dev(), numrow, transpose, B.dev(), B.numrow, B.transpose, s1, s2, S()));
    return C;
}

// in place: the result keeps this layout, B is read transposed iff the flags differ (cpp/cumatrix.cu:246-260)
void Matrix<CUDAfloat>::add(const Matrix<CUDAfloat>& B, float s1, float s2) {
    require_same_shape(*this, B);
    if (transpose == B.transpose) JZ_DO(jz_axpby(dev(), dev(), B.dev(), count(), s1, s2, S()));
    else JZ_DO(jz_axpby2d(dev(), numrow ? numrow : 1, numrow, numcol, dev(), numrow, 0, B.dev(), B.numrow, 1, s1, s2, S()));
}

// l / M (cpp/cumatrix.cu:263-303)
Matrix<CUDAfloat> Matrix<CUDAfloat>::eleminv(double l) const {
    Matrix<CUDAfloat> R(Raw{}, "elem_rec", numrow, numcol, transpose);
    JZ_DO(jz_eleminv(R.dev(), dev(), count(), float(l), S()));
    return R;
}
void Matrix<CUDAfloat>::eleminv(double l) { JZ_DO(jz_eleminv(dev(), dev(), count(), float(l), S())); }

float Matrix<CUDAfloat>::norm() const {  // cpp/cumatrix.cu:168-175 (cublasSnrm2): syncs
    float r = 0.0f;
    JZ_DO(jz_nrm2(dev(), count(), &r, S()));
    return r;
}

// ------------------------------------------------------------------ windows (cpp/cumatrix.cuh:188-221)
// The window is given in logical coordinates; on a flagged matrix rows and columns swap roles and the
// result carries the flag.  The copy is a strided 2-D move with 64-bit indexing.
Matrix<CUDAfloat> Matrix<CUDAfloat>::slice(size_t rstart, size_t rend, size_t cstart, size_t cend) const {
    if (transpose) { std::swap(rstart, cstart); std::swap(rend, cend); }
    const size_t r = rend - rstart, c = cend - cstart;
    Matrix<CUDAfloat> W(Raw{}, "submatrix", r, c, transpose);
    JZ_DO(jz_copy2d(W.dev(), r ? r : 1, dev() + cstart * numrow + rstart, numrow, r, c, 0, S()));
    return W;
}

// assignment into a window: M's buffer is taken in this matrix's physical orientation, as the
// reference's copyKernel does (cpp/cukernels.cu:72-90)
void Matrix<CUDAfloat>::slice(size_t rstart, size_t rend, size_t cstart, size_t cend, const Matrix<CUDAfloat>& M) {
    if (transpose) { std::swap(rstart, cstart); std::swap(rend, cend); }
    const size_t r = rend - rstart, c = cend - cstart;
    JZ_DO(jz_copy2d(dev() + cstart * numrow + rstart, numrow, M.dev(), r ? r : 1, r, c, 0, S()));
}

void copy(Matrix<CUDAfloat>& dest, const Matrix<CUDAfloat>& src) {  // cpp/cukernels.cu:241-256: dest is re-allocated
    dest.numrow = src.numrow;
    dest.numcol = src.numcol;
    dest.transpose = src.transpose;
    dest.elements.reset();
    dest.elements = Matrix<CUDAfloat>::new_storage(src.count());
    JZ_DO(jz_copy(dest.dev(), src.dev(), src.count(), S()));
}

// ------------------------------------------------------------------ reductions
// sum(M, 0): 1 x ncols, handed back as a flagged ncols x 1 buffer; sum(M, 1): nrows x 1 (cpp/cumatrix.cu:312-337)
Matrix<CUDAfloat> sum(const Matrix<CUDAfloat>& M, int dim) {
    const bool down_physical_columns = (dim == 0) != M.transpose;
    const size_t len = down_physical_columns ? M.numcol : M.numrow;
    Matrix<CUDAfloat> R(Matrix<CUDAfloat>::Raw{}, "sumM", len, 1, dim == 0);
    JZ_DO(jz_sum(R.dev(), M.dev(), M.numrow, M.numcol, M.numrow ? M.numrow : 1, down_physical_columns ? 0 : 1, S()));
    return R;
}

// ------------------------------------------------------------------ unary maps (cpp/cukernels.cu:156-239)
Matrix<CUDAfloat> jz_unary_new(int op, const char* name, const Matrix<CUDAfloat>& M) {
    Matrix<CUDAfloat> R(Matrix<CUDAfloat>::Raw{}, name, M.numrow, M.numcol, M.transpose);
    JZ_DO(jz_unary(op, R.dev(), M.dev(), M.count(), S()));
    return R;
}
Matrix<CUDAfloat> jz_unary_reuse(int op, Matrix<CUDAfloat>&& M) {
    JZ_DO(jz_unary(op, M.dev(), M.dev(), M.count(), S()));
    return std::move(M);
}

Matrix<CUDAfloat> exp(const Matrix<CUDAfloat>& M) { return jz_unary_new(JZ_EXP, "expM", M); }
Matrix<CUDAfloat> exp(Matrix<CUDAfloat>&& M) { return jz_unary_reuse(JZ_EXP, std::move(M)); }
Matrix<CUDAfloat> log(const Matrix<CUDAfloat>& M) { return jz_unary_new(JZ_LOG, "logM", M); }
Matrix<CUDAfloat> tanh(const Matrix<CUDAfloat>& M) { return jz_unary_new(JZ_TANH, "tanhM", M); }
Matrix<CUDAfloat> tanh(Matrix<CUDAfloat>&& M) { return jz_unary_reuse(JZ_TANH, std::move(M)); }
Matrix<CUDAfloat> d_tanh(const Matrix<CUDAfloat>& M) { return jz_unary_new(JZ_DTANH, "d_tanhM", M); }
Matrix<CUDAfloat> d_tanh(Matrix<CUDAfloat>&& M) { return jz_unary_reuse(JZ_DTANH, std::move(M)); }
Matrix<CUDAfloat> square(const Matrix<CUDAfloat>& M) { return jz_unary_new(JZ_SQUARE, "square", M); }
Matrix<CUDAfloat> square(Matrix<CUDAfloat>&& M) { return jz_unary_reuse(JZ_SQUARE, std::move(M)); }

// ------------------------------------------------------------------ Hadamard product (cpp/cukernels.cu:326-400)
// The result takes M1's layout and flag; with differing flags M2 is read through a tiled transpose in
// the same pass.  The rvalue overloads work in place only when the flags agree.
Matrix<CUDAfloat> hadmd(const Matrix<CUDAfloat>& M1, const Matrix<CUDAfloat>& M2) {
    require_same_shape(M1, M2);
    Matrix<CUDAfloat> R(Matrix<CUDAfloat>::Raw{}, "hadmd", M1.numrow, M1.numcol, M1.transpose);
    if (M1.transpose == M2.transpose) JZ_DO(jz_hadamard(R.dev(), M1.dev(), M2.dev(), M1.count(), S()));
    else JZ_DO(jz_hadamard2d(R.dev(), M1.numrow ? M1.numrow : 1, M1.numrow, M1.numcol, M1.dev(), M1.numrow, 0, M2.dev(),
                             M2.numrow, 1, S()));
    return R;
}
Matrix<CUDAfloat> hadmd(const Matrix<CUDAfloat>& M1, Matrix<CUDAfloat>&& M2) {
    require_same_shape(M1, M2);
    if (M1.transpose != M2.transpose) return hadmd(M1, static_cast<const Matrix<CUDAfloat>&>(M2));
    JZ_DO(jz_hadamard(M2.dev(), M2.dev(), M1.dev(), M1.count(), S()));
    return std::move(M2);
}
Matrix<CUDAfloat> hadmd(Matrix<CUDAfloat>&& M1, const Matrix<CUDAfloat>& M2) {
    require_same_shape(M1, M2);
    if (M1.transpose != M2.transpose) return hadmd(static_cast<const Matrix<CUDAfloat>&>(M1), M2);
    JZ_DO(jz_hadamard(M1.dev(), M1.dev(), M2.dev(), M1.count(), S()));
    return std::move(M1);
}
Matrix<CUDAfloat> hadmd(Matrix<CUDAfloat>&& M1, Matrix<CUDAfloat>&& M2) {
    return hadmd(static_cast<const Matrix<CUDAfloat>&>(M1), std::move(M2));
}

// ------------------------------------------------------------------ stacking (cpp/cukernels.cu:258-323)
// Empty inputs are dropped; an empty list or mismatching extents throw std::invalid_argument.
// Each input goes to its place in ONE strided (and, for flagged views, transposing) copy; vstack
// places rows directly instead of building hstack of the flipped views and transposing the result.
static void drop_empty(std::vector<MatrixView<CUDAfloat>>& v) {
    v.erase(std::remove_if(v.begin(), v.end(),
                           [](const MatrixView<CUDAfloat>& m) { return m.num_row() == 0 || m.num_col() == 0; }),
            v.end());
}

Matrix<CUDAfloat> hstack(std::vector<MatrixView<CUDAfloat>> matrices) {
    drop_empty(matrices);
    if (matrices.empty()) throw std::invalid_argument("hstack: input list is empty or contains only empty matrices");
    const size_t rows = matrices[0].num_row();
    size_t cols = 0;
    for (const auto& m : matrices) {
        if (m.num_row() != rows) throw std::invalid_argument("hstack: all matrices must have the same row count");
        cols += m.num_col();
    }
    Matrix<CUDAfloat> R(Matrix<CUDAfloat>::Raw{}, "hstack", rows, cols, false);
    size_t at = 0;
    for (const auto& m : matrices) {
        const float* src = reinterpret_cast<const float*>(m.data());
        // a flagged view stores its transpose: physical leading dimension = logical column count
        JZ_DO(jz_copy2d(R.dev() + at * rows, rows, src, m.get_transpose() ? m.num_col() : rows, rows, m.num_col(),
                        m.get_transpose() ? 1 : 0, S()));
        at += m.num_col();
    }
    return R;
}

const Matrix<CUDAfloat> vstack(std::vector<MatrixView<CUDAfloat>> matrices) {
    drop_empty(matrices);
    if (matrices.empty()) throw std::invalid_argument("hstack: input list is empty or contains only empty matrices");
    const size_t cols = matrices[0].num_col();
    size_t rows = 0;
    for (const auto& m : matrices) {
        if (m.num_col() != cols) throw std::invalid_argument("hstack: all matrices must have the same row count");
        rows += m.num_row();
    }
    Matrix<CUDAfloat> R(Matrix<CUDAfloat>::Raw{}, "vstack", rows, cols, false);  // materialised, not flagged
    size_t at = 0;
    for (const auto& m : matrices) {
        const float* src = reinterpret_cast<const float*>(m.data());
        JZ_DO(jz_copy2d(R.dev() + at, rows, src, m.get_transpose() ? cols : m.num_row(), m.num_row(), cols,
                        m.get_transpose() ? 1 : 0, S()));
        at += m.num_row();
    }
    return R;
}

// ------------------------------------------------------------------ printing and file IO (via the host)
std::ostream& operator<<(std::ostream& os, const Matrix<CUDAfloat>& M) {
    const Matrix<float> h = M.to_host();
    os << h.get_name() << " " << h.num_row() << " by " << h.num_col();
    for (size_t i = 0; i < h.num_row(); i++) {
        os << std::endl;
        for (size_t j = 0; j < h.num_col(); j++) os << h.elem(i, j) << " ";
    }
    return os;
}

template <>
void write(FILE* fp, const Matrix<CUDAfloat>& M) {
    write(fp, M.to_host());
}

template <>
void read(FILE* fp, Matrix<CUDAfloat>& M) {  // header (rows, cols, flag) + physical buffer, cpp/core.hpp:484-496
    Matrix<float> tmp("tmp", M.num_row(), M.num_col());
    read(fp, tmp);
    const size_t n = tmp.num_row() * tmp.num_col();
    if (n != M.count() || !M.elements) M.elements = Matrix<CUDAfloat>::new_storage(n);
    M.numrow = tmp.get_transpose() ? tmp.num_col() : tmp.num_row();
    M.numcol = tmp.get_transpose() ? tmp.num_row() : tmp.num_col();
    M.transpose = tmp.get_transpose() != 0;
    JZ_DO(jz_memcpy_h2d(M.dev(), reinterpret_cast<const float*>(tmp.data()), n, S()));
    JZ_DO(jz_sync(S()));
}
