// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin extern "C" shim over the UNMODIFIED reference headers (compiled from
// where they lie under /root/reference; nothing of the reference is copied
// here).  Every entry point wraps caller-owned float buffers as
// Matrix<float> views and runs the reference's own CPU code path
// (cpp/core.hpp, cpp/matrix.hpp, cpp/operators.hpp, cpp/juzhen.hpp) on them,
// so the outputs ARE the reference's outputs on this machine.
//
// Used for: (1) pinning oracle/jz_oracle.c, (2) generating tests/golden/*,
// (3) the "reference" CPU baseline in bench.py.  Built by oracle/Makefile into
// oracle/_ref/libjzref.so (git-ignored).
//
// All matrices cross this boundary in the reference's physical format:
// column-major `numrow x numcol` buffer + a `transpose` flag
// (cpp/core.hpp:96-98).  Results are written out in LOGICAL column-major
// order (out[j*R+i] = result.elem(i,j)) so the flag a result happens to carry
// does not matter to the caller.
#include "cpp/juzhen.hpp"

#include <cstdint>
#include <cstring>

int compute() { return 0; }  // cpp/juzhen.hpp:8 declares it; never called here.

namespace {

using MF = Matrix<float>;

// Non-owning view: public ctor cpp/core.hpp:113-115, T() flips the flag
// (cpp/core.hpp:346-351).
MF view(const float* p, size_t numrow, size_t numcol, int trans) {
    std::shared_ptr<float[]> sp(const_cast<float*>(p), [](float*) {});
    MF m("view", numrow, numcol, sp);
    if (trans) return m.T();
    return m;
}

void emit(const MF& r, float* out) {
    const size_t R = r.num_row(), C = r.num_col();
    for (size_t j = 0; j < C; j++)
        for (size_t i = 0; i < R; i++) out[j * R + i] = r.elem(i, j);
}

}  // namespace

extern "C" {

// op ids shared with include/jz_b200.h (jz_unary_op)
enum { OP_EXP = 0, OP_LOG = 1, OP_TANH = 2, OP_DTANH = 3, OP_SQUARE = 4, OP_SQRT = 5,
       OP_RELU = 6, OP_DRELU = 7 };

int ref_version() { return 1; }

// unary maps on a flat buffer -- cpp/matrix.hpp:252-424, cpp/juzhen.hpp:73-76,
// ml/util.cuh:72-80 (relu/d_relu lambdas restated here verbatim in meaning).
int ref_unary(int op, const float* in, float* out, size_t n) {
    const MF a = view(in, n, 1, 0);
    MF r;
    switch (op) {
        case OP_EXP: r = exp(a); break;
        case OP_LOG: r = log(a); break;
        case OP_TANH: r = tanh(a); break;
        case OP_DTANH: r = d_tanh(a); break;
        case OP_SQUARE: r = square(a); break;
        case OP_SQRT: r = sqrt(a); break;
        case OP_RELU: r = elemwise([=](float x) { return x > 0.0 ? x : 0.0; }, a); break;
        case OP_DRELU: r = elemwise([=](float x) { return x > 0.0 ? 1.0 : 0.0; }, a); break;
        default: return 1;
    }
    std::memcpy(out, r.data(), n * sizeof(float));
    return 0;
}

// s1*x + a -- Matrix::add(D b, D s1), cpp/core.hpp:452-469
int ref_affine(const float* in, float* out, size_t n, float s1, float a) {
    const MF x = view(in, n, 1, 0);  // const: select the allocating overload
    MF r = x.add(a, s1);
    std::memcpy(out, r.data(), n * sizeof(float));
    return 0;
}

// M / r via operator/(const Matrix&, double) -- cpp/operators.hpp:236-239
int ref_div_scalar(const float* in, float* out, size_t n, double r) {
    const MF a = view(in, n, 1, 0);
    MF q = a / r;
    std::memcpy(out, q.data(), n * sizeof(float));
    return 0;
}

// l / x -- Matrix::eleminv, cpp/core.hpp:471-482
int ref_eleminv(const float* in, float* out, size_t n, double l) {
    const MF x = view(in, n, 1, 0);
    MF r = x.eleminv(l);
    std::memcpy(out, r.data(), n * sizeof(float));
    return 0;
}

// s1*op(A) + s2*op(B) -- Matrix::add(const Matrix&, D, D) const, cpp/core.hpp:413-431
// A is (ar x ac, ta), B is (br x bc, tb) physical.  Returns 2 on the
// reference's std::invalid_argument.
int ref_axpby(const float* A, size_t ar, size_t ac, int ta, const float* B, size_t br, size_t bc,
              int tb, float s1, float s2, float* out) {
    try {
        const MF a = view(A, ar, ac, ta);
        MF r = a.add(view(B, br, bc, tb), s1, s2);
        emit(r, out);
    } catch (std::invalid_argument&) {
        return 2;
    }
    return 0;
}

// hadmd -- cpp/matrix.hpp:427-442
int ref_hadmd(const float* A, size_t ar, size_t ac, int ta, const float* B, size_t br, size_t bc,
              int tb, float* out) {
    try {
        const MF a = view(A, ar, ac, ta), b = view(B, br, bc, tb);  // const: never in place
        MF r = hadmd(a, b);
        emit(r, out);
    } catch (std::invalid_argument&) {
        return 2;
    }
    return 0;
}

// A / B elementwise -- cpp/operators.hpp:270-274 (eleminv(1) then hadmd)
int ref_div(const float* A, size_t ar, size_t ac, int ta, const float* B, size_t br, size_t bc,
            int tb, float* out) {
    try {
        const MF a = view(A, ar, ac, ta), b = view(B, br, bc, tb);
        MF r = a / b;
        emit(r, out);
    } catch (std::invalid_argument&) {
        return 2;
    }
    return 0;
}

// sum(M, dim) -- cpp/matrix.hpp:74-91 (cblas_sgemv / jz::gemv with a ones vector)
int ref_sum(const float* A, size_t ar, size_t ac, int ta, int dim, float* out) {
    MF r = sum(view(A, ar, ac, ta), dim);
    emit(r, out);
    return 0;
}

// reduce(max-functor, M, dim, 1) -- cpp/matrix.hpp:202-250 with the functor of
// ml/layer.hpp:254-259 (LogisticLayer column max, init -1e30f).
int ref_reduce_max(const float* A, size_t ar, size_t ac, int ta, int dim, float* out) {
    MF r = reduce(
        [](float* v, float* vdes, int lenv, int) {
            float m = -1e30f;
            for (int i = 0; i < lenv; i++) m = m > v[i] ? m : v[i];
            vdes[0] = m;
        },
        view(A, ar, ac, ta), dim, 1);
    emit(r, out);
    return 0;
}

// reduce with the serial-sum + max functor of
// tests/testElementwiseReduceTorchDump.cu:55-58 (k = 2 outputs per vector).
int ref_reduce_stats(const float* A, size_t ar, size_t ac, int ta, int dim, float* out) {
    MF r = reduce(
        [](float* src, float* dst, int n, int) {
            float s = 0.0f, mx = src[0];
            for (int i = 0; i < n; ++i) {
                s += src[i];
                mx = mx > src[i] ? mx : src[i];
            }
            dst[0] = s;
            dst[1] = mx;
        },
        view(A, ar, ac, ta), dim, 2);
    emit(r, out);
    return 0;
}

// A*B -- Matrix::dot, cpp/core.hpp:393-410 -> cblas_sgemm (cpp/helper.hpp:231-242)
int ref_gemm(const float* A, size_t ar, size_t ac, int ta, const float* B, size_t br, size_t bc,
             int tb, float* out) {
    try {
        MF r = view(A, ar, ac, ta) * view(B, br, bc, tb);
        std::memcpy(out, r.data(), r.num_row() * r.num_col() * sizeof(float));
    } catch (std::invalid_argument&) {
        return 2;
    }
    return 0;
}

// Output sub-block rows [r0, r1) x columns [c0, c1) of op(A)*op(B), through the reference's own rows() / columns() /
// dot (cpp/core.hpp:158-163, 393-410 -> cblas_sgemm): the OpenBLAS comparator BASELINE.md section 3 prescribes for
// GEMM parity at sizes where the full CPU product takes minutes (2048 x 2048 block of a 16384^3 product: 137 GFLOP).
int ref_gemm_subblock(const float* A, size_t ar, size_t ac, int ta, const float* B, size_t br, size_t bc, int tb,
                      size_t r0, size_t r1, size_t c0, size_t c1, float* out) {
    try {
        const MF a = view(A, ar, ac, ta), b = view(B, br, bc, tb);
        MF r = a.rows(r0, r1) * b.columns(c0, c1);
        emit(r, out);
    } catch (std::invalid_argument&) {
        return 2;
    }
    return 0;
}

// the README / config-1 chain: log(exp(X)+1.0f)/5.0f  (README example)
int ref_chain_softplus5(const float* in, float* out, size_t n) {
    const MF a = view(in, n, 1, 0);
    MF r = log(exp(a) + 1.0f) / 5.0f;
    std::memcpy(out, r.data(), n * sizeof(float));
    return 0;
}

// config 1 in full: log(exp(A*B/scale)+1)/5 on square n x n operands.
int ref_config1(const float* A, const float* B, size_t n, double scale, float* out) {
    const MF a = view(A, n, n, 0), b = view(B, n, n, 0);
    MF r = log(exp(a * b / scale) + 1.0f) / 5.0f;
    std::memcpy(out, r.data(), n * n * sizeof(float));
    return 0;
}

// materialised logical copy (what to_host()+elem() shows) -- covers T()
int ref_materialize(const float* A, size_t ar, size_t ac, int ta, float* out) {
    emit(view(A, ar, ac, ta), out);
    return 0;
}

// slice get -- cpp/core.hpp:353-363
int ref_slice(const float* A, size_t ar, size_t ac, int ta, size_t r0, size_t r1, size_t c0,
              size_t c1, float* out) {
    MF r = view(A, ar, ac, ta).slice(r0, r1, c0, c1);
    emit(r, out);
    return 0;
}

// slice set -- cpp/core.hpp:365-373; `dst` is modified in place (physical buffer)
int ref_slice_set(float* dst, size_t dr, size_t dc, int dt, size_t r0, size_t r1, size_t c0,
                  size_t c1, const float* S, size_t sr, size_t sc, int st) {
    MF d = view(dst, dr, dc, dt);
    d.slice(r0, r1, c0, c1, view(S, sr, sc, st));
    return 0;
}

// hstack / vstack of up to 8 inputs -- cpp/matrix.hpp:92-150
int ref_stack(int vertical, int count, const float** ptrs, const size_t* nr, const size_t* nc,
              const int* tr, float* out, size_t* out_rows, size_t* out_cols) {
    try {
        std::vector<MF> keep;
        keep.reserve(count);
        for (int i = 0; i < count; i++) keep.push_back(view(ptrs[i], nr[i], nc[i], tr[i]));
        std::vector<MatrixView<float>> views;
        for (auto& m : keep) views.emplace_back(m);
        MF r = vertical ? vstack<float>(views) : hstack<float>(views);
        *out_rows = r.num_row();
        *out_cols = r.num_col();
        emit(r, out);
    } catch (std::invalid_argument&) {
        return 2;
    }
    return 0;
}

// Frobenius norm -- cpp/core.hpp:337-344 (serial fp32 accumulation)
float ref_norm(const float* A, size_t n) { return view(A, n, 1, 0).norm(); }

// seeded inputs exactly as the reference draws them -- cpp/matrix.hpp:49-71
int ref_randn(unsigned seed, float* out, size_t n) {
    global_rand_gen.seed(seed);
    MF r = MF::randn(n, 1);
    std::memcpy(out, r.data(), n * sizeof(float));
    return 0;
}
int ref_rand(unsigned seed, float* out, size_t n) {
    global_rand_gen.seed(seed);
    MF r = MF::rand(n, 1);
    std::memcpy(out, r.data(), n * sizeof(float));
    return 0;
}

// softmax-CE head gradient exactly as LogisticLayer::grad composes it
// (ml/layer.hpp:252-264): mx = colmax, E = exp(X - 1*mx), Z = 1*sum(E,0),
// G = -(Y - E/Z)/nb.   X, Y are K x N plain.
int ref_softmax_ce_grad(const float* X, const float* Y, size_t K, size_t N, double nb,
                        float* out) {
    const MF input = view(X, K, N, 0), output = view(Y, K, N, 0);
    MF oneK1("oneK1", K, 1);
    oneK1.ones();
    auto mx = reduce(
        [](float* v, float* vdes, int lenv, int) {
            float m = -1e30f;
            for (int i = 0; i < lenv; i++) m = m > v[i] ? m : v[i];
            vdes[0] = m;
        },
        input, 0, 1);
    auto shifted = input - oneK1 * mx;
    auto E = exp(std::move(shifted));
    auto Z = oneK1 * sum(E, 0);
    MF g = -(output - E / std::move(Z)) / nb;
    emit(g, out);
    return 0;
}

// column softmax composite E / (1 * sum(E,0)) with max subtraction (same
// building blocks as above, ml/layer.hpp:254-262).
int ref_softmax_cols(const float* X, size_t K, size_t N, float* out) {
    const MF input = view(X, K, N, 0);
    MF oneK1("oneK1", K, 1);
    oneK1.ones();
    auto mx = reduce(
        [](float* v, float* vdes, int lenv, int) {
            float m = -1e30f;
            for (int i = 0; i < lenv; i++) m = m > v[i] ? m : v[i];
            vdes[0] = m;
        },
        input, 0, 1);
    auto shifted = input - oneK1 * mx;
    auto E = exp(std::move(shifted));
    auto Z = oneK1 * sum(E, 0);
    MF s = E / std::move(Z);
    emit(s, out);
    return 0;
}

// tests/testbasic.cu:6-12 expression on caller-provided 2x3 A,B (plain).
int ref_testbasic_expr(const float* A, const float* B, float* out) {
    const MF a = view(A, 2, 3, 0), b = view(B, 2, 3, 0);
    auto C = log(exp(-a / b) + exp(hadmd(b, a))) - (a.T() * b).rows(0, 2);
    emit(C, out);
    return 0;
}

#ifndef JUZHEN_NO_BLAS
extern void openblas_set_num_threads(int);
extern int openblas_get_num_threads(void);
extern char* openblas_get_config(void);
int ref_blas_threads(int n) {
    if (n > 0) openblas_set_num_threads(n);
    return openblas_get_num_threads();
}
const char* ref_blas_config() { return openblas_get_config(); }
#else
int ref_blas_threads(int) { return 1; }
const char* ref_blas_config() { return "JUZHEN_NO_BLAS (cpp/cpulinalg.hpp)"; }
#endif

}  // extern "C"
