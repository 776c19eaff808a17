// shell_capi.cu -- TEST INFRASTRUCTURE: a flat C wrapper over the C++ shell (juzhen_b200/cpp/cumatrix.cuh: the
// Matrix<CUDAfloat> class, the reference's operators.hpp / juzhen.hpp on top of it), so that the Python parity tests can
// drive the SAME code a C++ user of the reference compiles against -- operator dispatch, lvalue / rvalue overloads, the
// lazy transpose flag, deferred fills / products / elementwise chains -- and pin it against the reference's golden
// vectors and against the eager C-ABI mirror (juzhen_b200/matrix.py) bit for bit (tests/test_shell_gpu.py).
// Built by juzhen_b200/cpp/build_dropin.py into build/dropin/lib/libjz_shell_capi.so.  Not part of the product.
//
// Every function returns a handle (a heap Matrix<CUDAfloat>*) or NULL on error; shell_last_error() has the text and
// shell_last_error_is_shape() says whether it was the reference's std::invalid_argument (tests/testbasic.cu:114-249).
#include <cstring>
#include <string>
#include <vector>

#include "../cpp/juzhen.hpp"
#include "../ml/layer.hpp"

int compute() { return 0; }   // cpp/juzhen.hpp declares it; this library has no main

namespace {
typedef Matrix<CUDAfloat> CM;
std::string g_err;
int g_shape = 0;

struct Pools {   // what main() holds in the launcher (cpp/launcher.cu:78): the pools' lifetime is the process's
    Memory<int> mi; Memory<float> mf; Memory<double> md; Memory<CUDAfloat> mc;
};
Pools* g_pools = nullptr;

template <class F>
CM* guarded(F f) {
    g_err.clear();
    g_shape = 0;
    try {
        return new CM(f());
    } catch (const std::invalid_argument& e) {
        g_err = e.what();
        g_shape = 1;
    } catch (const std::exception& e) {
        g_err = e.what();
    }
    return nullptr;
}
CM& MX(void* h) { return *static_cast<CM*>(h); }
}  // namespace

extern "C" {
#define SHELL_API __attribute__((visibility("default")))

SHELL_API int shell_init(int device) {
    if (jz_init(device) != JZ_OK) { g_err = jz_last_error(); return 1; }
    if (!g_pools) g_pools = new Pools();
    return 0;
}
SHELL_API const char* shell_last_error(void) { return g_err.c_str(); }
SHELL_API int shell_last_error_is_shape(void) { return g_shape; }

// column-major host data -> device matrix (the explicit Matrix(const Matrix<float>&) constructor)
SHELL_API void* shell_from_host(const float* data, size_t r, size_t c) {
    return guarded([&] {
        Matrix<float> h("h", r, c);
        if (r * c) std::memcpy(const_cast<float*>(h.data()), data, r * c * sizeof(float));
        return CM(h);
    });
}
SHELL_API void* shell_named(size_t r, size_t c) { return guarded([&] { return CM("named", r, c); }); }
SHELL_API void* shell_static(const char* what, size_t r, size_t c) {
    const std::string w(what);
    return guarded([&] { return w == "ones" ? CM::ones(r, c) : w == "zeros" ? CM::zeros(r, c) : w == "randn" ? CM::randn(r, c) : CM::rand(r, c); });
}
SHELL_API void shell_free(void* h) { delete static_cast<CM*>(h); }
SHELL_API void shell_info(void* h, size_t* rows, size_t* cols, int* trans, const void** dev) {
    *rows = MX(h).num_row(); *cols = MX(h).num_col(); *trans = int(MX(h).get_transpose()); *dev = MX(h).data();
}
// logical matrix, column-major, through to_host() (Matrix<float> keeps the flag: read it by element)
SHELL_API int shell_to_host(void* h, float* out) {
    try {
        Matrix<float> hm = MX(h).to_host();
        const size_t r = hm.num_row(), c = hm.num_col();
        for (size_t j = 0; j < c; j++)
            for (size_t i = 0; i < r; i++) out[j * r + i] = hm.elem(i, j);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
SHELL_API float shell_norm(void* h) { return MX(h).norm(); }

// unary / scalar / binary operators by name.  `move_a` / `move_b`: pass that operand as an rvalue (std::move), which
// selects the reference's && overloads (in place on the operand's buffer); the handle stays valid but empty.
SHELL_API void* shell_unary(const char* op, void* a, int move_a) {
    const std::string o(op);
    return guarded([&]() -> CM {
        CM& A = MX(a);
        if (o == "exp") return move_a ? exp(std::move(A)) : exp(A);
        if (o == "log") return log(A);                       // no && overload in the reference
        if (o == "tanh") return move_a ? tanh(std::move(A)) : tanh(A);
        if (o == "d_tanh") return move_a ? d_tanh(std::move(A)) : d_tanh(A);
        if (o == "square") return move_a ? square(std::move(A)) : square(A);
        if (o == "sqrt") return move_a ? sqrt(std::move(A)) : sqrt(A);          // cpp/juzhen.hpp:76-82 (elemwise<F>)
        if (o == "neg") return move_a ? -std::move(A) : -A;
        if (o == "T") return A.T();
        if (o == "copy") return CM(A);
        if (o == "sum0") return sum(A, 0);
        if (o == "sum1") return sum(A, 1);
        if (o == "relu") return elemwise([=] __device__(float x) { return x > 0.0f ? x : 0.0f; }, A);
        if (o == "d_relu") return elemwise([=] __device__(float x) { return x > 0.0f ? 1.0f : 0.0f; }, A);
        if (o == "colmax")    // the functor LogisticLayer uses (ml/layer.hpp:254-259): maximum of the vector, floored at -1e30
            return reduce([] __device__(float* v, float* vdes, int lenv, int lendes) {
                float mx = -1e30f;
                for (int i = 0; i < lenv; i++) mx = v[i] > mx ? v[i] : mx;
                vdes[0] = mx; }, A, 0, 1);
        throw std::runtime_error("shell_unary: unknown op " + o);
    });
}
SHELL_API void* shell_scalar(const char* op, void* a, double s, int move_a) {
    const std::string o(op);
    return guarded([&]() -> CM {
        CM& A = MX(a);
        if (o == "add") return move_a ? std::move(A) + s : A + s;
        if (o == "radd") return move_a ? s + std::move(A) : s + A;
        if (o == "sub") return move_a ? std::move(A) - s : A - s;
        if (o == "rsub") return move_a ? s - std::move(A) : s - A;
        if (o == "mul") return move_a ? std::move(A) * s : A * s;
        if (o == "rmul") return move_a ? s * std::move(A) : s * A;
        if (o == "div") return move_a ? std::move(A) / s : A / s;
        if (o == "rdiv") return move_a ? s / std::move(A) : s / A;
        throw std::runtime_error("shell_scalar: unknown op " + o);
    });
}
SHELL_API void* shell_binary(const char* op, void* a, void* b, int move_a, int move_b) {
    const std::string o(op);
    return guarded([&]() -> CM {
        CM& A = MX(a);
        CM& B = MX(b);
        if (o == "add") return move_a ? std::move(A) + B : move_b ? A + std::move(B) : A + B;
        if (o == "sub") return move_a ? std::move(A) - B : move_b ? A - std::move(B) : A - B;
        if (o == "mul") return A * B;
        if (o == "div") return move_a ? std::move(A) / B : move_b ? A / std::move(B) : A / B;
        if (o == "hadmd") return move_a ? hadmd(std::move(A), B) : move_b ? hadmd(A, std::move(B)) : hadmd(A, B);
        throw std::runtime_error("shell_binary: unknown op " + o);
    });
}
// compound assignment on the handle itself: += -= (matrix), += -= *= /= (scalar)
SHELL_API int shell_inplace(const char* op, void* a, void* b, double s) {
    const std::string o(op);
    g_err.clear(); g_shape = 0;
    try {
        CM& A = MX(a);
        if (o == "iadd") A += MX(b);
        else if (o == "isub") A -= MX(b);
        else if (o == "iadds") A += s;
        else if (o == "isubs") A -= s;
        else if (o == "zeros") A.zeros();
        else if (o == "ones") A.ones();
        else if (o == "fill") fill(A, s);
        else throw std::runtime_error("shell_inplace: unknown op " + o);
        return 0;
    } catch (const std::invalid_argument& e) { g_err = e.what(); g_shape = 1; }
    catch (const std::exception& e) { g_err = e.what(); }
    return 1;
}
SHELL_API void* shell_slice(void* a, size_t r0, size_t r1, size_t c0, size_t c1) {
    return guarded([&] { return MX(a).slice(r0, r1, c0, c1); });
}
SHELL_API int shell_slice_set(void* a, size_t r0, size_t r1, size_t c0, size_t c1, void* src) {
    g_err.clear(); g_shape = 0;
    try { MX(a).slice(r0, r1, c0, c1, MX(src)); return 0; }
    catch (const std::invalid_argument& e) { g_err = e.what(); g_shape = 1; }
    catch (const std::exception& e) { g_err = e.what(); }
    return 1;
}
SHELL_API void* shell_stack(int vertical, void** hs, int n) {
    return guarded([&]() -> CM {
        std::vector<MatrixView<CUDAfloat>> v;
        for (int i = 0; i < n; i++) v.push_back(MatrixView<CUDAfloat>(MX(hs[i])));
        if (vertical) return vstack(v);
        return hstack(v);
    });
}
// the README / config-1 expression and the testbasic golden expression, written exactly as a C++ user writes them
SHELL_API void* shell_expr_softplus5(void* a, void* b, double div) {
    return guarded([&]() -> CM { return log(exp(MX(a) * MX(b) / div) + 1.0f) / 5.0f; });
}
SHELL_API void* shell_expr_testbasic(void* a, void* b) {
    return guarded([&]() -> CM {
        CM& A = MX(a);
        CM& B = MX(b);
        return log(exp(-A / B) + exp(hadmd(B, A))) - (A.T() * B).rows(0, 2);   // tests/testbasic.cu:28-55
    });
}
// the softmax-CE head: the reference's OWN LogisticLayer<CUDAfloat> (ml/layer.hpp:232-283, unchanged) on logits X and
// one-hot Y -- grad() is the ~18-operator sequence the shell's lazy layer recognises and runs as one kernel; eval() the loss
SHELL_API void* shell_logistic_grad(void* x, void* y) {
    return guarded([&]() -> CM {
        Juzhen::LogisticLayer<CUDAfloat> head(MX(x).num_col(), MX(y));
        return head.grad(MX(x));
    });
}
SHELL_API void* shell_logistic_loss(void* x, void* y) {
    return guarded([&]() -> CM {
        Juzhen::LogisticLayer<CUDAfloat> head(MX(x).num_col(), MX(y));
        head.eval(MX(x));
        return head.value();
    });
}
}  // extern "C"
