// jz_lazy.hpp -- deferred evaluation behind Matrix<CUDAfloat> ("fusion").
//
// The reference's operators are eager: `log(exp(A*B)+1.0f)/5.0f` is a GEMM, a zero-fill, and four
// separate passes over the result.  The API cannot change -- operators return Matrix by
// value -- so fusion hides in the storage object every Matrix handle points at:
//
//   * an in-place elementwise op (rvalue overloads, add(a,s1), scale, eleminv) only APPENDS a step to
//     the storage's pending program;
//   * an out-of-place elementwise op and dot() create a storage whose contents are DEFINED but not yet
//     computed (a Producer: "these steps applied to that storage" / "this GEMM");
//   * fills (the zero-init contract of the public constructor, ones()) and random draws are deferred too: a
//     matrix that is constructed and then replaced or destroyed without being read -- every dummy weight,
//     bias and Adam buffer of the per-iteration LogisticLayer in examples/demo_mnist.cu:118 -- costs no launch;
//   * an unread deferred PRODUCT is launched when the named matrix holding it is assigned over (a benchmark loop
//     `out = a * b` that never looks at `out` must still time the GEMM); it is dropped only when it dies as a
//     temporary or at scope exit;
//   * anything that needs real bytes (a read by a non-elementwise op, to_host, data(), printing)
//     materialises: the whole program runs as ONE jz_chain pass, or as the epilogue of ONE jz_gemm_chain
//     when the values come from a product whose un-materialised temporary has already died.
//
// Results are bit-identical to the eager order (jz_chain applies the same roundings step by step).
// Safety rules: a storage whose raw pointer escaped through data() is never deferred (callers launch
// their own kernels on it); before a storage's value changes, every deferred reader of it is
// materialised first.  JZ_EAGER=1 turns all of this off.
#pragma once
#include <jz_b200.h>

#include <cstddef>
#include <memory>
#include <vector>

namespace jzb200 {

struct Storage;
using StoragePtr = std::shared_ptr<Storage>;

struct Producer {
    // COLMAX .. SOFTMAX: the column-softmax head the reference spells with a dozen operators (LogisticLayer::grad,
    // ml/layer.hpp:252-264): mx = reduce(max, X, 0) -> X - 1*mx -> exp -> sum(.,0) -> 1*sum -> E / Z -> Y - S -> -(.)/nb.
    // Each stage is DEFINED here without running; when the chain completes, one jz_softmax_cols(-_axpby) pass computes
    // it from X.  A stage somebody reads in the meantime materialises on its own (see cumatrix.cu).
    enum Kind { GEMM, MAP, FILL, RAND, COLMAX, SHIFTED, COLSUMEXP, SOFTMAX } kind = MAP;
    // GEMM: C(m x n) = op(A)(m x k) * op(B)(k x n), column-major, lda/ldb = physical rows
    StoragePtr a, b;
    int ta = 0, tb = 0;
    size_t m = 0, n = 0, k = 0, lda = 0, ldb = 0;
    bool consumed = false;   // its value already went into a fused pass (broadcast add): need not run if it dies unread
    // optional broadcast stage of the GEMM epilogue: C = bias_s1*(A*B) + bias_s2*bias[i] (dim 1) / bias[j] (dim 0), the
    // W*x + b*ones(1,N) idiom (ml/layer.hpp:79,120) folded into the product that has not run yet
    StoragePtr bias;
    int bias_dim = 0;
    float bias_s1 = 1.0f, bias_s2 = 0.0f;
    // MAP: out[i] = steps(src[i]) over the flat physical buffer
    // COLMAX (1 x cols) / SHIFTED, SOFTMAX (rows x cols) / COLSUMEXP (cols): functions of the plain rows x cols matrix `src`
    StoragePtr src;
    size_t rows = 0, cols = 0;
    // SOFTMAX only: value = ax_s1 * softmax(src) + ax_s2 * other when `other` is set (the Y - S of the loss gradient)
    StoragePtr other;
    float ax_s1 = 1.0f, ax_s2 = 0.0f;
    std::vector<jz_step> steps;
    // FILL: every element = value.  RAND: counter-based stream (seed, offset) fixed when the matrix was made,
    // so the values do not depend on when -- or whether -- the kernel runs.
    float value = 0.0f;
    bool normal = false;
    unsigned long long seed = 0, offset = 0;
};

struct Storage {
    float* ptr = nullptr;
    size_t count = 0;
    bool escaped = false;                         // raw pointer handed out by data()
    bool konst_known = false;                     // every element is known to equal `konst` (set by fills, cleared
    float konst = 0.0f;                           // by anything that may change the bytes)
    std::vector<jz_step> pending;                 // deferred in-place program
    std::unique_ptr<Producer> producer;           // deferred definition of the contents
    std::vector<std::weak_ptr<Storage>> readers;  // storages whose producer reads this one

    explicit Storage(size_t count);
    ~Storage();
    Storage(const Storage&) = delete;
    Storage& operator=(const Storage&) = delete;

    bool lazy_ok() const;           // deferral allowed on this storage
    void materialize();             // contents become real bytes
    void flush_readers();           // deferred readers take their snapshot now
    void before_write();            // flush_readers + materialize: safe to modify the bytes in place
    void append(const jz_step& s);  // in-place elementwise step (deferred when allowed)
    void retire_by_assignment();    // the owning matrix is being assigned over: launch an unread product (see cumatrix.cu)
    void define(std::unique_ptr<Producer> p);  // the WHOLE buffer is redefined (fill / random draw): old work is dropped
    float* escape();                // materialise for an outside reader/writer; disables deferral for good
};

bool lazy_enabled();
void add_reader(const StoragePtr& source, const StoragePtr& reader);

// what Matrix<CUDAfloat>::elements is: a shared handle on a Storage that looks enough like the
// reference's std::shared_ptr<CUDAfloat[]> for MatrixView (cpp/core.hpp:60-64) -- get() materialises.
template <class Elem>
class LazyBuf {
    StoragePtr st;

   public:
    LazyBuf() = default;
    LazyBuf(std::nullptr_t) {}
    explicit LazyBuf(StoragePtr s) : st(std::move(s)) {}
    Elem* get() const {
        if (!st) return nullptr;
        st->materialize();
        return reinterpret_cast<Elem*>(st->ptr);
    }
    Elem& operator[](size_t i) const { return get()[i]; }
    explicit operator bool() const { return bool(st); }
    long use_count() const { return st.use_count(); }
    void reset() { st.reset(); }
    LazyBuf& operator=(std::nullptr_t) {
        st.reset();
        return *this;
    }
    const StoragePtr& storage() const { return st; }
};

}  // namespace jzb200
