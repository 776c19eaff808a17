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
#include <cstdlib>
#include <cstring>

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

// ------------------------------------------------------------------ deferred evaluation (jz_lazy.hpp)
namespace jzb200 {

bool lazy_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("JZ_EAGER");
        return !(e && *e && std::strcmp(e, "0") != 0);
    }();
    return on;
}

Storage::Storage(size_t n) : ptr(reinterpret_cast<float*>(Memory<CUDAfloat>::allocate(n))), count(n) {}
Storage::~Storage() { Memory<CUDAfloat>::free(reinterpret_cast<CUDAfloat*>(ptr)); }

// An unread product is LAUNCHED when the named matrix holding it is assigned over, so that timing loops of the form
// `out = a * b` (tests/benchmarkCoreOps.cu:64-69, the CPU loop of examples/demo_gemm.cu) measure what they mean
// to.  It is dropped only when it dies as a temporary or at scope exit -- e.g. the input gradient W^T*t of the first
// layer, which backprop() returns and examples/demo_mnist.cu:120 discards.  Dead elementwise work, fills and draws
// are always dropped: that is the point of deferring them.
void Storage::retire_by_assignment() {
    if (producer && producer->kind == Producer::GEMM && !producer->consumed) materialize();
}

bool Storage::lazy_ok() const { return lazy_enabled() && !escaped; }

void add_reader(const StoragePtr& source, const StoragePtr& reader) {
    auto& v = source->readers;
    v.erase(std::remove_if(v.begin(), v.end(), [](const std::weak_ptr<Storage>& w) { return w.expired(); }), v.end());
    v.push_back(reader);
}

// one pass over `count` floats applying `prog` (any length; single steps use the specialised kernels)
static void run_program(float* out, const float* in, size_t count, const std::vector<jz_step>& prog) {
    if (prog.empty()) {
        if (out != in) JZ_DO(jz_copy(out, in, count, S()));
        return;
    }
    size_t at = 0;
    while (at < prog.size()) {
        const size_t len = std::min(prog.size() - at, size_t(JZ_MAX_CHAIN));
        if (len == 1) {
            const jz_step& st = prog[at];
            if (st.kind == JZ_STEP_AFFINE) JZ_DO(jz_affine(out, in, count, st.s1, st.a, S()));
            else if (st.kind == JZ_STEP_ELEMINV) JZ_DO(jz_eleminv(out, in, count, st.s1, S()));
            else JZ_DO(jz_unary(st.kind, out, in, count, S()));
        } else {
            JZ_DO(jz_chain(out, in, count, prog.data() + at, int(len), S()));
        }
        in = out;  // later chunks continue in place
        at += len;
    }
}

static void run_gemm(Producer& g, float* out, const std::vector<jz_step>& epilogue) {
    g.a->materialize();
    g.b->materialize();
    const size_t fused = std::min(epilogue.size(), size_t(JZ_MAX_CHAIN));
    if (g.bias) {   // product + broadcast add + elementwise program: one kernel
        g.bias->materialize();
        JZ_DO(jz_gemm_bias_chain(g.ta, g.tb, g.m, g.n, g.k, 1.0f, g.a->ptr, g.lda, g.b->ptr, g.ldb, out, g.m ? g.m : 1, g.bias->ptr,
                                 g.bias_dim, g.bias_s1, g.bias_s2, epilogue.data(), int(fused), -1, S()));
    } else {
        JZ_DO(jz_gemm_chain(g.ta, g.tb, g.m, g.n, g.k, 1.0f, g.a->ptr, g.lda, g.b->ptr, g.ldb, out, g.m ? g.m : 1,
                            epilogue.data(), int(fused), -1, S()));
    }
    if (fused < epilogue.size())
        run_program(out, out, g.m * g.n, std::vector<jz_step>(epilogue.begin() + fused, epilogue.end()));
}

static void run_generator(const Producer& p, float* out, size_t count) {
    if (p.kind == Producer::FILL) JZ_DO(jz_fill(out, count, p.value, S()));
    else if (p.normal) JZ_DO(jz_rand_normal(out, count, p.seed, p.offset, S()));
    else JZ_DO(jz_rand_uniform(out, count, p.seed, p.offset, S()));
}

// ---- the column-softmax head, stage by stage (see Producer::Kind).  Each function computes its stage from `src` with
//      the kernels the eager path would have used, so a stage that IS read mid-way costs what it always cost.
static void run_colmax(const Producer& p, float* out) {
    p.src->materialize();
    JZ_DO(jz_max(out, p.src->ptr, p.rows, p.cols, p.rows ? p.rows : 1, 0, S()));
}
static void run_shifted(const Producer& p, float* out) {   // x - colmax(x)
    p.src->materialize();
    float* mx = reinterpret_cast<float*>(Memory<CUDAfloat>::allocate(p.cols));
    JZ_DO(jz_max(mx, p.src->ptr, p.rows, p.cols, p.rows ? p.rows : 1, 0, S()));
    JZ_DO(jz_add_bcast(out, p.src->ptr, p.rows, p.cols, mx, 0, 1.0f, -1.0f, S()));
    Memory<CUDAfloat>::free(reinterpret_cast<CUDAfloat*>(mx));
}
static void run_colsumexp(const Producer& p, float* out) {   // sum_i exp(x_ij - colmax_j)
    float* e = reinterpret_cast<float*>(Memory<CUDAfloat>::allocate(p.rows * p.cols));
    run_shifted(p, e);
    JZ_DO(jz_unary(JZ_EXP, e, e, p.rows * p.cols, S()));
    JZ_DO(jz_sum(out, e, p.rows, p.cols, p.rows ? p.rows : 1, 0, S()));
    Memory<CUDAfloat>::free(reinterpret_cast<CUDAfloat*>(e));
}
// softmax(src) [* ax_s1 + ax_s2 * other], then the pending program.  The loss-gradient form -(Y - S)/nb
// (LogisticLayer::grad, ml/layer.hpp:263) is ONE kernel; anything else is the softmax kernel plus the generic passes.
static void run_softmax(Producer& p, float* out, std::vector<jz_step>& prog) {
    p.src->materialize();
    const size_t n = p.rows * p.cols;
    if (p.other) {
        p.other->materialize();
        const bool ce = p.ax_s1 == -1.0f && p.ax_s2 == 1.0f && prog.size() == 2 && prog[0].kind == JZ_STEP_AFFINE && prog[0].s1 == -1.0f &&
                        prog[0].a == 0.0f && prog[1].kind == JZ_STEP_AFFINE && prog[1].a == 0.0f;
        if (ce) {
            JZ_DO(jz_softmax_ce_grad_scaled(out, p.src->ptr, p.other->ptr, p.rows, p.cols, prog[1].s1, S()));
            prog.clear();
            return;
        }
        JZ_DO(jz_softmax_cols(out, p.src->ptr, p.rows, p.cols, p.rows ? p.rows : 1, S()));
        JZ_DO(jz_axpby(out, out, p.other->ptr, n, p.ax_s1, p.ax_s2, S()));
    } else {
        JZ_DO(jz_softmax_cols(out, p.src->ptr, p.rows, p.cols, p.rows ? p.rows : 1, S()));
    }
}

void Storage::materialize() {
    if (producer && producer->kind >= Producer::COLMAX) {
        std::unique_ptr<Producer> p = std::move(producer);
        std::vector<jz_step> prog;
        prog.swap(pending);
        switch (p->kind) {
            case Producer::COLMAX: run_colmax(*p, ptr); break;
            case Producer::SHIFTED: run_shifted(*p, ptr); break;
            case Producer::COLSUMEXP: run_colsumexp(*p, ptr); break;
            default: run_softmax(*p, ptr, prog); break;
        }
        if (!prog.empty()) run_program(ptr, ptr, count, prog);
        return;
    }
    if (producer && (producer->kind == Producer::FILL || producer->kind == Producer::RAND)) {
        std::unique_ptr<Producer> p = std::move(producer);
        run_generator(*p, ptr, count);   // then the pending in-place program, below
    }
    if (producer) {
        std::unique_ptr<Producer> p = std::move(producer);
        std::vector<jz_step> prog;
        if (p->kind == Producer::GEMM) {
            prog.swap(pending);
            run_gemm(*p, ptr, prog);
            return;
        }
        // The producer holds the only remaining handle on its source: the source was a temporary that has
        // since died, so nobody will ever ask for ITS bytes -- fold its unfinished work into this pass.
        // A dead source that is itself "steps applied to X" is spliced out entirely.
        while (p->src.use_count() == 1 && p->src->producer && p->src->producer->kind == Producer::MAP) {
            Storage& dead = *p->src;
            std::vector<jz_step> steps = dead.producer->steps;
            steps.insert(steps.end(), dead.pending.begin(), dead.pending.end());
            steps.insert(steps.end(), p->steps.begin(), p->steps.end());
            p->steps.swap(steps);
            StoragePtr next = dead.producer->src;
            p->src = next;  // releases the dead temporary (its buffer goes back to the pool)
        }
        Storage& src = *p->src;
        const bool orphan = p->src.use_count() == 1;
        if (orphan && src.producer && src.producer->kind == Producer::GEMM) {
            std::unique_ptr<Producer> g = std::move(src.producer);
            prog = src.pending;
            src.pending.clear();
            prog.insert(prog.end(), p->steps.begin(), p->steps.end());
            prog.insert(prog.end(), pending.begin(), pending.end());
            pending.clear();
            run_gemm(*g, ptr, prog);  // the product lands directly in this buffer, program as its epilogue
            return;
        }
        if (!(orphan && !src.producer)) src.materialize();  // a live source keeps its own bytes up to date
        prog = src.pending;                                  // (empty unless orphan)
        prog.insert(prog.end(), p->steps.begin(), p->steps.end());
        prog.insert(prog.end(), pending.begin(), pending.end());
        pending.clear();
        run_program(ptr, src.ptr, count, prog);
        return;
    }
    if (!pending.empty()) {
        std::vector<jz_step> prog;
        prog.swap(pending);
        run_program(ptr, ptr, count, prog);
    }
}

void Storage::flush_readers() {
    if (readers.empty()) return;
    std::vector<std::weak_ptr<Storage>> rs;
    rs.swap(readers);
    for (auto& w : rs)
        if (StoragePtr r = w.lock())
            if (r->producer) r->materialize();
}

void Storage::before_write() {
    flush_readers();
    materialize();
    konst_known = false;
}

void Storage::append(const jz_step& s) {
    flush_readers();  // they were defined on the value before this step
    konst_known = false;
    if (!lazy_ok()) {
        materialize();
        run_program(ptr, ptr, count, std::vector<jz_step>{s});
        return;
    }
    if (pending.size() >= 2 * size_t(JZ_MAX_CHAIN)) materialize();
    pending.push_back(s);
}

void Storage::define(std::unique_ptr<Producer> p) {
    flush_readers();   // deferred readers were defined on the old value
    pending.clear();   // ... which nobody else can ask for any more
    producer.reset();
    konst_known = p->kind == Producer::FILL;
    konst = p->value;
    if (lazy_ok()) {
        producer = std::move(p);
        return;
    }
    run_generator(*p, ptr, count);
}

float* Storage::escape() {
    before_write();
    escaped = true;
    return ptr;
}

}  // namespace jzb200

using jzb200::Producer;
using jzb200::StoragePtr;

static inline jz_step step_unary(int op) { return jz_step{op, 0.0f, 0.0f}; }
static inline jz_step step_affine(float s1, float a) { return jz_step{JZ_STEP_AFFINE, s1, a}; }
static inline jz_step step_eleminv(float l) { return jz_step{JZ_STEP_ELEMINV, l, 0.0f}; }

// ------------------------------------------------------------------ storage and lifetime
jzb200::LazyBuf<CUDAfloat> Matrix<CUDAfloat>::new_storage(size_t count) {
    return jzb200::LazyBuf<CUDAfloat>(StoragePtr(new jzb200::Storage(count)));
}

// out-of-place elementwise result: defined as "step applied to this storage", computed when needed
Matrix<CUDAfloat> Matrix<CUDAfloat>::mapped(const char* nm, const jz_step& step) const {
    Matrix<CUDAfloat> R(Raw{}, nm, numrow, numcol, transpose);
    if (store().lazy_ok()) {
        std::unique_ptr<Producer> p(new Producer());
        p->kind = Producer::MAP;
        p->src = elements.storage();
        p->steps.push_back(step);
        R.store().producer = std::move(p);
        jzb200::add_reader(elements.storage(), R.elements.storage());
    } else {
        jzb200::run_program(R.store().ptr, dev(), count(), std::vector<jz_step>{step});
    }
    return R;
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
    // staged through pinned memory: M's buffer is free again on return (as after the reference's synchronous
    // cudaMemcpy) but the host does not wait for the device to drain first
    JZ_DO(jz_upload(store().ptr, M.elements.get(), count(), S()));
}

Matrix<CUDAfloat>::Matrix(const Matrix<CUDAfloat>& M)
    : Matrix(Raw{}, ("copy of" + M.name).c_str(), M.numrow, M.numcol, M.transpose) {
    JZ_DO(jz_memcpy_d2d(store().ptr, M.dev(), count(), S()));
}

Matrix<CUDAfloat>::Matrix(Matrix<CUDAfloat>&& M) noexcept
    : numcol(M.numcol), numrow(M.numrow), transpose(M.transpose), name(std::move(M.name)),
      elements(std::move(M.elements)) {
    M.elements = nullptr;
}

Matrix<CUDAfloat>& Matrix<CUDAfloat>::operator=(const Matrix<CUDAfloat>& M) {
    if (this == &M) return *this;
    name = "copy of " + M.name;
    const float* from = M.dev();
    if (elements && elements.use_count() == 1) store().retire_by_assignment();
    elements = new_storage(M.count());  // a fresh block, like the reference (cpp/cumatrix.cu:99-116)
    numrow = M.numrow;
    numcol = M.numcol;
    transpose = M.transpose;
    JZ_DO(jz_memcpy_d2d(store().ptr, from, count(), S()));
    return *this;
}

Matrix<CUDAfloat>& Matrix<CUDAfloat>::operator=(Matrix<CUDAfloat>&& M) noexcept {
    if (this == &M) return *this;
    if (elements && elements.use_count() == 1) store().retire_by_assignment();
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
void Matrix<CUDAfloat>::ones() { fill(*this, 1.0); }
void Matrix<CUDAfloat>::zeros() { fill(*this, 0.0); }

Matrix<CUDAfloat>& fill(Matrix<CUDAfloat>& M, double a) {
    std::unique_ptr<Producer> p(new Producer());
    p->kind = Producer::FILL;
    p->value = float(a);
    M.store().define(std::move(p));
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
    std::unique_ptr<Producer> p(new Producer());
    p->kind = Producer::RAND;
    p->normal = true;
    p->seed = GPUSampler::seed;
    p->offset = GPUSampler::offset;
    M.store().define(std::move(p));
    GPUSampler::offset += m * n;
    return M;
}

Matrix<CUDAfloat> Matrix<CUDAfloat>::rand(size_t m, size_t n) {
    Matrix<CUDAfloat> M(Raw{}, "rand", m, n, false);
    std::unique_ptr<Producer> p(new Producer());
    p->kind = Producer::RAND;
    p->seed = GPUSampler::seed;
    p->offset = GPUSampler::offset;
    M.store().define(std::move(p));
    GPUSampler::offset += m * n;
    return M;
}

// ------------------------------------------------------------------ GEMM (cpp/cumatrix.cu:177-197)
Matrix<CUDAfloat> Matrix<CUDAfloat>::dot(const Matrix<CUDAfloat>& B) const {
    if (num_col() != B.num_row()) throw std::invalid_argument("Matrix dimensions are not compatible");
    const size_t m = num_row(), n = B.num_col(), k = num_col();
    Matrix<CUDAfloat> C(Raw{}, "dot", m, n, false);
    if (store().lazy_ok() && B.store().lazy_ok()) {
        // defined now, launched when somebody needs the bytes: an elementwise chain applied to the result in
        // the meantime becomes the GEMM's epilogue (one kernel for log(exp(A*B)+1)/5)
        std::unique_ptr<Producer> g(new Producer());
        g->kind = Producer::GEMM;
        g->a = elements.storage();
        g->b = B.elements.storage();
        g->ta = transpose;
        g->tb = B.transpose;
        g->m = m; g->n = n; g->k = k;
        g->lda = numrow; g->ldb = B.numrow;
        C.store().producer = std::move(g);
        jzb200::add_reader(elements.storage(), C.elements.storage());
        if (B.elements.storage() != elements.storage()) jzb200::add_reader(B.elements.storage(), C.elements.storage());
        return C;
    }
    JZ_DO(jz_gemm(transpose, B.transpose, m, n, k, 1.0f, dev(), numrow, B.dev(), B.numrow, 0.0f, C.store().ptr,
                  m ? m : 1, -1, S()));
    return C;
}

// ------------------------------------------------------------------ affine / axpby / reciprocal
// s1*M + a (cpp/cumatrix.cu:199-215); layout and flag preserved
Matrix<CUDAfloat> Matrix<CUDAfloat>::add(float a, float s1) const { return mapped("add", step_affine(s1, a)); }
void Matrix<CUDAfloat>::add(float a, float s1) { store().append(step_affine(s1, a)); }

// M *= s1.  The CPU oracle's scale is add(0, s1) (cpp/core.hpp:148-149): s1*x + 0.0f.
void Matrix<CUDAfloat>::scale(float s1) { store().append(step_affine(s1, 0.0f)); }

// s1*this + s2*B into a fresh, non-transposed matrix of the logical shape (cpp/cumatrix.cu:227-244)
Matrix<CUDAfloat> Matrix<CUDAfloat>::add(const Matrix<CUDAfloat>& B, float s1, float s2) const {
    require_same_shape(*this, B);
    const size_t r = num_row(), c = num_col();
    Matrix<CUDAfloat> C(Raw{}, "add", r, c, false);
    float* out = C.store().ptr;
    if (!transpose && !B.transpose) JZ_DO(jz_axpby(out, dev(), B.dev(), count(), s1, s2, S()));
    else JZ_DO(jz_axpby2d(out, r ? r : 1, r, c, dev(), numrow, transpose, B.dev(), B.numrow, B.transpose, s1, s2, S()));
    return C;
}

// The reference spells "add a column (row) vector to every column (row)" as a rank-1 GEMM against a vector of
// ones followed by a full add: W*x + b*ones(1,N) (ml/layer.hpp:79,120), input - oneK1*mx (ml/layer.hpp:260),
// m = ones(n,1)*mean (cpp/juzhen.hpp:88-104).  When one side of an in-place add is such a product that has not
// been computed yet, the ones vector, the m x n outer product and the separate add collapse into ONE broadcast
// pass.  Bit-identical to the unfused order: u[i]*1.0f is exact, and the pass applies the same
// separately rounded s1*x + s2*y.
//   returns the storage of the vector and dim (1: out(i,j) gets vec[i]; 0: vec[j]) if `st` is a deferred u*1^T / 1*v^T
static StoragePtr deferred_broadcast(const jzb200::Storage& st, size_t rows, size_t cols, int& dim) {
    const Producer* g = st.producer.get();
    if (!g || g->kind != Producer::GEMM || g->k != 1 || !st.pending.empty() || g->m != rows || g->n != cols) return nullptr;
    const size_t su = g->ta ? g->lda : 1, sv = g->tb ? 1 : g->ldb;   // element strides of u = op(A)(:,0), v = op(B)(0,:)
    auto is_one = [](const StoragePtr& s) { return s->konst_known && s->konst == 1.0f && s->pending.empty(); };
    if (is_one(g->b) && (su == 1 || rows == 1)) { dim = 1; return g->a; }
    if (is_one(g->a) && (sv == 1 || cols == 1)) { dim = 0; return g->b; }
    return nullptr;
}

bool Matrix<CUDAfloat>::add_broadcast(const Matrix<CUDAfloat>& B, float s1, float s2) {
    if (transpose || B.transpose || !jzb200::lazy_enabled() || count() == 0) return false;
    const StoragePtr mine = elements.storage(), theirs = B.elements.storage();
    if (mine == theirs) return false;
    int dim = 0;
    // this = s1*this + s2*(u 1^T): one pass over this
    if (const Producer* g = theirs->producer.get(); g && g->a != mine && g->b != mine) {
        if (StoragePtr vec = deferred_broadcast(*theirs, numrow, numcol, dim)) {
            // `this` is itself a product that has not run yet (W*x + b*ones(1,N)): the broadcast becomes a stage of
            // its epilogue, and whatever elementwise steps follow (tanh, d_tanh) join it there -- one kernel in all
            if (mine->lazy_ok() && vec != mine) {
                mine->flush_readers();   // deferred readers were defined on the product without the broadcast
                Producer* mg = mine->producer.get();
                if (mg && mg->kind == Producer::GEMM && mg->k > 1 && !mg->bias && mine->pending.empty() && mg->a != vec && mg->b != vec) {
                    mg->bias = vec;
                    mg->bias_dim = dim;
                    mg->bias_s1 = s1;
                    mg->bias_s2 = s2;
                    jzb200::add_reader(vec, mine);   // if the vector changes first, the product runs with today's values
                    if (theirs->producer) theirs->producer->consumed = true;
                    return true;
                }
            }
            vec->materialize();
            float* x = wdev();
            JZ_DO(jz_add_bcast(x, x, numrow, numcol, vec->ptr, dim, s1, s2, S()));
            if (theirs->producer) theirs->producer->consumed = true;   // still defined (a later reader can run it)
            return true;
        }
    }
    // this = s1*(u 1^T) + s2*B, this being the not-yet-computed product: one pass from B into this buffer
    if (StoragePtr vec = deferred_broadcast(*mine, numrow, numcol, dim)) {
        if (vec == theirs || !mine->lazy_ok()) return false;
        // X - ones(K,1) * colmax(X) with the column max itself still deferred: define "X shifted by its column max"
        // (first stage of the softmax head, Producer::SHIFTED) instead of running the max and the broadcast now
        if (dim == 0 && s1 == -1.0f && s2 == 1.0f && vec->producer && vec->producer->kind == Producer::COLMAX && vec->pending.empty() &&
            vec->producer->src == theirs && theirs->lazy_ok() && vec->producer->rows == numrow && vec->producer->cols == numcol) {
            mine->flush_readers();
            if (mine->producer && mine->producer->kind == Producer::GEMM) {
                std::unique_ptr<Producer> p(new Producer());
                p->kind = Producer::SHIFTED;
                p->src = theirs;
                p->rows = numrow;
                p->cols = numcol;
                mine->producer = std::move(p);
                mine->konst_known = false;
                jzb200::add_reader(theirs, mine);
                return true;
            }
        }
        const float* b = B.dev();
        mine->flush_readers();
        if (!mine->producer) return false;   // a deferred reader needed the product itself after all
        vec->materialize();
        mine->producer.reset();
        JZ_DO(jz_add_bcast(mine->ptr, b, numrow, numcol, vec->ptr, dim, s2, s1, S()));
        return true;
    }
    return false;
}

// in place: the result keeps this layout, B is read transposed iff the flags differ (cpp/cumatrix.cu:246-260)
void Matrix<CUDAfloat>::add(const Matrix<CUDAfloat>& B, float s1, float s2) {
    require_same_shape(*this, B);
    if (add_broadcast(B, s1, s2)) return;
    {   // Y - softmax(X) with the softmax still deferred: it becomes part of the softmax pass (Producer::SOFTMAX)
        jzb200::Storage& st = store();
        Producer* p = st.producer.get();
        if (p && p->kind == Producer::SOFTMAX && !p->other && st.pending.empty() && st.lazy_ok() && !transpose && !B.transpose &&
            B.elements.storage() != elements.storage() && B.elements.storage() != p->src) {
            st.flush_readers();
            if (st.producer.get() == p) {
                p->other = B.elements.storage();
                p->ax_s1 = s1;
                p->ax_s2 = s2;
                jzb200::add_reader(B.elements.storage(), elements.storage());
                return;
            }
        }
    }
    const float* b = B.dev();  // before wdev(): B may be a deferred view of this very storage
    float* x = wdev();
    if (transpose == B.transpose) JZ_DO(jz_axpby(x, x, b, count(), s1, s2, S()));
    else JZ_DO(jz_axpby2d(x, numrow ? numrow : 1, numrow, numcol, x, numrow, 0, b, B.numrow, 1, s1, s2, S()));
}

// l / M (cpp/cumatrix.cu:263-303)
Matrix<CUDAfloat> Matrix<CUDAfloat>::eleminv(double l) const { return mapped("elem_rec", step_eleminv(float(l))); }
void Matrix<CUDAfloat>::eleminv(double l) { store().append(step_eleminv(float(l))); }

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
    JZ_DO(jz_copy2d(W.store().ptr, r ? r : 1, dev() + cstart * numrow + rstart, numrow, r, c, 0, S()));
    return W;
}

// assignment into a window: M's buffer is taken in this matrix's physical orientation, as the
// reference's copyKernel does (cpp/cukernels.cu:72-90)
void Matrix<CUDAfloat>::slice(size_t rstart, size_t rend, size_t cstart, size_t cend, const Matrix<CUDAfloat>& M) {
    if (transpose) { std::swap(rstart, cstart); std::swap(rend, cend); }
    const size_t r = rend - rstart, c = cend - cstart;
    const float* from = M.dev();
    // Assignment in LOGICAL coordinates, the CPU oracle's semantics (cpp/core.hpp:365-373: elem(i, j) on both sides).  The
    // reference's GPU copyKernel reads M's physical buffer whatever M's flag (cpp/cumatrix.cuh:208-222), which is the same
    // thing whenever the two flags agree; when they differ the block is M's physical buffer transposed.
    const bool mixed = M.transpose != transpose;
    JZ_DO(jz_copy2d(wdev() + cstart * numrow + rstart, numrow, from, M.numrow ? M.numrow : 1, r, c, mixed ? 1 : 0, S()));
}

void copy(Matrix<CUDAfloat>& dest, const Matrix<CUDAfloat>& src) {  // cpp/cukernels.cu:241-256: dest is re-allocated
    dest.numrow = src.numrow;
    dest.numcol = src.numcol;
    dest.transpose = src.transpose;
    const float* from = src.dev();
    dest.elements.reset();
    dest.elements = Matrix<CUDAfloat>::new_storage(src.count());
    JZ_DO(jz_copy(dest.store().ptr, from, src.count(), S()));
}

// ------------------------------------------------------------------ reductions
// sum(M, 0): 1 x ncols, handed back as a flagged ncols x 1 buffer; sum(M, 1): nrows x 1 (cpp/cumatrix.cu:312-337)
Matrix<CUDAfloat> sum(const Matrix<CUDAfloat>& M, int dim) {
    const bool down_physical_columns = (dim == 0) != M.transpose;
    const size_t len = down_physical_columns ? M.numcol : M.numrow;
    Matrix<CUDAfloat> R(Matrix<CUDAfloat>::Raw{}, "sumM", len, 1, dim == 0);
    {   // sum(exp(X - colmax), 0) with the shifted matrix still deferred: the softmax denominators, also deferred
        jzb200::Storage& st = M.store();
        const Producer* p = st.producer.get();
        if (dim == 0 && !M.transpose && p && p->kind == Producer::SHIFTED && st.pending.size() == 1 && st.pending[0].kind == JZ_EXP &&
            st.lazy_ok()) {
            std::unique_ptr<Producer> q(new Producer());
            q->kind = Producer::COLSUMEXP;
            q->src = p->src;
            q->rows = p->rows;
            q->cols = p->cols;
            R.store().producer = std::move(q);
            jzb200::add_reader(p->src, R.elements.storage());
            return R;
        }
    }
    JZ_DO(jz_sum(R.store().ptr, M.dev(), M.numrow, M.numcol, M.numrow ? M.numrow : 1, down_physical_columns ? 0 : 1, S()));
    return R;
}

// ------------------------------------------------------------------ unary maps (cpp/cukernels.cu:156-239)
Matrix<CUDAfloat> jz_unary_new(int op, const char* name, const Matrix<CUDAfloat>& M) {
    return M.mapped(name, step_unary(op));
}
Matrix<CUDAfloat> jz_unary_reuse(int op, Matrix<CUDAfloat>&& M) {
    M.store().append(step_unary(op));
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
    float* out = R.store().ptr;
    if (M1.transpose == M2.transpose) JZ_DO(jz_hadamard(out, M1.dev(), M2.dev(), M1.count(), S()));
    else JZ_DO(jz_hadamard2d(out, M1.numrow ? M1.numrow : 1, M1.numrow, M1.numcol, M1.dev(), M1.numrow, 0, M2.dev(),
                             M2.numrow, 1, S()));
    return R;
}
// E / Z of the softmax head: M1 = exp(X - colmax) still deferred, M2 = 1 / (ones(K,1) * sum(E,0)) still deferred -> M2's
// storage is redefined as softmax(X) (Producer::SOFTMAX), nothing runs
static bool define_softmax(const Matrix<CUDAfloat>& M1, Matrix<CUDAfloat>& M2, jzb200::Storage& s1, jzb200::Storage& s2,
                           const StoragePtr& s2ptr) {
    const Producer* e = s1.producer.get();
    const Producer* g = s2.producer.get();
    if (!e || !g || e->kind != Producer::SHIFTED || g->kind != Producer::GEMM || !s1.lazy_ok() || !s2.lazy_ok()) return false;
    if (s1.pending.size() != 1 || s1.pending[0].kind != JZ_EXP) return false;
    if (s2.pending.size() != 1 || s2.pending[0].kind != JZ_STEP_ELEMINV || s2.pending[0].s1 != 1.0f) return false;
    if (g->k != 1 || g->bias || g->m != e->rows || g->n != e->cols) return false;
    const jzb200::Storage& ones = *g->a;
    if (!(ones.konst_known && ones.konst == 1.0f && ones.pending.empty())) return false;
    const jzb200::Storage& den = *g->b;
    if (!den.producer || den.producer->kind != Producer::COLSUMEXP || !den.pending.empty() || den.producer->src != e->src) return false;
    s2.flush_readers();
    if (s2.producer.get() != g) return false;
    std::unique_ptr<Producer> p(new Producer());
    p->kind = Producer::SOFTMAX;
    p->src = e->src;
    p->rows = e->rows;
    p->cols = e->cols;
    s2.pending.clear();
    s2.producer = std::move(p);
    s2.konst_known = false;
    jzb200::add_reader(s2.producer->src, s2ptr);
    (void)M1; (void)M2;
    return true;
}

Matrix<CUDAfloat> hadmd(const Matrix<CUDAfloat>& M1, Matrix<CUDAfloat>&& M2) {
    require_same_shape(M1, M2);
    if (M1.transpose != M2.transpose) return hadmd(M1, static_cast<const Matrix<CUDAfloat>&>(M2));
    if (!M1.transpose && M1.elements.storage() != M2.elements.storage() &&
        define_softmax(M1, M2, M1.store(), M2.store(), M2.elements.storage()))
        return std::move(M2);
    const float* other = M1.dev();
    float* x = M2.wdev();
    JZ_DO(jz_hadamard(x, x, other, M1.count(), S()));
    return std::move(M2);
}
Matrix<CUDAfloat> hadmd(Matrix<CUDAfloat>&& M1, const Matrix<CUDAfloat>& M2) {
    require_same_shape(M1, M2);
    if (M1.transpose != M2.transpose) return hadmd(static_cast<const Matrix<CUDAfloat>&>(M1), M2);
    const float* other = M2.dev();
    float* x = M1.wdev();
    JZ_DO(jz_hadamard(x, x, other, M1.count(), S()));
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
        JZ_DO(jz_copy2d(R.store().ptr + at * rows, rows, src, m.get_transpose() ? m.num_col() : rows, rows, m.num_col(),
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
        JZ_DO(jz_copy2d(R.store().ptr + at, rows, src, m.get_transpose() ? cols : m.num_row(), m.num_row(), cols,
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
    JZ_DO(jz_memcpy_h2d(M.wdev(), reinterpret_cast<const float*>(tmp.data()), n, S()));
    JZ_DO(jz_sync(S()));
}
