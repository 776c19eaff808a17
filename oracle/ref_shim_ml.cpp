// oracle/ref_shim_ml.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Second shim over the UNMODIFIED reference headers, this one over ml/layer.hpp: the reference's own CPU formulation
// (generic Matrix<float> operators) of the transformer helpers whose CUDA kernels the jz_* entry points of SURVEY row
// 8f-3 replace -- row_softmax (ml/layer.hpp:2344-2367), LayerNorm<float>::forward / backward (:2572-2700) and the softmax
// backward expression of TransformerLayer::backward (:3351-3352).
// Used to pin oracle/jz_oracle.c's restatements of the CUDA kernels (jzo_softmax_rows_batched, jzo_layernorm_*) against
// what the reference itself computes on the CPU, and to generate tests/golden/ref_ml_golden.npz.
// Built by `make -C oracle ref-ml` into oracle/_ref/libjzref_ml.so (git-ignored).
#include "cpp/juzhen.hpp"
#include "ml/layer.hpp"

int compute() { return 0; }

namespace {
using MF = Matrix<float>;
MF owned(const float* p, size_t numrow, size_t numcol) {   // a copy the reference code may keep or move from
    MF m("in", numrow, numcol);
    std::memcpy((void*)m.data(), p, numrow * numcol * sizeof(float));
    return m;
}
void emit(const MF& r, float* out) {
    const size_t R = r.num_row(), C = r.num_col();
    for (size_t j = 0; j < C; j++)
        for (size_t i = 0; i < R; i++) out[j * R + i] = r.elem(i, j);
}
}  // namespace

extern "C" {

int refml_version() { return 1; }

// one seq x seq block: out(a, :) = softmax over keys of x(a, :)
int refml_row_softmax(const float* x, size_t rows, size_t cols, float* out) {
    emit(Juzhen::row_softmax(owned(x, rows, cols)), out);
    return 0;
}

int refml_layernorm_forward(const float* x, const float* gamma, const float* beta, size_t dim, size_t N, float* y,
                            float* xhat, float* inv_std) {
    Juzhen::LayerNorm<float> ln((int)dim, (int)N);
    ln.gamma = owned(gamma, dim, 1);
    ln.beta = owned(beta, dim, 1);
    emit(ln.forward(owned(x, dim, N)), y);
    emit(ln.cached_xhat, xhat);
    emit(ln.cached_inv, inv_std);
    return 0;
}

int refml_layernorm_backward(const float* dy, const float* gamma, const float* xhat, const float* inv_std, size_t dim,
                             size_t N, float* dx) {
    Juzhen::LayerNorm<float> ln((int)dim, (int)N);
    ln.gamma = owned(gamma, dim, 1);
    ln.cached_xhat = owned(xhat, dim, N);
    ln.cached_inv = owned(inv_std, 1, N);
    emit(ln.backward(owned(dy, dim, N), /*update=*/false), dx);
    return 0;
}

// dS = A .* (dA - rowsum(A .* dA) * ones(1, seq)) * scale: the reference's CPU spelling of the softmax backward of one
// attention block, written with its own operators exactly as in TransformerLayer::backward (ml/layer.hpp:3351-3352)
int refml_softmax_backward(const float* A, const float* dA, size_t seq, float scale, float* dS) {
    const MF Ai = owned(A, seq, seq), dAi = owned(dA, seq, seq);
    MF ones_1seq("ones", 1, seq);
    ones_1seq.ones();
    auto AodA = hadmd(Ai, dAi);
    auto row_sum = sum(AodA, 1);
    auto dSi = hadmd(Ai, dAi - row_sum * ones_1seq) * scale;
    emit(dSi, dS);
    return 0;
}

// One Adam step through the reference's own adam_update<float> (ml/util.cuh:165-257; the fused CPU branch :223-245,
// which the reference documents as bit-identical to its generic operator formulation).  g, m, v are updated in place;
// `iteration` is the step counter t (bias corrections 1/(1 - beta^t) are computed inside, in double).
int refml_adam_update(float* g, float* m, float* v, size_t n, float alpha, float beta1, float beta2, float eps, int iteration) {
    adam_state<float> st((double)alpha, n, 1);
    st.iteration = iteration;
    st.alpha = alpha; st.beta1 = beta1; st.beta2 = beta2; st.eps = eps;
    st.m = owned(m, n, 1);
    st.v = owned(v, n, 1);
    MF upd = adam_update(owned(g, n, 1), st);
    std::memcpy(g, upd.data(), n * sizeof(float));
    std::memcpy(m, st.m.data(), n * sizeof(float));
    std::memcpy(v, st.v.data(), n * sizeof(float));
    return 0;
}

}  // extern "C"
