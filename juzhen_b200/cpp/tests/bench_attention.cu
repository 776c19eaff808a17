// bench_attention.cu -- same-GPU comparison of the reference's OWN transformer helper kernels (ml/layer.hpp:2373-2538,
// compiled unchanged from the reference tree through the staged include path) against their jz_* replacements.
// Algorithmic bytes: softmax 8 B/elem (read + write), softmax backward 12 (A, dA^T, dS), LayerNorm forward 12
// (x, y, xhat), LayerNorm backward 12 (dy, xhat, dx).  Prints one line per case; run on the GPU box by
// scripts/gpu_attention_bench.sh.  Not a test: returns 0 unless results disagree grossly.
#include <cmath>
#include <cstdio>
#include <vector>

#include <cublas_v2.h>   // comparison bar only: the library call the reference makes (ml/layer.hpp:2896-2926)

#include "../cpp/juzhen.hpp"
#include "../ml/layer.hpp"

static float* dalloc(size_t n, unsigned long long seed) {
    float* p = nullptr;
    cudaMalloc(&p, n * sizeof(float));
    jz_rand_normal(p, n, seed, 0, nullptr);
    return p;
}
template <class F>
static double time_ms(F f, int reps) {
    for (int i = 0; i < 3; i++) f();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}
static double max_diff(const float* a, const float* b, size_t n) {
    std::vector<float> ha(n), hb(n);
    cudaMemcpy(ha.data(), a, n * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hb.data(), b, n * 4, cudaMemcpyDeviceToHost);
    double m = 0;
    for (size_t i = 0; i < n; i++) m = std::max(m, (double)std::fabs(ha[i] - hb[i]));
    return m;
}
static void report(const char* what, size_t S, size_t B, double bytes, double ms_ref, double ms_ours, double diff) {
    std::printf("%-22s %5zu x %-6zu reference %8.3f ms %7.1f GB/s | juzhen-b200 %8.3f ms %7.1f GB/s | x%.1f  max|diff| %.2e\n", what, S, B,
                ms_ref, bytes / ms_ref / 1e6, ms_ours, bytes / ms_ours / 1e6, ms_ref / ms_ours, diff);
}

int compute() {
    using namespace Juzhen;
    int bad = 0;
    const size_t cases[][2] = {{128, 512}, {256, 256}, {512, 128}, {1024, 32}};   // (seq_len, batch*heads)
    for (auto& c : cases) {
        const size_t S = c[0], B = c[1], n = S * S * B;
        float *x = dalloc(n, 1), *y0 = dalloc(n, 2), *y1 = dalloc(n, 3), *d = dalloc(n, 4), *g0 = dalloc(n, 5), *g1 = dalloc(n, 6);
        double t0 = time_ms([&] { cuda_softmax_rows_batched(x, y0, (int)S, (int)B); }, 20);
        double t1 = time_ms([&] { jz_softmax_rows_batched(y1, x, S, B, 0, -1e9f, nullptr); }, 20);
        double df = max_diff(y0, y1, n);
        report("softmax rows", S, B, 8.0 * n, t0, t1, df);
        bad += df > 1e-5;
        t0 = time_ms([&] { cuda_softmax_backward_rows(y0, d, g0, (int)S, (int)B, 0.125f); }, 20);
        t1 = time_ms([&] { jz_softmax_rows_backward(g1, y0, d, S, B, 0.125f, nullptr); }, 20);
        df = max_diff(g0, g1, n);
        report("softmax rows backward", S, B, 12.0 * n, t0, t1, df);
        bad += df > 1e-4;
        for (float* p : {x, y0, y1, d, g0, g1}) cudaFree(p);
    }
    // ---- the two strided-batched products of TransformerLayer's attention (ml/layer.hpp:2896-2926), one head's worth:
    //      scores = scale * Q^T K   (op T, N: m = n = seq, k = d_h, lda = ldb = d_k)   and   H = V A^T (op N, T: m = d_h, n = k = seq)
    //      jz_gemm_strided_batched (3xTF32 = fp32 accuracy, and TF32) against cublasSgemmStridedBatched in its default
    //      fp32 math (what the reference runs) and with TF32 tensor-op math (NVIDIA_TF32=1)
    {
        cublasHandle_t h;
        cublasCreate(&h);
        const size_t gcases[][3] = {{64, 128, 2048}, {128, 128, 1024}, {256, 128, 512}, {512, 128, 128}, {1024, 128, 32}, {256, 64, 512}};   // seq, d_h, batch
        for (auto& c : gcases) {
            const size_t S = c[0], DH = c[1], B = c[2], DK = DH;   // one head: d_k = d_h
            const size_t nq = DK * S * B, ns = S * S * B;
            float *q = dalloc(nq, 21), *k = dalloc(nq, 22), *v = dalloc(nq, 23), *s0 = dalloc(ns, 24), *s1 = dalloc(ns, 25), *h0 = dalloc(nq, 26), *h1 = dalloc(nq, 27);
            const float scale = 1.0f / std::sqrt((float)DH), zero = 0.0f, one = 1.0f;
            const double fl = 2.0 * S * S * DH * B;
            for (int pass = 0; pass < 2; pass++) {   // 0: QK^T, 1: V A^T
                auto cublas_call = [&] {
                    if (pass == 0) cublasSgemmStridedBatched(h, CUBLAS_OP_T, CUBLAS_OP_N, (int)S, (int)S, (int)DH, &scale, q, (int)DK, (long long)(DK * S), k, (int)DK,
                                                             (long long)(DK * S), &zero, s0, (int)S, (long long)(S * S), (int)B);
                    else cublasSgemmStridedBatched(h, CUBLAS_OP_N, CUBLAS_OP_T, (int)DH, (int)S, (int)S, &one, v, (int)DK, (long long)(DK * S), s0, (int)S,
                                                   (long long)(S * S), &zero, h0, (int)DK, (long long)(DK * S), (int)B);
                };
                auto ours_call = [&](int mode) {
                    if (pass == 0) jz_gemm_strided_batched(1, 0, S, S, DH, scale, q, DK, DK * S, k, DK, DK * S, 0.0f, s1, S, S * S, B, mode, nullptr);
                    else jz_gemm_strided_batched(0, 1, DH, S, S, 1.0f, v, DK, DK * S, s0, S, S * S, 0.0f, h1, DK, DK * S, B, mode, nullptr);
                };
                cublasSetMathMode(h, CUBLAS_DEFAULT_MATH);
                const double t_fp32 = time_ms(cublas_call, 20);
                cublasSetMathMode(h, CUBLAS_TF32_TENSOR_OP_MATH);
                const double t_tf32 = time_ms(cublas_call, 20);
                cublasSetMathMode(h, CUBLAS_DEFAULT_MATH);
                cublas_call();   // fp32 result as the comparator
                const double t_3x = time_ms([&] { ours_call(0); }, 20);
                const int path = jz_gemm_last_path();
                const double df = pass == 0 ? max_diff(s0, s1, ns) : max_diff(h0, h1, nq);
                const double t_1x = time_ms([&] { ours_call(1); }, 20);
                std::printf("batched gemm %-6s seq %4zu d_h %3zu batch %4zu | cuBLAS fp32 %7.3f ms %6.1f TF | cuBLAS TF32 %7.3f ms %6.1f TF | juzhen-b200 3xTF32 %7.3f ms %6.1f TF (x%.1f vs fp32, path %d, max|diff| %.1e) | TF32 %7.3f ms %6.1f TF (x%.2f vs cuBLAS TF32)\n",
                            pass == 0 ? "Q^T K" : "V A^T", S, DH, B, t_fp32, fl / t_fp32 / 1e9, t_tf32, fl / t_tf32 / 1e9, t_3x, fl / t_3x / 1e9, t_fp32 / t_3x, path, df,
                            t_1x, fl / t_1x / 1e9, t_tf32 / t_1x);
                bad += df > 1e-3 * std::sqrt((double)(pass == 0 ? DH : S));
            }
            for (float* p : {q, k, v, s0, s1, h0, h1}) cudaFree(p);
        }
        cublasDestroy(h);
    }
    const size_t ln[][2] = {{256, 65536}, {512, 32768}, {1024, 16384}, {4096, 4096}};   // (dim, tokens)
    for (auto& c : ln) {
        const size_t D = c[0], N = c[1], n = D * N;
        float *x = dalloc(n, 7), *ga = dalloc(D, 8), *be = dalloc(D, 9);
        float *y0 = dalloc(n, 10), *h0 = dalloc(n, 11), *i0 = dalloc(N, 12), *y1 = dalloc(n, 13), *h1 = dalloc(n, 14), *i1 = dalloc(N, 15);
        float *dy = dalloc(n, 16), *dx0 = dalloc(n, 17), *dx1 = dalloc(n, 18);
        double t0 = time_ms([&] { cuda_layernorm_forward(x, ga, be, y0, h0, i0, (int)D, (int)N); }, 10);
        double t1 = time_ms([&] { jz_layernorm_forward(y1, h1, i1, x, ga, be, D, N, nullptr); }, 10);
        double df = max_diff(y0, y1, n);
        report("layernorm forward", D, N, 12.0 * n, t0, t1, df);
        bad += df > 1e-3;
        t0 = time_ms([&] { cuda_layernorm_backward(dy, ga, h0, i0, dx0, (int)D, (int)N); }, 10);
        t1 = time_ms([&] { jz_layernorm_backward(dx1, dy, ga, h0, i0, D, N, nullptr); }, 10);
        df = max_diff(dx0, dx1, n);
        report("layernorm backward", D, N, 12.0 * n, t0, t1, df);
        bad += df > 1e-2;
        for (float* p : {x, ga, be, y0, h0, i0, y1, h1, i1, dy, dx0, dx1}) cudaFree(p);
    }
    return bad ? 1 : 0;
}
