// bench_attention.cu -- same-GPU comparison of the reference's OWN transformer helper kernels (ml/layer.hpp:2373-2538,
// compiled unchanged from the reference tree through the staged include path) against their jz_* replacements.
// Algorithmic bytes: softmax 8 B/elem (read + write), softmax backward 12 (A, dA^T, dS), LayerNorm forward 12
// (x, y, xhat), LayerNorm backward 12 (dy, xhat, dx).  Prints one line per case; run on the GPU box by
// scripts/gpu_attention_bench.sh.  Not a test: returns 0 unless results disagree grossly.
#include <cmath>
#include <cstdio>
#include <vector>

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
