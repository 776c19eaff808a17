// bench_mnist_step.cu -- BASELINE configs[3]: one demo_mnist training step (784 -> 1024 -> 128 -> 10 MLP with
// Layer / LinearLayer / LogisticLayer from the reference's UNCHANGED ml/layer.hpp, Adam inside backprop, exactly the
// loop body of examples/demo_mnist.cu:103-121) at a batch size taken from the environment (MNIST_BATCH = 32 | 8192 |
// 60000; the demo itself hard-codes 32).  Synthetic data of MNIST's shape: X = randn(784, N), Y = one-hot of N random
// labels.  Built twice from this one source -- against this backend (juzhen_b200/cpp/build_dropin.py) and against the
// reference's own CUDA/cuBLAS sources (oracle/build_ref_cuda.sh) -- and timed with CUDA events on the stream both
// backends use (legacy default), after warm-up.  FLOPs per step are counted as 6 * N * (#weights) (forward, input
// gradient and weight gradient of every layer), the usual convention.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <list>
#include <vector>

#include "../cpp/juzhen.hpp"
#include "../ml/layer.hpp"

using namespace std;
using namespace Juzhen;

int compute() {
    const char* e = std::getenv("MNIST_BATCH");
    const int N = e && *e ? std::atoi(e) : 32;
    const int d = 28 * 28, k = 10;
    GPUSampler sampler(1);
    // host-side synthetic batch, uploaded once (the demo re-uploads its 32-column slice every step; at N = 60000 the
    // whole training set IS one batch)
    Matrix<float> Xh("X", d, N), Yh("Y", k, N);
    Yh.zeros();
    {
        std::mt19937 gen(7);
        std::normal_distribution<float> nd(0.f, 1.f);
        for (int j = 0; j < N; j++) {
            for (int i = 0; i < d; i++) Xh.elem(i, j) = nd(gen);
            Yh.elem(gen() % k, j) = 1.0f;
        }
    }
    Matrix<CUDAfloat> X(Xh), Y(Yh);
    Layer<CUDAfloat> L0(1024, d, N), L1(128, 1024, N);
    LinearLayer<CUDAfloat> L2(k, 128, N);
    list<Layer<CUDAfloat>*> trainnn({&L2, &L1, &L0});
    auto step = [&] {
        forward(trainnn, X);
        LogisticLayer<CUDAfloat> L3(N, Y);
        trainnn.push_front(&L3);
        backprop(trainnn, X);
        trainnn.pop_front();
    };
    const int warm = 5, iters = N <= 64 ? 2000 : (N <= 8192 ? 100 : 20);
    for (int i = 0; i < warm; i++) step();
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const auto w0 = std::chrono::steady_clock::now();
    cudaEventRecord(e0, 0);
    for (int i = 0; i < iters; i++) step();
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    const double wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count() / iters;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    const double weights = double(d) * 1024 + 1024.0 * 128 + 128.0 * k;
    const double tflops = 6.0 * N * weights / (ms * 1e-3) / 1e12;
    // the training loss after the timed steps, as a sanity check that both builds learn the same problem
    LogisticLayer<CUDAfloat> L3(N, Y);
    forward(trainnn, X);
    L3.eval(L2.value());
    const float loss = L3.value().to_host().elem(0, 0);
    std::printf("bench_mnist_step batch=%d ms_per_step=%.4f wall_ms_per_step=%.4f TFLOP/s=%.2f steps=%d loss_after=%.4f\n", N, ms, wall_ms,
                tflops, iters + warm, loss);
    return 0;
}
