// bench_overhead.cu -- host-side cost per operator of the C++ shell (launch-bound regime: 32 x 32 matrices, the device
// work per op is ~2 us).  Wall time per op over long loops, device drained only at the end; compiles against either
// backend (the reference's CUDA build or this one), so the two shells can be compared directly.
#include <chrono>
#include <cstdio>

#include "../cpp/juzhen.hpp"

template <class F>
static double us_per_op(F f, int iters) {
    for (int i = 0; i < 200; i++) f();
    cudaDeviceSynchronize();
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; i++) f();
    cudaDeviceSynchronize();
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / iters;
}

int compute() {
    const int N = 20000;
    auto hA = Matrix<float>::randn(32, 32), hB = Matrix<float>::randn(32, 32);
    CM A(hA), B(hB), C("c", 32, 32);
    std::printf("%-44s %8.2f us\n", "C = A + B            (alloc + 1 kernel)", us_per_op([&] { C = A + B; }, N));
    std::printf("%-44s %8.2f us\n", "C = hadmd(A, B)", us_per_op([&] { C = hadmd(A, B); }, N));
    std::printf("%-44s %8.2f us\n", "C = tanh(A)", us_per_op([&] { C = tanh(A); C.norm(); }, 2000));
    std::printf("%-44s %8.2f us\n", "C = tanh(A); C += B   (map + in-place add)", us_per_op([&] { C = tanh(A); C += B; }, N));
    std::printf("%-44s %8.2f us\n", "C = A * B            (32^3 product)", us_per_op([&] { C = A * B; C += B; }, N));
    std::printf("%-44s %8.2f us\n", "C = tanh(A * B + 1.0f) (product + chain)", us_per_op([&] { C = tanh(A * B + 1.0f); C += B; }, N));
    std::printf("%-44s %8.2f us\n", "s = sum(A, 1)", us_per_op([&] { CM s = sum(A, 1); C += B; }, N));
    std::printf("%-44s %8.2f us\n", "CM T(\"t\", 32, 32)    (ctor + dtor)", us_per_op([&] { CM T("t", 32, 32); }, N));
    std::printf("%-44s %8.2f us\n", "CM X(hA)             (upload)", us_per_op([&] { CM X(hA); }, N));
    return 0;
}
