// test_fusion.cu -- acceptance test of the deferred-evaluation layer (jz_lazy.hpp) through the reference's
// public C++ API only.  Same conventions as the reference's tests: define compute(), return non-zero on
// failure.  Checks (1) values against the Matrix<float> CPU path of the same
// expression, (2) aliasing / ordering hazards that a lazy implementation could get wrong, (3) that the
// README expression log(exp(A*B)+1)/5 really runs as ONE GEMM launch with a fused epilogue, and writes a
// dump that tests/test_dropin_gpu.py compares bit-for-bit between JZ_EAGER=1 and the default lazy mode.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <random>
#include <string>
#include <vector>

#include "../cpp/juzhen.hpp"

static int failures = 0;
static std::vector<Matrix<float>> dump;

static float max_abs_diff(const Matrix<float>& a, const Matrix<float>& b) {
    if (a.num_row() != b.num_row() || a.num_col() != b.num_col()) return INFINITY;
    float m = 0.0f;
    for (size_t j = 0; j < a.num_col(); j++)
        for (size_t i = 0; i < a.num_row(); i++) m = std::max(m, std::fabs(a.elem(i, j) - b.elem(i, j)));
    return m;
}
static float rel_fro(const Matrix<float>& a, const Matrix<float>& b) {
    double num = 0, den = 0;
    for (size_t j = 0; j < a.num_col(); j++)
        for (size_t i = 0; i < a.num_row(); i++) {
            const double d = double(a.elem(i, j)) - double(b.elem(i, j));
            num += d * d;
            den += double(b.elem(i, j)) * double(b.elem(i, j));
        }
    return float(std::sqrt(num / (den > 0 ? den : 1)));
}
static void check(bool ok, const std::string& what) {
    std::cout << (ok ? "[PASS] " : "[FAIL] ") << what << std::endl;
    if (!ok) failures++;
}
static Matrix<float> keep(const Matrix<float>& m) {
    dump.push_back(m);
    return m;
}

int compute() {
    global_rand_gen.seed(7);
    GPUSampler sampler(7);
    const size_t n = 512;
    auto A = Matrix<float>::randn(n, n), B = Matrix<float>::randn(n, n);
    CM dA(A), dB(B);

    // (1) the README chain: one GEMM launch (plus its two operand pre-passes) and nothing else
    {
        jz_sync(nullptr);
        const unsigned long long before = jz_launch_count();
        CM R = log(exp(dA * dB / (float)n) + 1.0f) / 5.0f;
        Matrix<float> got = keep(R.to_host());
        const unsigned long long launches = jz_launch_count() - before;
        Matrix<float> want = log(exp(A * B / (float)n) + 1.0f) / 5.0f;
        std::cout << "    chain launches: " << launches << ", rel_fro vs CPU " << rel_fro(got, want) << std::endl;
        check(rel_fro(got, want) < 1e-5f, "log(exp(A*B/n)+1)/5 matches the CPU path (1e-5)");
        const char* eager = std::getenv("JZ_EAGER");
        if (!(eager && *eager && std::string(eager) != "0"))
            check(launches <= 3, "fused: GEMM + operand pre-passes only (<= 3 launches)");
    }
    // (2) pure elementwise rvalue chain on an existing matrix: one pass
    {
        const unsigned long long before = jz_launch_count();
        CM R = tanh(exp(-dA / 3.0f) * 0.5f - 1.0f);
        Matrix<float> got = keep(R.to_host());
        const unsigned long long launches = jz_launch_count() - before;
        Matrix<float> want = tanh(exp(-A / 3.0f) * 0.5f - 1.0f);
        std::cout << "    elementwise chain launches: " << launches << std::endl;
        check(max_abs_diff(got, want) < 2e-6f, "tanh(exp(-A/3)*0.5-1) matches the CPU path");
    }
    // (3) a deferred reader must see the OLD value of a source that is modified afterwards
    {
        CM X(A);
        CM Y = exp(X);       // lvalue overload: defined on X's current value
        X += 1.0f;           // in-place update of X
        Matrix<float> y = keep(Y.to_host()), x = keep(X.to_host());
        check(max_abs_diff(y, exp(A)) < 1e-5f * 50, "exp(X) taken before X += 1 sees the old X");
        check(max_abs_diff(x, A + 1.0f) == 0.0f, "X += 1 applied exactly once");
    }
    // (4) a deferred GEMM must use the operand values at the time of the call
    {
        CM P(A), Q(B);
        CM C = P * Q;
        P.zeros();
        Q = Q * 2.0f;
        Matrix<float> c = keep(C.to_host());
        check(rel_fro(c, A * B) < 1e-5f, "A*B evaluated with the operands as they were at the call");
    }
    // (5) T() aliases share pending work
    {
        CM X(A);
        auto XT = X.T();
        X.scale(2.0f);
        Matrix<float> xt = keep(XT.to_host());
        check(max_abs_diff(xt, (A * 2.0f).T()) == 0.0f, "T() view observes an in-place update of its buffer");
    }
    // (6) copies are snapshots
    {
        CM C = dA * dB;
        CM D = C;
        C = exp(std::move(C) / 100.0f);
        Matrix<float> d = keep(D.to_host()), c = keep(C.to_host());
        check(rel_fro(d, A * B) < 1e-5f, "copy of a deferred product is the product");
        check(rel_fro(c, exp(A * B / 100.0f)) < 1e-5f, "moved-from chain continues on the original buffer");
    }
    // (7) programs longer than one fused pass
    {
        CM X(A);
        Matrix<float> W = A;
        for (int i = 0; i < 21; i++) {
            X = tanh(std::move(X) * 1.5f + 0.25f);
            W = tanh(std::move(W) * 1.5f + 0.25f);
        }
        check(max_abs_diff(keep(X.to_host()), W) < 1e-5f, "63-step program split over several passes");
    }
    // (8) data() hands out real bytes and pins the storage to eager mode; rvalue reuse keeps the pointer
    {
        CM X = exp(dA / 10.0f);
        const void* p = (const void*)X.data();
        Matrix<float> viaptr("h", n, n);
        cudaMemcpy((void*)viaptr.data(), p, n * n * sizeof(float), cudaMemcpyDeviceToHost);
        check(max_abs_diff(viaptr, exp(A / 10.0f)) < 1e-5f, "data() points at materialised values");
        CM Y = square(std::move(X));
        check((const void*)Y.data() == p, "rvalue overload reuses the buffer after data()");
        keep(Y.to_host());
    }
    // (9) stacking and slicing deferred operands; sums of deferred values
    {
        CM H = hstack({exp(dA / 10.0f), dB + 1.0f});
        Matrix<float> want = hstack<float>({exp(A / 10.0f), B + 1.0f});
        check(max_abs_diff(keep(H.to_host()), want) < 1e-5f, "hstack of deferred operands");
        CM S = sum(exp(dA / 10.0f), 0);
        Matrix<float> ws = sum(exp(A / 10.0f), 0);
        check(rel_fro(keep(S.to_host()), ws) < 1e-5f, "sum of a deferred operand");
        CM sl = (dA * dB).slice(3, 40, 5, 17);
        check(rel_fro(keep(sl.to_host()), (A * B).slice(3, 40, 5, 17)) < 1e-5f, "slice of a deferred product");
    }
    // (10) self-referential updates
    {
        CM X(A);
        X = X * X.T() / (float)n;
        X = exp(X / 50.0f) - X;
        Matrix<float> W = A * A.T() / (float)n;
        W = exp(W / 50.0f) - W;
        check(rel_fro(keep(X.to_host()), W) < 1e-5f, "X = f(X) with X on both sides");
    }

    // (11) deferred fills: the constructor's zero-init contract, ones(), partial overwrite, and no launch at all
    //      for matrices that die (or are replaced) unread
    {
        const char* eager = std::getenv("JZ_EAGER");
        const bool lazy = !(eager && *eager && std::string(eager) != "0");
        const unsigned long long before = jz_launch_count();
        {
            CM W("weights", 64, 48), b("bias", 64, 1), v("val", 64, 32);
            W = CM::randn(64, 48) * .001;   // the Layer<D> constructor idiom (ml/layer.hpp:67-76)
            b = CM::randn(64, 1) * .001;
            v.zeros();
        }
        const unsigned long long dead = jz_launch_count() - before;
        std::cout << "    launches for constructed-and-dropped matrices: " << dead << std::endl;
        if (lazy) check(dead == 0, "matrices that die unread cost no launch");
        CM Z("z", 37, 5);
        Matrix<float> z = keep(Z.to_host());
        check(max_abs_diff(z, Matrix<float>("z", 37, 5)) == 0.0f, "public constructor is observably zero-filled");
        CM O = CM::ones(9, 4);
        O.slice(2, 5, 1, 3, CM(Matrix<float>::randn(3, 2) * 0.0f + 7.0f));   // partial overwrite of a deferred fill
        Matrix<float> o = keep(O.to_host());
        bool ok = true;
        for (size_t j = 0; j < 4; j++)
            for (size_t i = 0; i < 9; i++) ok &= o.elem(i, j) == ((i >= 2 && i < 5 && j >= 1 && j < 3) ? 7.0f : 1.0f);
        check(ok, "window assignment into a deferred ones()");
        CM F2("f", 6, 6);
        F2.ones();
        F2 += 2.0f;            // in-place step on a deferred fill
        CM G = exp(F2 * 0.0f) + F2;   // deferred readers of a deferred fill
        F2.zeros();            // redefinition after readers were taken
        check(max_abs_diff(keep(G.to_host()), Matrix<float>::ones(6, 6) * 4.0f) == 0.0f, "readers of a deferred fill see it");
        check(max_abs_diff(keep(F2.to_host()), Matrix<float>::zeros(6, 6)) == 0.0f, "refill after readers");
        CM bias = CM::ones(5, 1) * 0.5f, row = CM::ones(1, 7);
        check(max_abs_diff(keep((bias * row).to_host()), Matrix<float>::ones(5, 7) * 0.5f) == 0.0f, "b * ones(1, N) broadcast");
    }
    // (12) deferred draws: values are fixed at creation (counter-based stream), whatever runs or dies in between
    {
        GPUSampler s1(11);
        CM R1 = CM::randn(33, 17);
        { CM dropped = CM::randn(100, 100); }
        CM R2 = CM::rand(8, 8);
        Matrix<float> r2a = R2.to_host(), r1a = R1.to_host();   // materialised in the opposite order
        GPUSampler s2(11);
        CM Q1 = CM::randn(33, 17);
        Matrix<float> q1 = Q1.to_host();
        CM Qd = CM::randn(100, 100);
        Matrix<float> qd = Qd.to_host();
        CM Q2 = CM::rand(8, 8);
        Matrix<float> q2 = Q2.to_host();
        check(max_abs_diff(r1a, q1) == 0.0f && max_abs_diff(r2a, q2) == 0.0f, "draws do not depend on evaluation order");
        double mean = 0, var = 0;
        for (size_t j = 0; j < 100; j++) for (size_t i = 0; i < 100; i++) mean += qd.elem(i, j);
        mean /= 1e4;
        for (size_t j = 0; j < 100; j++) for (size_t i = 0; i < 100; i++) var += (qd.elem(i, j) - mean) * (qd.elem(i, j) - mean);
        var /= 1e4;
        check(std::fabs(mean) < 0.05 && std::fabs(var - 1.0) < 0.06, "randn is standard normal");
        keep(r1a); keep(r2a);
    }

    // (13) the broadcast idioms (rank-1 product against ones + full add) as one pass, bit-identical to the CPU path
    {
        const char* eager = std::getenv("JZ_EAGER");
        const bool lazy = !(eager && *eager && std::string(eager) != "0");
        auto W = Matrix<float>::randn(64, 48), X = Matrix<float>::randn(48, 32), b = Matrix<float>::randn(64, 1);
        auto mx = Matrix<float>::randn(1, 32), v = Matrix<float>::randn(1, 32);
        CM dW(W), dX(X), db(b), dmx(mx), dv(v);
        jz_sync(nullptr);
        unsigned long long before = jz_launch_count();
        CM H = tanh(dW * dX + db * CM::ones(1, 32));                 // Layer<D>::eval, ml/layer.hpp:118-121
        Matrix<float> h = keep(H.to_host());
        unsigned long long launches = jz_launch_count() - before;
        std::cout << "    tanh(W*x + b*ones) launches: " << launches << std::endl;
        if (lazy) check(launches <= 3, "layer forward: GEMM + broadcast add + tanh (<= 3 launches)");
        Matrix<float> wantH = tanh(W * X + b * Matrix<float>::ones(1, 32));
        std::cout << "    max |diff| vs CPU: " << max_abs_diff(h, wantH) << std::endl;
        check(max_abs_diff(h, wantH) < 2e-5f, "tanh(W*x + b*ones(1,N)) matches the CPU path (different fp32 summation order)");
        {   // and bit-identical to the unfused order on the same device product
            CM P0 = dW * dX;
            Matrix<float> p0 = P0.to_host();
            check(max_abs_diff(h, tanh(p0 + b * Matrix<float>::ones(1, 32))) < 5e-7f, "broadcast add equals product + b*ones on the device product");
        }
        CM P = dW * dX;
        Matrix<float> Ph = P.to_host();
        CM one_k1("oneK1", 64, 1);
        one_k1.ones();
        before = jz_launch_count();
        CM S1 = P - one_k1 * dmx;                                     // LogisticLayer::grad, ml/layer.hpp:260
        Matrix<float> s1h = keep(S1.to_host());
        launches = jz_launch_count() - before;
        std::cout << "    X - ones*mx launches: " << launches << std::endl;
        if (lazy) check(launches <= 2, "column-max subtraction: ones fill (first use) + one broadcast pass");
        check(max_abs_diff(s1h, Ph - Matrix<float>::ones(64, 1) * mx) == 0.0f, "X - oneK1*mx is bit-identical to the CPU path");
        CM S2 = P - one_k1 * dmx;                                     // second use: oneK1 is real bytes now, still known to be ones
        check(max_abs_diff(keep(S2.to_host()), s1h) == 0.0f, "same with a materialised ones vector");
        one_k1 += 1.0f;                                               // no longer ones: must take the general path
        CM S3 = P - one_k1 * dmx;
        check(max_abs_diff(keep(S3.to_host()), Ph - (Matrix<float>::ones(64, 1) * 2.0f) * mx) < 1e-5f, "modified vector is not mistaken for ones");
        CM S4 = P + db * dv;                                          // genuine rank-1 update
        check(max_abs_diff(keep(S4.to_host()), Ph + b * v) < 1e-5f, "general rank-1 product still exact");
        CM Pt = (dX.T() * dW.T());                                    // (W*X)^T materialised non-transposed: 32 x 64
        CM S5 = Pt.T() + db * CM::ones(1, 32);                        // flagged operand: falls back to the transposing add
        check(max_abs_diff(keep(S5.to_host()), Ph + b * Matrix<float>::ones(1, 32)) < 1e-5f, "flagged operand takes the general path");
    }

    // (14) a timing loop that never reads its result still runs one product per iteration (tests/benchmarkCoreOps.cu:64-69)
    {
        CM out("bench_gemm", n, n);
        jz_sync(nullptr);
        const unsigned long long before = jz_launch_count();
        for (int i = 0; i < 5; i++) out = dA * dB;
        const unsigned long long launches = jz_launch_count() - before;
        std::cout << "    `out = a * b` x5 launches: " << launches << std::endl;
        check(launches >= 4, "an unread product is launched when its variable is assigned over");
        check(rel_fro(keep(out.to_host()), A * B) < 1e-5f, "and the last one is the product");
    }

    // (14b) Layer::eval / Layer::grad (ml/layer.hpp:79,120): tanh(W*x + b*ones(1,N)) is ONE kernel -- product, broadcast
    //       stage and activation in the GEMM epilogue -- for a latency-bound (small-product kernel) and a tensor-core shape
    for (size_t N : {size_t(40), size_t(512)}) {
        const size_t m = N == 40 ? 96 : 384, k = N == 40 ? 200 : 320;
        auto W = Matrix<float>::randn(m, k) * 0.1, x = Matrix<float>::randn(k, N), b = Matrix<float>::randn(m, 1);
        CM dW(W), dx(x), db(b);
        jz_sync(nullptr);
        const unsigned long long before = jz_launch_count();
        CM R = tanh(dW * dx + db * CM::ones(1, N));
        Matrix<float> got = keep(R.to_host());
        const unsigned long long launches = jz_launch_count() - before;
        Matrix<float> want = tanh(W * x + b * Matrix<float>::ones(1, N));
        std::cout << "    tanh(W*x + b*ones) N=" << N << ": launches " << launches << ", rel_fro vs CPU " << rel_fro(got, want) << std::endl;
        check(rel_fro(got, want) < 1e-5f, "tanh(W*x + b*ones(1,N)) matches the CPU path");
        const char* eager = std::getenv("JZ_EAGER");
        if (!(eager && *eager && std::string(eager) != "0")) check(launches == 1, "fused: one GEMM launch with bias + activation epilogue");
        CM G = d_tanh(dW * dx + db * CM::ones(1, N));
        db += db;   // the deferred product must run with the bias it was defined on
        Matrix<float> got2 = keep(G.to_host());
        check(rel_fro(got2, d_tanh(W * x + b * Matrix<float>::ones(1, N))) < 1e-5f, "a bias modified after the deferral does not leak into it");
    }

    // (14c) the softmax-CE head exactly as LogisticLayer::grad spells it (ml/layer.hpp:252-264): every stage is deferred and the
    //       whole expression is ONE kernel; reading an intermediate mid-way falls back to the stage-by-stage kernels
    {
        const size_t K = 10, N = 96, nb = 32;
        auto X = Matrix<float>::randn(K, N) * 3.0, Yh = Matrix<float>::zeros(K, N);
        for (size_t j = 0; j < N; j++) Yh.elem(j % K, j) = 1.0f;
        CM dX(X), dY(Yh), oneK1("oneK1", K, 1);
        oneK1.ones();
        Matrix<float> hone("one", K, 1);
        hone.ones();
        auto maxf = [] __GPU_CPU__(float* v, float* vdes, int lenv, int) {
            float m = -1e30f;
            for (int i = 0; i < lenv; i++) m = m > v[i] ? m : v[i];
            vdes[0] = m;
        };
        auto head = [&](auto& in, auto& one, auto& out) {
            auto mx = reduce(maxf, in, 0, 1);
            auto shifted = in - one * mx;
            auto E = exp(std::move(shifted));
            auto Z = one * sum(E, 0);
            return -(out - E / std::move(Z)) / (double)nb;
        };
        jz_sync(nullptr);
        const unsigned long long before = jz_launch_count();
        CM G = head(dX, oneK1, dY);
        Matrix<float> got = G.to_host();   // (not in the bitwise dump: the fused kernel sums a column in another order than the composite)
        const unsigned long long launches = jz_launch_count() - before;
        Matrix<float> want = head(X, hone, Yh);
        std::cout << "    softmax-CE head launches: " << launches << ", max abs diff vs CPU " << max_abs_diff(got, want) << std::endl;
        check(max_abs_diff(got, want) < 1e-7f, "softmax-CE head matches the CPU path");
        const char* eager = std::getenv("JZ_EAGER");
        if (!(eager && *eager && std::string(eager) != "0")) check(launches == 1, "fused: the whole head is one kernel");
        // an intermediate that is read: E is looked at before the division
        auto mx = reduce(maxf, dX, 0, 1);
        auto shifted = dX - oneK1 * mx;
        auto E = exp(std::move(shifted));
        Matrix<float> Eh = E.to_host();
        auto Z = oneK1 * sum(E, 0);
        CM S = E / std::move(Z);
        Matrix<float> hmx = reduce(maxf, X, 0, 1);
        Matrix<float> hE = exp(X - hone * hmx);
        Matrix<float> hS = hE / (hone * sum(hE, 0));
        check(max_abs_diff(Eh, hE) < 1e-6f && max_abs_diff(S.to_host(), hS) < 1e-6f && max_abs_diff(mx.to_host(), hmx) == 0.0f,
              "stages of the head read mid-way (column max, exp(shifted), softmax) match the CPU path");
        // the source changes after the head was defined: the deferred result must use the OLD values
        CM dX2(X);
        CM G2 = head(dX2, oneK1, dY);
        dX2 += dX2;
        check(max_abs_diff(G2.to_host(), want) < 1e-7f, "a source modified after the deferral does not leak into the head");
    }

    // (15) random programs over a pool of matrices: every aliasing pattern the API allows (source == destination, T()
    //      views of the destination, copies taken before a source changes, rvalue chains, products feeding chains,
    //      broadcast idioms, refills).  The dump of this section must be bit-identical between the deferred and the
    //      eager (JZ_EAGER=1) run -- tests/test_dropin_gpu.py compares them -- and finite.
    {
        std::mt19937 rng(20261017);
        const size_t d = 48;
        std::vector<CM> pool;
        for (int i = 0; i < 6; i++) pool.emplace_back(CM(Matrix<float>::randn(d, d)));
        bool finite = true;
        for (int step = 0; step < 400; step++) {
            const int op = rng() % 17, i = rng() % 6, j = rng() % 6, k = rng() % 6;
            switch (op) {
                case 0: pool[k] = tanh(pool[i]); break;
                case 1: pool[k] = tanh(std::move(pool[k]) * 0.5f + 0.1f); break;
                case 2: pool[k] = tanh(pool[i] + pool[j]); break;
                case 3: pool[k] = tanh(pool[i] - pool[j] * 0.5f); break;
                case 4: pool[k] = hadmd(pool[i], pool[j]); break;
                case 5: pool[k] = tanh(pool[i] * pool[j] / (float)d); break;
                case 6: pool[k] = tanh(pool[i].T() * pool[j] / (float)d + 0.05f); break;
                case 7: pool[k] += pool[i]; pool[k] = tanh(std::move(pool[k])); break;
                case 8: pool[k] = exp(-square(pool[i])); break;
                case 9: pool[k] = tanh(pool[i].T() + pool[j]); break;
                case 10: pool[k].zeros(); pool[k] += pool[i]; break;
                case 11: pool[k] = tanh(CM::ones(d, d) * 0.25f + pool[i] * pool[j].T() / (float)d); break;
                case 12: { CM t = pool[i]; pool[i] = tanh(std::move(pool[i]) * 1.5f); pool[k] = t - pool[i]; pool[k] = tanh(std::move(pool[k])); break; }
                case 13: pool[k] = tanh(pool[i] * pool[j] + sum(pool[j], 1) * CM::ones(1, d)); break;          // W*x + b*ones(1,N)
                case 14: { CM bcol = sum(pool[i], 1) / (float)d; CM g = d_tanh(pool[i].T() * pool[j] + bcol * CM::ones(1, d));
                           bcol += bcol; pool[k] = hadmd(g, pool[j]); break; }                                  // bias changes after the deferral
                case 15: pool[k] = tanh(pool[i] * pool[j] - CM::ones(d, 1) * sum(pool[i], 0)); break;          // per-column broadcast
                default: pool[k] = tanh(pool[i] - CM::ones(d, 1) * sum(pool[i], 0) / (float)d); break;
            }
            if (step % 20 == 19) {
                Matrix<float> h = keep(pool[rng() % 6].to_host());
                for (size_t c = 0; c < d; c++) for (size_t r = 0; r < d; r++) finite &= std::isfinite(h.elem(r, c));
            }
        }
        for (auto& m : pool) keep(m.to_host());
        check(finite, "400-step random program stays finite (dump compared bitwise against the eager run)");
    }

    const std::string path = std::string(PROJECT_DIR) + "/res/test_fusion_dump.bin";
    if (FILE* f = fopen(path.c_str(), "wb")) {
        for (const auto& m : dump) fwrite(m.data(), sizeof(float), m.num_row() * m.num_col(), f);
        fclose(f);
    }
    std::cout << (failures ? "FAILED" : "ALL PASSED") << std::endl;
    return failures ? 1 : 0;
}
