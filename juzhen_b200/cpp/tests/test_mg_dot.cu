// test_mg_dot.cu -- a C++ Matrix<CUDAfloat> program sharded over N GPUs through the jz_mg_* ABI (juzhen_b200/cpp/jz_mg.hpp):
// one PROCESS per GPU (fork before any CUDA call), handles exchanged through a shared-memory page, no torch / NCCL / MPI.
//   1. C = log(exp(A*B/k)+1) column-sharded with the all-gather fused into the tcgen05 epilogue, every rank's gathered
//      image compared with the single-GPU product of the same operands (same kernel, possibly another split-K plan:
//      1e-6 relative), and bit-for-bit across ranks;
//   2. column sums of a ROW-sharded matrix: local sum(X_r, 0) + jz_mg_allreduce_sum, against the stacked matrix.
// Own main() (the launcher's main() initialises CUDA before compute() and cannot fork).  Usage: test_mg_dot [world]
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../cpp/juzhen.hpp"
#include "../cpp/jz_mg.hpp"

int compute() { return 0; }   // cpp/juzhen.hpp declares it; this program has its own main

struct SharedPage {
    std::atomic<int> arrived[64];      // one counter per exchange round
    std::atomic<int> failures;
    unsigned char bytes[64][8][JZ_MG_HANDLE_BYTES];   // [round][rank]
    unsigned checksum[8];
};

static int run_rank(int world, int rank, SharedPage* page) {
    if (jz_init(rank) != JZ_OK) { std::fprintf(stderr, "rank %d: %s\n", rank, jz_last_error()); return 1; }
    Memory<int> host_int; Memory<float> host_f32; Memory<double> host_f64; Memory<CUDAfloat> device_pool;
    int round = 0;
    Juzhen::mg::Exchange ex = [&](const void* mine, void* all, size_t bytes) {
        const int r = round++;
        std::memcpy(page->bytes[r][rank], mine, bytes);
        page->arrived[r].fetch_add(1);
        while (page->arrived[r].load() < world) usleep(50);
        for (int q = 0; q < world; q++) std::memcpy(static_cast<unsigned char*>(all) + size_t(q) * bytes, page->bytes[r][q], bytes);
    };
    int bad = 0;
    {
        Juzhen::mg::Comm comm(world, rank, ex);
        // ---- 1. sharded product + fused chain
        const size_t m = 1536, n = 2048, k = 1024;
        global_rand_gen.seed(0);   // every rank draws the same host operands (cpp/matrix.hpp:49-71)
        auto Ah = Matrix<float>::randn(m, k), Bh = Matrix<float>::randn(k, n);
        Matrix<CUDAfloat> A(Ah), B(Bh);
        size_t j0, j1;
        comm.block_range(n, j0, j1);
        Matrix<CUDAfloat> Bblk = B.columns(j0, j1);
        const jz_step steps[4] = {{JZ_STEP_AFFINE, 1.0f / k, 0.0f}, {JZ_EXP, 0, 0}, {JZ_STEP_AFFINE, 1.0f, 1.0f}, {JZ_LOG, 0, 0}};
        Juzhen::mg::Replicated C(comm, m, n);
        for (int rep = 0; rep < 3; rep++) C.dot_allgather(A, Bblk, steps, 4);
        Matrix<CUDAfloat> got = C.matrix();
        Matrix<CUDAfloat> want = log(exp(A * B / (double)k) + 1.0);
        const float err = (got - want).norm() / want.norm();
        auto gh = got.to_host();
        unsigned cs = 0;
        for (size_t i = 0; i < m * n; i++) { unsigned u; std::memcpy(&u, gh.data() + i, 4); cs = cs * 31u + u; }
        page->checksum[rank] = cs;
        std::printf("rank %d/%d: sharded dot + chain %zux%zux%zu, columns [%zu, %zu): rel err vs single-GPU product %.2e, checksum %08x\n", rank,
                    world, m, n, k, j0, j1, err, cs);
        bad += !(err < 1e-6f);
        // ---- 2. column sums over row-sharded data
        const size_t rows = 3000, cols = 517;
        global_rand_gen.seed(100 + rank);
        auto Xh = Matrix<float>::randn(rows, cols);
        Matrix<CUDAfloat> X(Xh);
        Matrix<CUDAfloat> total = Juzhen::mg::allreduce_sum(comm, sum(X, 0));
        auto th = total.to_host();
        double worst = 0;
        std::vector<double> ref(cols, 0.0);
        for (int r = 0; r < world; r++) {
            global_rand_gen.seed(100 + r);
            auto Xr = Matrix<float>::randn(rows, cols);
            for (size_t j = 0; j < cols; j++)
                for (size_t i = 0; i < rows; i++) ref[j] += Xr.elem(i, j);
        }
        for (size_t j = 0; j < cols; j++) worst = std::max(worst, std::fabs(th.elem(0, j) - ref[j]));   // sum(X, 0) is a logical 1 x cols row
        std::printf("rank %d/%d: row-sharded column sums + allreduce: max abs err %.2e\n", rank, world, worst);
        bad += !(worst < 1e-5 * std::sqrt(double(rows * world)) * 4);
        comm.barrier();
        jz_sync(nullptr);
    }
    return bad;
}

int main(int argc, char** argv) {
    setvbuf(stdout, nullptr, _IOLBF, 0);
    const int world = argc > 1 ? std::atoi(argv[1]) : 2;
    if (world < 1 || world > 8) { std::fprintf(stderr, "world must be 1..8\n"); return 2; }
    auto* page = static_cast<SharedPage*>(mmap(nullptr, sizeof(SharedPage), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0));
    if (page == MAP_FAILED) { std::perror("mmap"); return 2; }
    std::memset(page, 0, sizeof(SharedPage));
    std::vector<pid_t> kids;
    for (int r = 0; r < world; r++) {
        pid_t p = fork();
        if (p == 0) {
            const int rc = run_rank(world, r, page);
            std::fflush(stdout);   // _exit skips the stdio flush
            std::fflush(stderr);
            _exit(rc ? 1 : 0);
        }
        kids.push_back(p);
    }
    int failed = 0;
    for (size_t r = 0; r < kids.size(); r++) {
        int st = 0;
        waitpid(kids[r], &st, 0);
        const bool ok = WIFEXITED(st) && WEXITSTATUS(st) == 0;
        if (!ok) {
            if (WIFSIGNALED(st)) std::printf("rank %zu: killed by signal %d\n", r, WTERMSIG(st));
            else std::printf("rank %zu: exit status %d\n", r, WEXITSTATUS(st));
        }
        failed += !ok;
    }
    bool same = true;
    for (int r = 1; r < world; r++) same &= page->checksum[r] == page->checksum[0];
    std::printf("gathered images bit-identical across ranks: %s\n", same ? "yes" : "NO");
    std::printf(failed || !same ? "MG_CPP_FAIL\n" : "MG_CPP_OK\n");
    return failed || !same ? 1 : 0;
}
