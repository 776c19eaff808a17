// jz_mg.hpp -- multi-GPU helpers for C++ programs written against Matrix<CUDAfloat> (one process per GPU): thin RAII
// wrappers over the jz_mg_* entry points of include/jz_b200.h.  The reference has nothing of the kind (SURVEY section
// 2: no collectives); this is the C++ face of SURVEY 8(e):
//
//     Juzhen::mg::Comm comm(world, rank, exchange);      // exchange: all-gather of opaque bytes between the processes
//     Juzhen::mg::Replicated C(comm, m, n);              // every rank holds a full m x n image, peers' images mapped
//     C.dot_allgather(A, B_block, steps, nsteps);        // this rank's column block of chain(A * B), stored into
//                                                        // every image by the GEMM epilogue, then a barrier
//     Matrix<CUDAfloat> full = C.matrix();               // the gathered result, as an ordinary matrix (a copy)
//
// `exchange(mine, all, bytes)` must fill all[r*bytes .. (r+1)*bytes) with rank r's `mine` on every rank (MPI_Allgather,
// a pipe fan-out, shared memory ...): the library never talks to other processes itself.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#include <jz_b200.h>

namespace Juzhen {
namespace mg {

using Exchange = std::function<void(const void* mine, void* all, size_t bytes)>;

inline void check(int rc, const char* what) {
    if (rc != JZ_OK) {
        std::fprintf(stderr, "jz_mg: %s failed: %s\n", what, jz_last_error());
        std::exit(1);   // device errors are fatal, as everywhere in the reference (cpp/cumatrix.cuh:34-55)
    }
}

// a device buffer every rank allocates with the same size, with every peer's copy mapped into this process
class Shared {
   public:
    Shared(int world, int rank, const Exchange& ex, size_t count) : world_(world), rank_(rank), ptrs_(world, nullptr) {
        check(jz_malloc(&ptrs_[rank], count, nullptr), "jz_malloc");
        check(jz_fill(ptrs_[rank], count, 0.0f, nullptr), "jz_fill");
        check(jz_sync(nullptr), "jz_sync");
        std::vector<unsigned char> mine(JZ_MG_HANDLE_BYTES), all(size_t(JZ_MG_HANDLE_BYTES) * world);
        check(jz_mg_export(ptrs_[rank], mine.data()), "jz_mg_export");
        ex(mine.data(), all.data(), JZ_MG_HANDLE_BYTES);
        for (int r = 0; r < world; r++)
            if (r != rank) check(jz_mg_import(all.data() + size_t(r) * JZ_MG_HANDLE_BYTES, &ptrs_[r]), "jz_mg_import");
    }
    ~Shared() {
        jz_sync(nullptr);
        for (int r = 0; r < world_; r++)
            if (r != rank_ && ptrs_[r]) jz_mg_release(ptrs_[r]);
        if (ptrs_[rank_]) jz_free(ptrs_[rank_], nullptr);
    }
    Shared(const Shared&) = delete;
    Shared& operator=(const Shared&) = delete;
    float* local() const { return ptrs_[rank_]; }
    float* const* images() const { return ptrs_.data(); }

   private:
    int world_, rank_;
    std::vector<float*> ptrs_;
};

class Comm {
   public:
    Comm(int world, int rank, Exchange ex) : world_(world), rank_(rank), ex_(std::move(ex)), flags_(world, rank, ex_, 64), epoch_(0) {}
    int world() const { return world_; }
    int rank() const { return rank_; }
    const Exchange& exchange() const { return ex_; }
    // device-side barrier on the legacy default stream (the stream Matrix<CUDAfloat> works on)
    void barrier() {
        check(jz_mg_barrier(reinterpret_cast<unsigned* const*>(flags_.images()), world_, rank_, ++epoch_, nullptr), "jz_mg_barrier");
    }
    void block_range(size_t n, size_t& begin, size_t& end) const { check(jz_mg_block_range(n, world_, rank_, &begin, &end), "jz_mg_block_range"); }

   private:
    int world_, rank_;
    Exchange ex_;
    Shared flags_;   // 64 words, zero-filled: the flag array of jz_mg_barrier
    unsigned epoch_;
};

// a dense m x n matrix replicated on every rank, filled by column-sharded products
class Replicated {
   public:
    Replicated(Comm& comm, size_t m, size_t n) : comm_(comm), m_(m), n_(n), buf_(comm.world(), comm.rank(), comm.exchange(), m * n) {}
    // C = chain(alpha * op(A) * B) with B given as this rank's column block (k x (j1 - j0), see Comm::block_range)
    void dot_allgather(const Matrix<CUDAfloat>& A, const Matrix<CUDAfloat>& B_block, const jz_step* steps = nullptr, int nsteps = 0,
                       float alpha = 1.0f, int mode = -1) {
        size_t j0, j1;
        comm_.block_range(n_, j0, j1);
        if (A.num_row() != m_ || A.num_col() != B_block.num_row() || B_block.num_col() != j1 - j0 || B_block.get_transpose())
            throw std::invalid_argument("Matrix dimensions are not compatible");
        comm_.barrier();   // nobody is still reading the previous contents of any image
        check(jz_mg_gemm_allgather(A.get_transpose(), m_, n_, A.num_col(), alpha, reinterpret_cast<const float*>(A.data()),
                                   A.get_transpose() ? A.num_col() : A.num_row(), reinterpret_cast<const float*>(B_block.data()),
                                   B_block.num_row(), buf_.images(), comm_.world(), comm_.rank(), steps, nsteps, mode, nullptr),
              "jz_mg_gemm_allgather");
        comm_.barrier();   // every rank's columns have landed in every image
    }
    Matrix<CUDAfloat> matrix() const {
        Matrix<CUDAfloat> r("gathered", m_, n_);
        check(jz_memcpy_d2d(const_cast<float*>(reinterpret_cast<const float*>(r.data())), buf_.local(), m_ * n_, nullptr), "jz_memcpy_d2d");
        return r;
    }
    float* data() const { return buf_.local(); }

   private:
    Comm& comm_;
    size_t m_, n_;
    Shared buf_;
};

// sum over ranks of a per-rank vector (e.g. sum(X_r, 0) of a row-sharded X): bitwise identical on every rank
inline Matrix<CUDAfloat> allreduce_sum(Comm& comm, const Matrix<CUDAfloat>& partial) {
    const size_t n = partial.num_row() * partial.num_col();
    Shared img(comm.world(), comm.rank(), comm.exchange(), n);
    check(jz_memcpy_d2d(img.local(), reinterpret_cast<const float*>(partial.data()), n, nullptr), "jz_memcpy_d2d");
    comm.barrier();
    Matrix<CUDAfloat> out("allreduce", partial.get_transpose() ? partial.num_col() : partial.num_row(),
                          partial.get_transpose() ? partial.num_row() : partial.num_col());
    check(jz_mg_allreduce_sum(const_cast<float*>(reinterpret_cast<const float*>(out.data())), img.images(), n, comm.world(), comm.rank(), nullptr),
          "jz_mg_allreduce_sum");
    comm.barrier();   // img may be unmapped only after every rank has read it
    return partial.get_transpose() ? Matrix<CUDAfloat>(out.T()) : out;
}

}  // namespace mg
}  // namespace Juzhen
