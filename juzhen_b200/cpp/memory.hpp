// memory.hpp -- block recycling for Matrix<D> (drop-in for the reference's cpp/memory.hpp:30-143).
//
// Interface kept: `Memory<D>::allocate(n)`, `Memory<D>::free(p)`, and an RAII object whose
// destructor releases everything (created once in main(), cpp/launcher.cu:66-78).
//
//  * host types (float/int/double):  an exact-size free list, like the reference
//    (cpp/memory.hpp:50-97) -- a freed block is handed back to the next request of the same
//    element count, nothing is returned to the system before the RAII object dies.
//  * CUDAfloat:  a thin front for the STREAM-ORDERED device pool inside libjz_b200.so
//    (jz_malloc / jz_free, juzhen_b200/csrc/jz_runtime.cu).  The pool keeps the reference's
//    "free then allocate the same size returns the same block" behaviour that rvalue operator
//    chains rely on, adds cross-stream ordering by events, and trims itself on OOM.
#pragma once
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <unordered_map>
#include <vector>

#include "helper.hpp"

template <class D>
class Memory {
    struct Book {
        std::unordered_map<D *, size_t> in_use;                // block -> element count
        std::unordered_map<size_t, std::vector<D *>> spare;    // element count -> idle blocks
        bool released = false;                                 // the RAII owner in main() is gone
    };
    // leaked on purpose: matrices with static storage may be released after main() returns
    static Book &book() {
        static Book *b = new Book();
        return *b;
    }

   public:
    static D *allocate(size_t count) {
        Book &b = book();
        auto hit = b.spare.find(count);
        if (hit != b.spare.end() && !hit->second.empty()) {
            D *p = hit->second.back();
            hit->second.pop_back();
            b.in_use.emplace(p, count);
            return p;
        }
        D *p = nullptr;
        try {
            p = new D[count ? count : 1];  // a zero-sized request still yields a valid block
        } catch (std::bad_alloc &e) {
            LOG_ERROR("host allocation of {} elements failed: {}", count, e.what());
            ERROR_OUT;
        }
        b.in_use.emplace(p, count);
        return p;
    }

    static void free(D *p) {
        Book &b = book();
        auto it = b.in_use.find(p);
        if (it == b.in_use.end()) {
            if (b.released) return;  // a static-duration matrix outlived main(): its block is already gone
            LOG_ERROR("Memory<{}>::free: unknown block {}", datatype(D), (void *)p);
            ERROR_OUT;
        }
        b.spare[it->second].push_back(p);
        b.in_use.erase(it);
    }

    ~Memory() {
        Book &b = book();
        size_t elems = 0;
        for (auto &kv : b.in_use) { delete[] kv.first; elems += kv.second; }
        for (auto &kv : b.spare)
            for (D *p : kv.second) { delete[] p; elems += kv.first; }
        b.in_use.clear();
        b.spare.clear();
        b.released = true;
        LOG_INFO("Total {} memory released: {:.2f} MB.", datatype(D), elems * sizeof(D) / 1024.0 / 1024.0);
    }
};

#ifdef CUDA
#include <jz_b200.h>

// The stream every Matrix<CUDAfloat> operation is issued on.  NULL = the legacy default stream,
// which is what the reference uses and what unchanged callers (ml/layer.hpp, ml/util.cuh, the
// examples' own kernels) launch on, so ordering with their raw <<<>>> launches is preserved.
inline jz_stream_t &jz_cpp_stream() {
    static jz_stream_t s = nullptr;
    return s;
}

template <>
class Memory<CUDAfloat> {
   public:
    static CUDAfloat *allocate(size_t count) {
        float *p = nullptr;
        const int rc = jz_malloc(&p, count, jz_cpp_stream());
        if (rc != JZ_OK) {
            std::fprintf(stderr, "Memory<CUDAfloat>::allocate(%zu floats) failed: %s\n", count, jz_last_error());
            ERROR_OUT;
        }
        return reinterpret_cast<CUDAfloat *>(p);
    }
    static void free(CUDAfloat *p) {
        if (jz_free(reinterpret_cast<float *>(p), jz_cpp_stream()) != JZ_OK) {
            std::fprintf(stderr, "Memory<CUDAfloat>::free failed: %s\n", jz_last_error());
            ERROR_OUT;
        }
    }
    // scope exit in main(): give every block (live and cached) back to the driver
    ~Memory() {
        size_t live = 0, cached = 0;
        jz_pool_stats(&live, &cached, nullptr, nullptr);
        LOG_INFO("Total CUDAfloat memory released: {:.2f} MB.", (live + cached) / 1024.0 / 1024.0);
        jz_shutdown();
    }
};
#endif
