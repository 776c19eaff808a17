// jz_rng.cu -- counter-based Philox4x32-10 generator behind Matrix<CUDAfloat>::rand/randn
// (SURVEY 8a row a19).  Replaces cuRAND XORWOW incl. its odd-count scratch cudaMalloc +
// D2D copy (cpp/cumatrix.cu:358-420): any count, no scratch, 4 B/elem written.
// The reference's GPU stream is not reproducible against its CPU mt19937 stream either, so
// RNG is not parity-pinned; tests check moments and determinism per (seed, offset).
#include "jz_common.cuh"

namespace jz {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__device__ __forceinline__ void philox4x32_10(uint64_t index, uint64_t seed, uint64_t offset, uint32_t (&out)[4]) {
    uint32_t c[4] = {uint32_t(index), uint32_t(index >> 32), uint32_t(offset), uint32_t(offset >> 32)};
    uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

template <bool NORMAL>
__global__ void __launch_bounds__(256) rng_kernel(float* x, size_t n, uint64_t seed, uint64_t offset) {
    pdl_enter();
    const size_t n4 = (n + 3) >> 2;
    for (size_t q = size_t(blockIdx.x) * 256 + threadIdx.x; q < n4; q += size_t(gridDim.x) * 256) {
        uint32_t r[4];
        philox4x32_10(q, seed, offset, r);
        float v[4];
        if (NORMAL) {
#pragma unroll
            for (int p = 0; p < 2; p++) {
                const float u1 = float((r[2 * p] >> 8) + 1u) * 5.9604644775390625e-8f;   // (0, 1]
                const float u2 = float(r[2 * p + 1] >> 8) * 5.9604644775390625e-8f;      // [0, 1)
                const float rad = sqrtf(-2.0f * logf(u1));
                float sn, cs;
                sincospif(2.0f * u2, &sn, &cs);
                v[2 * p] = rad * cs;
                v[2 * p + 1] = rad * sn;
            }
        } else {
#pragma unroll
            for (int p = 0; p < 4; p++) v[p] = float(r[p] >> 8) * 5.9604644775390625e-8f;    // [0, 1)
        }
        const size_t base = q << 2;
        if (base + 3 < n && aligned16(x)) {
            *reinterpret_cast<float4*>(x + base) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int p = 0; p < 4; p++)
                if (base + p < n) x[base + p] = v[p];
        }
    }
}

template <bool NORMAL>
static int launch_rng(float* x, size_t n, uint64_t seed, uint64_t offset, cudaStream_t s) {
    if (n == 0) return JZ_OK;
    if (!x) return fail(JZ_ERR_ARG, "rng: null pointer");
    const size_t cap = size_t(ctx().sm_count) * 8;
    const size_t blocks = ceil_div((n + 3) >> 2, size_t(256));
    JZ_LAUNCH((rng_kernel<NORMAL>), unsigned(blocks < cap ? blocks : cap), 256, 0, s, x, n, seed, offset);
    return JZ_OK;
}

}  // namespace jz

using namespace jz;

extern "C" {

int jz_rand_uniform(float* x, size_t n, uint64_t seed, uint64_t offset, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    return launch_rng<false>(x, n, seed, offset, as_stream(stream));
}

int jz_rand_normal(float* x, size_t n, uint64_t seed, uint64_t offset, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    return launch_rng<true>(x, n, seed, offset, as_stream(stream));
}

}  // extern "C"
