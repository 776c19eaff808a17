// jz_math.cuh -- per-element device math shared by the map kernels, the fused chain,
// the GEMM epilogue and the accuracy sweep.
//
// Parity target (BASELINE.json north_star): <= 2 ulp against the reference's g++ CPU build,
// which evaluates exp/log/tanh in DOUBLE libm and rounds once, i.e. the
// comparator is the correctly rounded fp32 value.  Arithmetic ops are written with explicit
// round-to-nearest intrinsics so nvcc cannot contract them into FMAs: the x86-64 reference
// has no FMA, and __fmul_rn/__fadd_rn make affine/axpby/hadamard BIT-EXACT with it.
#pragma once
#include <cuda_runtime.h>
#include "../../include/jz_b200.h"

namespace jz {

// s1*x + a, two roundings (cpp/core.hpp:452-469)
__device__ __forceinline__ float affine_rn(float x, float s1, float a) {
    return __fadd_rn(__fmul_rn(s1, x), a);
}

// l / x : the reference divides in double and rounds to float (cpp/core.hpp:471-482).
// For fp32 operands double rounding of a quotient is innocuous (53 >= 2*24+2), so the
// IEEE fp32 division gives the same bits.
__device__ __forceinline__ float eleminv_rn(float x, float l) { return __fdiv_rn(l, x); }

// exp(t) for t in [-88, 0] with < 1 ulp error (CUDA's expf is documented at 2 ulp, which alone would
// use up the whole d_tanh budget).  t = n*ln2 + r with |r| <= ln2/2: n from the round-to-nearest
// magic-number trick, r by a two-term Cody-Waite reduction (the first FMA is exact), exp(r) =
// 1 + r + r^2*P5(r) (fit error 0.012 ulp), and the 2^n scaling is an integer add on the exponent.
__device__ __forceinline__ float exp_neg_1ulp(float t) {
    const float z = __fmaf_rn(t, 1.4426950408889634f, 12582912.0f);
    const float n = __fsub_rn(z, 12582912.0f);
    float r = __fmaf_rn(n, -0.693145751953125f, t);
    r = __fmaf_rn(n, -1.4286068203094173e-06f, r);
    float p = 0.00019899278413504362f;
    p = __fmaf_rn(p, r, 0.0013933645095676184f);
    p = __fmaf_rn(p, r, 0.0083332983776927f);
    p = __fmaf_rn(p, r, 0.04166646674275398f);
    p = __fmaf_rn(p, r, 0.1666666716337204f);
    p = __fmaf_rn(p, r, 0.5f);
    p = __fmaf_rn(p, __fmul_rn(r, r), r);
    p = __fadd_rn(1.0f, p);
    return __int_as_float(__float_as_int(p) + (__float_as_int(z) << 23));
}

// d_tanh(x) = 1 - tanh(x)^2 = 4e / (1+e)^2 with e = exp(-2|x|).
// The reference evaluates 1 - tanh^2 in double (cpp/matrix.hpp:301-324); in fp32 that form
// cancels catastrophically for |x| >~ 5, so the quotient form is used with a compensated
// denominator: s = fl(1+e) with exact residual serr (Fast2Sum, 1 >= e), t = fl(s*s) with exact
// residual terr (FMA), an FMA-refined quotient q = 4e/t and one correction step for t's residual.
__device__ __forceinline__ float dtanh_acc(float x) {
    const float ax = fabsf(x);
    if (!(ax <= 40.0f)) {
        // exp(-2|x|) approaches the subnormal range (its fp32 quantisation alone would cost > 2 ulp),
        // or x is inf / NaN.  Never taken for activations in practice; do it in double.
        return float(4.0 * exp(-2.0 * double(ax)));   // (1+e)^2 == 1 to 1e-34 here
    }
    const float e = exp_neg_1ulp(-2.0f * ax);
    const float s = __fadd_rn(1.0f, e);
    const float serr = __fsub_rn(e, __fsub_rn(s, 1.0f));
    const float t = __fmul_rn(s, s);
    const float terr = __fmaf_rn(s, s, -t);
    const float c = __fmaf_rn(2.0f * s, serr, terr);  // (1+e)^2 = t + c (+ serr^2 ~ 2^-48)
    const float num = 4.0f * e;
    float rc;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(t));  // t in [1, 4]
    const float q0 = __fmul_rn(num, rc);
    const float q = __fmaf_rn(__fmaf_rn(-q0, t, num), rc, q0);  // residual-corrected quotient
    return __fmaf_rn(-q, __fmul_rn(c, rc), q);
}

// The same computation on two elements per instruction (packed fp32 pipe forms, see log_main2 below): every lane goes
// through the same sequence of individually rounded operations as dtanh_acc, so the bits are the same; 25 of its ~30
// instructions are floating-point and pack.  Both lanes must be in the main range (|x| <= 40).
__device__ __forceinline__ float2 jz_splat2(float c) { return make_float2(c, c); }
__device__ __forceinline__ float2 jz_neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 exp_neg_1ulp2(float2 t) {
    const float2 z = __ffma2_rn(t, jz_splat2(1.4426950408889634f), jz_splat2(12582912.0f));
    const float2 n = __fadd2_rn(z, jz_splat2(-12582912.0f));
    float2 r = __ffma2_rn(n, jz_splat2(-0.693145751953125f), t);
    r = __ffma2_rn(n, jz_splat2(-1.4286068203094173e-06f), r);
    float2 p = jz_splat2(0.00019899278413504362f);
    p = __ffma2_rn(p, r, jz_splat2(0.0013933645095676184f));
    p = __ffma2_rn(p, r, jz_splat2(0.0083332983776927f));
    p = __ffma2_rn(p, r, jz_splat2(0.04166646674275398f));
    p = __ffma2_rn(p, r, jz_splat2(0.1666666716337204f));
    p = __ffma2_rn(p, r, jz_splat2(0.5f));
    p = __ffma2_rn(p, __fmul2_rn(r, r), r);
    p = __fadd2_rn(jz_splat2(1.0f), p);
    return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(z.x) << 23)),
                       __int_as_float(__float_as_int(p.y) + (__float_as_int(z.y) << 23)));
}
__device__ __forceinline__ float2 dtanh_acc2(float2 x) {   // |x.x|, |x.y| <= 40
    const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
    const float2 e = exp_neg_1ulp2(make_float2(-2.0f * ax.x, -2.0f * ax.y));
    const float2 s = __fadd2_rn(jz_splat2(1.0f), e);
    const float2 serr = __fadd2_rn(e, jz_neg2(__fadd2_rn(s, jz_splat2(-1.0f))));
    const float2 t = __fmul2_rn(s, s);
    const float2 terr = __ffma2_rn(s, s, jz_neg2(t));
    const float2 c = __ffma2_rn(make_float2(2.0f * s.x, 2.0f * s.y), serr, terr);
    const float2 num = make_float2(4.0f * e.x, 4.0f * e.y);
    float2 rc;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc.x) : "f"(t.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc.y) : "f"(t.y));
    const float2 q0 = __fmul2_rn(num, rc);
    const float2 q = __ffma2_rn(__ffma2_rn(jz_neg2(q0), t, num), rc, q0);
    return __ffma2_rn(jz_neg2(q), __fmul2_rn(c, rc), q);
}
template <int N>
__device__ __forceinline__ void dtanh_tile(float (&v)[N]) {
    bool main_range = (N % 2) == 0;
#pragma unroll
    for (int i = 0; i < N; i++) main_range &= fabsf(v[i]) <= 40.0f;   // false for NaN too
    if (main_range) {
#pragma unroll
        for (int i = 0; i + 1 < N; i += 2) {
            const float2 r = dtanh_acc2(make_float2(v[i], v[i + 1]));
            v[i] = r.x;
            v[i + 1] = r.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; i++) v[i] = dtanh_acc(v[i]);
    }
}

// log(x) in ~20 instructions on the main path (CUDA's logf is ~28 and makes the fused softplus chain
// ALU-bound): x = m * 2^e with m in [2/3, 4/3), f = m - 1 (exact),
// log1p(f) = f + f^2 * (-1/2 + f * Q7(f)), result = fma(e, ln2, log1p(f)).  Measured worst case 0.92 ulp from the
// true value over every mantissa at ten exponents (and by the exhaustive device sweep in the tests).
// Zero, negative, denormal, inf and NaN inputs take the library path.
__device__ __forceinline__ bool log_needs_library(float x) {
    return unsigned(__float_as_int(x) - 0x00800000) >= 0x7f000000u;
}
__device__ __forceinline__ float log_main(float x) {  // positive normal x only
    const int ix = __float_as_int(x);
    const int e = (ix - 0x3f2aaaab) >> 23;
    const float f = __fsub_rn(__int_as_float(ix - (e << 23)), 1.0f);
    float q = -0.128597691655159f;
    q = __fmaf_rn(q, f, 0.1401381939649582f);
    q = __fmaf_rn(q, f, -0.12192238867282867f);
    q = __fmaf_rn(q, f, 0.13999086618423462f);
    q = __fmaf_rn(q, f, -0.16679848730564117f);
    q = __fmaf_rn(q, f, 0.20010896027088165f);
    q = __fmaf_rn(q, f, -0.24999813735485077f);
    q = __fmaf_rn(q, f, 0.3333320617675781f);
    const float w = __fmaf_rn(f, q, -0.5f);
    const float r = __fmaf_rn(__fmul_rn(f, f), w, f);
    return __fmaf_rn(float(e), 0.6931471805599453f, r);
}
__device__ __forceinline__ float log_1ulp(float x) { return log_needs_library(x) ? logf(x) : log_main(x); }
// Two elements per instruction: sm_100's packed fp32 pipe forms (FFMA2 / FADD2 / FMUL2 = fma / add / mul .rn.f32x2, each
// lane rounded exactly like the scalar instruction, so the result has the same bits as log_main on either lane).  The
// fused chain is ISSUE-bound (~38 instructions per element against 8 bytes, ncu: sm 84 %, dram 61-68 %); the thirteen
// floating-point instructions of the polynomial become 6.5 per element.  Only fma-into-fma sequences are packed:
// ptxas contracts a mul.rn.f32x2 followed by an add.rn.f32x2 into one FFMA2 (seen in SASS), which would break the
// separately rounded `s1*x + a` of the affine steps -- those stay scalar (__fmul_rn / __fadd_rn are never contracted).
__device__ __forceinline__ float2 splat2(float c) { return make_float2(c, c); }
__device__ __forceinline__ float2 log_main2(float2 x) {  // positive normal x only, both lanes
    const int ix0 = __float_as_int(x.x), ix1 = __float_as_int(x.y);
    const int e0 = (ix0 - 0x3f2aaaab) >> 23, e1 = (ix1 - 0x3f2aaaab) >> 23;
    const float2 f = __fadd2_rn(make_float2(__int_as_float(ix0 - (e0 << 23)), __int_as_float(ix1 - (e1 << 23))), splat2(-1.0f));
    float2 q = splat2(-0.128597691655159f);
    q = __ffma2_rn(q, f, splat2(0.1401381939649582f));
    q = __ffma2_rn(q, f, splat2(-0.12192238867282867f));
    q = __ffma2_rn(q, f, splat2(0.13999086618423462f));
    q = __ffma2_rn(q, f, splat2(-0.16679848730564117f));
    q = __ffma2_rn(q, f, splat2(0.20010896027088165f));
    q = __ffma2_rn(q, f, splat2(-0.24999813735485077f));
    q = __ffma2_rn(q, f, splat2(0.3333320617675781f));
    const float2 w = __ffma2_rn(f, q, splat2(-0.5f));
    const float2 r = __ffma2_rn(__fmul2_rn(f, f), w, f);
    return __ffma2_rn(make_float2(float(e0), float(e1)), splat2(0.6931471805599453f), r);
}
// register-tile form: ONE branch per tile instead of one per element (the per-element form costs ~8 extra
// instructions per element in branch bookkeeping, measured in the fused chain)
template <int N>
__device__ __forceinline__ void log_tile(float (&v)[N]) {
    bool special = false;
#pragma unroll
    for (int i = 0; i < N; i++) special |= log_needs_library(v[i]);
    if (!special) {
        if constexpr (N % 2 == 0) {
#pragma unroll
            for (int i = 0; i < N; i += 2) {
                const float2 r = log_main2(make_float2(v[i], v[i + 1]));
                v[i] = r.x;
                v[i + 1] = r.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < N; i++) v[i] = log_main(v[i]);
        }
    } else {   // rare: decide per element, so that the result of an element never depends on its tile mates
#pragma unroll  // (a rolled loop would index v[] dynamically and push the whole tile into local memory)
        for (int i = 0; i < N; i++) v[i] = log_1ulp(v[i]);
    }
}

template <int OP>
__device__ __forceinline__ float unary_op(float x) {
    if constexpr (OP == JZ_EXP) return expf(x);
    else if constexpr (OP == JZ_LOG) return log_1ulp(x);
    else if constexpr (OP == JZ_TANH) return tanhf(x);
    else if constexpr (OP == JZ_DTANH) return dtanh_acc(x);
    else if constexpr (OP == JZ_SQUARE) return __fmul_rn(x, x);
    else if constexpr (OP == JZ_SQRT) return __fsqrt_rn(x);
    else if constexpr (OP == JZ_RELU) return x > 0.0f ? x : 0.0f;
    else if constexpr (OP == JZ_DRELU) return x > 0.0f ? 1.0f : 0.0f;
    else return x;
}

// runtime-dispatched step over a small register tile (switch amortised over N values)
template <int N>
__device__ __forceinline__ void apply_step(float (&v)[N], int kind, float s1, float a) {
    switch (kind) {
#define JZ_CASE(OP)                                         \
    case OP:                                                \
        _Pragma("unroll") for (int i = 0; i < N; i++) v[i] = unary_op<OP>(v[i]); \
        break;
        JZ_CASE(JZ_EXP)
        case JZ_LOG:
            if constexpr (N <= 16) {
                log_tile<N>(v);
            } else {  // GEMM epilogue tiles: register pressure matters more than issue slots there
#pragma unroll
                for (int i = 0; i < N; i++) v[i] = log_1ulp(v[i]);
            }
            break;
        JZ_CASE(JZ_TANH)
        case JZ_DTANH:
            if constexpr (N <= 16) {
                dtanh_tile<N>(v);
            } else {
#pragma unroll
                for (int i = 0; i < N; i++) v[i] = dtanh_acc(v[i]);
            }
            break;
        JZ_CASE(JZ_SQUARE)
        JZ_CASE(JZ_SQRT)
        JZ_CASE(JZ_RELU)
        JZ_CASE(JZ_DRELU)
#undef JZ_CASE
        case JZ_STEP_AFFINE:
            if (s1 == 1.0f && N % 2 == 0) {   // x + a: 1 * x is exact, so only the addition rounds -- two lanes per FADD2
#pragma unroll
                for (int i = 0; i + 1 < N; i += 2) {
                    const float2 r = __fadd2_rn(make_float2(v[i], v[i + 1]), make_float2(a, a));
                    v[i] = r.x;
                    v[i + 1] = r.y;
                }
            } else {
#pragma unroll
                for (int i = 0; i < N; i++) v[i] = affine_rn(v[i], s1, a);
            }
            break;
        case JZ_STEP_ELEMINV:
#pragma unroll
            for (int i = 0; i < N; i++) v[i] = eleminv_rn(v[i], s1);
            break;
        default: break;
    }
}

struct ChainParams {
    int n;
    int kind[JZ_MAX_CHAIN];
    float s1[JZ_MAX_CHAIN];
    float a[JZ_MAX_CHAIN];
    // GEMM epilogues only: a broadcast stage applied BEFORE the steps, x = bias_s1*x + bias_s2*bias[dim == 1 ? row : col]
    // (W*x + b*ones(1,N), ml/layer.hpp:79,120), rounded exactly like the separate jz_add_bcast pass
    const float* bias;
    int bias_dim;
    float bias_s1, bias_s2;
};

__device__ __forceinline__ float apply_bias(float x, const ChainParams& c, size_t row, size_t col) {
    return __fadd_rn(__fmul_rn(c.bias_s1, x), __fmul_rn(c.bias_s2, c.bias[c.bias_dim == 1 ? row : col]));
}

// Applies the steps in order to a register tile.  `c` must live in SHARED memory (stage it with
// stage_chain): indexing a kernel-parameter struct with the runtime step counter makes nvcc emit a
// compare-and-select ladder over the whole constant bank (measured: 76 of 162 instructions per
// element in the first chain kernel), whereas an LDS per step is free.
template <int N>
__device__ __forceinline__ void apply_chain(float (&v)[N], const ChainParams& c) {
    const int n = c.n;
#pragma unroll 1
    for (int s = 0; s < n; s++) apply_step<N>(v, c.kind[s], c.s1[s], c.a[s]);
}

// copy of the kernel-parameter chain into shared memory (call before __syncthreads).  ONE thread copies the struct
// with static offsets (constant-bank loads): indexing the parameter with the thread id instead makes nvcc spill a
// private copy of the whole struct to every thread's local memory first (128 bytes of stack, measured as a drop of
// the fused chain from 0.86 to 0.68 of the HBM peak when the struct grew past 25 words).
__device__ __forceinline__ void stage_chain(ChainParams* dst, const ChainParams& src, int tid) {
    if (tid == 0) *dst = src;
}

inline int make_chain(ChainParams& c, const jz_step* steps, int nsteps) {
    if (nsteps < 0 || nsteps > JZ_MAX_CHAIN || (nsteps > 0 && !steps)) return JZ_ERR_ARG;
    c.n = nsteps;
    c.bias = nullptr;
    c.bias_dim = 0;
    c.bias_s1 = c.bias_s2 = 0.0f;
    for (int i = 0; i < nsteps; i++) {
        const int k = steps[i].kind;
        if (!((k >= 0 && k < JZ_UNARY_COUNT) || k == JZ_STEP_AFFINE || k == JZ_STEP_ELEMINV)) return JZ_ERR_ARG;
        c.kind[i] = k;
        c.s1[i] = steps[i].s1;
        c.a[i] = steps[i].a;
    }
    return JZ_OK;
}

}  // namespace jz
