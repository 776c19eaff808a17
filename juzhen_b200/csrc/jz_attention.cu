// jz_attention.cu -- the transformer helper kernels of ml/layer.hpp: batched row softmax with the
// causal mask, its backward, LayerNorm forward / input-gradient.  The reference runs each as one THREAD per row
// (or per column) looping serially over the other dimension (ml/layer.hpp:2373-2445, 2483-2538): three passes over
// global memory for the softmax, stride-`dim` (uncoalesced) accesses for LayerNorm, and only seq_len*batch threads
// in flight.  Here every global access is a coalesced 128-byte row segment, the reduced dimension is split over
// the warps of a CTA (softmax) or the lanes of a warp (LayerNorm), and the softmax tile lives in shared memory
// between its passes: one read and one write of HBM per element.
//
// Layouts are the reference's: attention scores are (seq_len, seq_len*batchN) column-major -- block `blk` is the
// contiguous seq_len x seq_len column-major matrix at x + blk*seq_len^2, query rows a, keys b, element (a, b) at
// a + b*seq_len; LayerNorm tensors are (dim, N) column-major, one sample per contiguous column.
#include "jz_common.cuh"
#include "jz_math.cuh"

namespace jz {

constexpr int AT_WARPS = 8;

// fold one float per (lane, warp) over the W warps of the CTA; result valid in every thread
template <int W, class F>
__device__ __forceinline__ float fold_warps(float v, float (*red)[32], F f, float init) {
    const int lane = threadIdx.x, w = threadIdx.y;
    __syncthreads();   // previous use of `red` is over
    red[w][lane] = v;
    __syncthreads();
    float t = init;
#pragma unroll
    for (int q = 0; q < W; q++) t = f(t, red[q][lane]);
    return t;
}

// y(a, b) = softmax over b of x(a, b) [with x(a, b) := mask_val for b > a when CAUSAL], per block.
// grid (ceil(S/32), batch), block (32, 8): lanes own rows, warps stride over keys.  STAGED: the 32 x S tile is kept
// in shared memory (S*128 bytes) between the max, exp-sum and normalise passes; otherwise y is the scratch, as in
// the reference.  ml/layer.hpp:2373-2398 (+ causal_mask_kernel :2400-2412 fused as a flag).
template <bool CAUSAL, bool STAGED, int W>
__global__ void __launch_bounds__(32 * W) softmax_rows_kernel(float* y, const float* x, int S, float mask_val) {
    pdl_enter();
    extern __shared__ float tile[];   // [S][32] when STAGED
    __shared__ float red[W][32];
    const int lane = threadIdx.x, w = threadIdx.y;
    const int a = blockIdx.x * 32 + lane;
    const bool ok = a < S;
    const size_t base = size_t(blockIdx.y) * size_t(S) * size_t(S) + size_t(ok ? a : 0);
    float m = -1e30f;
#pragma unroll 8
    for (int b = w; b < S; b += W) {
        float v = x[base + size_t(b) * S];
        if (CAUSAL && b > a) v = mask_val;
        if (STAGED) tile[b * 32 + lane] = v;
        m = m > v ? m : v;
    }
    m = fold_warps<W>(m, red, [](float p, float q) { return p > q ? p : q; }, -1e30f);
    float s = 0.0f;
#pragma unroll 8
    for (int b = w; b < S; b += W) {
        float v;
        if (STAGED) v = tile[b * 32 + lane];
        else { v = x[base + size_t(b) * S]; if (CAUSAL && b > a) v = mask_val; }
        const float e = expf(v - m);
        if (STAGED) tile[b * 32 + lane] = e;
        else if (ok) y[base + size_t(b) * S] = e;
        s += e;
    }
    s = fold_warps<W>(s, red, [](float p, float q) { return p + q; }, 0.0f);
    const float inv = 1.0f / (s + 1e-12f);
    if (!ok) return;
#pragma unroll 8
    for (int b = w; b < S; b += W) {
        const float e = STAGED ? tile[b * 32 + lane] : y[base + size_t(b) * S];
        y[base + size_t(b) * S] = e * inv;
    }
}

// s(a, b) = mask_val for b > a (ml/layer.hpp:2400-2412), in place, for callers that want the masked scores themselves
__global__ void __launch_bounds__(256) causal_mask_kernel(float* s, size_t S, size_t total, float mask_val) {
    pdl_enter();
    for (size_t idx = size_t(blockIdx.x) * 256 + threadIdx.x; idx < total; idx += size_t(gridDim.x) * 256) {
        const size_t local = idx % (S * S);
        if (local / S > local % S) s[idx] = mask_val;
    }
}

// dS(a, b) = A(a, b) * (dA(a, b) - sum_b' A(a, b') dA(a, b')) * scale with dA given TRANSPOSED per block:
// dA(a, b) = dAT[b + a*S] (ml/layer.hpp:2418-2445).  The 32 columns of dAT this CTA needs are contiguous runs of S
// floats: they are read coalesced along b and parked transposed in shared memory ([S][33]), so both operands are
// consumed with lanes along a.
__global__ void __launch_bounds__(32 * AT_WARPS) softmax_rows_backward_kernel(float* dS, const float* A, const float* dAT,
                                                                              int S, float scale, unsigned row_blocks) {
    pdl_enter();
    extern __shared__ float dtile[];   // [S][33]: dtile[b*33 + a_local] = dA(a0 + a_local, b)
    __shared__ float red[AT_WARPS][32];
    const int lane = threadIdx.x, w = threadIdx.y;
    const int a0 = int(blockIdx.x % row_blocks) * 32;   // linear grid: row block + row_blocks * batch member
    const size_t blk = size_t(blockIdx.x / row_blocks) * size_t(S) * size_t(S);
    // stage: warp w copies columns a_local = w, w + 8, ... of dAT (each S contiguous floats), lanes along b
    for (int al = w; al < 32; al += AT_WARPS) {
        const int a = a0 + al;
#pragma unroll 8
        for (int b = lane; b < S; b += 32) dtile[b * 33 + al] = a < S ? dAT[blk + size_t(a) * S + b] : 0.0f;
    }
    __syncthreads();
    const int a = a0 + lane;
    const bool ok = a < S;
    const size_t base = blk + size_t(ok ? a : 0);
    float rs = 0.0f;
#pragma unroll 8
    for (int b = w; b < S; b += AT_WARPS) rs += A[base + size_t(b) * S] * dtile[b * 33 + lane];
    rs = fold_warps<AT_WARPS>(rs, red, [](float p, float q) { return p + q; }, 0.0f);
    if (!ok) return;
#pragma unroll 8
    for (int b = w; b < S; b += AT_WARPS) {
        const float av = A[base + size_t(b) * S];
        dS[base + size_t(b) * S] = av * (dtile[b * 33 + lane] - rs) * scale;
    }
}

// Same, for key ranges that do not fit in shared memory at once (seq_len > 1551; the reference kernel has no limit):
// keys are staged `chunk` at a time, once for the row sums and once more for the output (the second read of dAT comes
// from L2).  The per-warp partial sums accumulate over the chunks in the same b order as the one-pass kernel.
__global__ void __launch_bounds__(32 * AT_WARPS) softmax_rows_backward_long_kernel(float* dS, const float* A, const float* dAT,
                                                                                   int S, float scale, unsigned row_blocks, int chunk) {
    pdl_enter();
    extern __shared__ float dtile[];   // [chunk][33]
    __shared__ float red[AT_WARPS][32];
    const int lane = threadIdx.x, w = threadIdx.y;
    const int a0 = int(blockIdx.x % row_blocks) * 32;
    const size_t blk = size_t(blockIdx.x / row_blocks) * size_t(S) * size_t(S);
    const int a = a0 + lane;
    const bool ok = a < S;
    const size_t base = blk + size_t(ok ? a : 0);
    float rs = 0.0f;
    for (int pass = 0; pass < 2; pass++) {
        for (int b0 = 0; b0 < S; b0 += chunk) {
            const int nb = S - b0 < chunk ? S - b0 : chunk;
            for (int al = w; al < 32; al += AT_WARPS) {
                const int aa = a0 + al;
#pragma unroll 8
                for (int b = lane; b < nb; b += 32) dtile[b * 33 + al] = aa < S ? dAT[blk + size_t(aa) * S + b0 + b] : 0.0f;
            }
            __syncthreads();
            if (pass == 0) {
#pragma unroll 8
                for (int b = w; b < nb; b += AT_WARPS) rs += A[base + size_t(b0 + b) * S] * dtile[b * 33 + lane];
            } else if (ok) {
#pragma unroll 8
                for (int b = w; b < nb; b += AT_WARPS) {
                    const float av = A[base + size_t(b0 + b) * S];
                    dS[base + size_t(b0 + b) * S] = av * (dtile[b * 33 + lane] - rs) * scale;
                }
            }
            __syncthreads();
        }
        if (pass == 0) rs = fold_warps<AT_WARPS>(rs, red, [](float p, float q) { return p + q; }, 0.0f);
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One WARP per column c of the (dim, N) tensor (ml/layer.hpp:2483-2510):
//   mu = mean(x_c), var = mean((x_c - mu)^2), inv = rsqrt(var + 1e-5), xhat = (x - mu)*inv, y = gamma*xhat + beta.
// The column is re-read from L1 for the second and third pass (it was just loaded); HBM sees one read, two writes.
// VEC: dim % 4 == 0 and 16-byte aligned tensors -> 128-bit accesses.
template <bool VEC>
__global__ void __launch_bounds__(32 * AT_WARPS) layernorm_forward_kernel(float* y, float* xhat, float* inv_std, const float* x,
                                                                          const float* gamma, const float* beta, int dim, size_t N) {
    pdl_enter();
    const int lane = threadIdx.x;
    const size_t c = size_t(blockIdx.x) * AT_WARPS + threadIdx.y;
    if (c >= N) return;
    const float* xc = x + c * size_t(dim);
    float mu = 0.0f, var = 0.0f;
    if constexpr (VEC) {
        const float4* x4 = reinterpret_cast<const float4*>(xc);
        const int d4 = dim >> 2;
#pragma unroll 4
        for (int i = lane; i < d4; i += 32) { const float4 v = x4[i]; mu += (v.x + v.y) + (v.z + v.w); }
        mu = warp_sum(mu) / dim;
#pragma unroll 4
        for (int i = lane; i < d4; i += 32) {
            const float4 v = x4[i];
            const float a0 = v.x - mu, a1 = v.y - mu, a2 = v.z - mu, a3 = v.w - mu;
            var += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
        }
    } else {
        for (int i = lane; i < dim; i += 32) mu += xc[i];
        mu = warp_sum(mu) / dim;
        for (int i = lane; i < dim; i += 32) { const float d = xc[i] - mu; var += d * d; }
    }
    var = warp_sum(var) / dim;
    const float inv = rsqrtf(var + 1e-5f);
    if (lane == 0) inv_std[c] = inv;
    if constexpr (VEC) {
        const float4* x4 = reinterpret_cast<const float4*>(xc);
        const float4* g4 = reinterpret_cast<const float4*>(gamma);
        const float4* b4 = reinterpret_cast<const float4*>(beta);
        float4* h4 = reinterpret_cast<float4*>(xhat + c * size_t(dim));
        float4* y4 = reinterpret_cast<float4*>(y + c * size_t(dim));
        const int d4 = dim >> 2;
#pragma unroll 4
        for (int i = lane; i < d4; i += 32) {
            const float4 v = x4[i], g = g4[i], b = b4[i];
            const float4 h = make_float4((v.x - mu) * inv, (v.y - mu) * inv, (v.z - mu) * inv, (v.w - mu) * inv);
            h4[i] = h;
            y4[i] = make_float4(g.x * h.x + b.x, g.y * h.y + b.y, g.z * h.z + b.z, g.w * h.w + b.w);
        }
    } else {
        for (int i = lane; i < dim; i += 32) {
            const float xh = (xc[i] - mu) * inv;
            xhat[c * size_t(dim) + i] = xh;
            y[c * size_t(dim) + i] = gamma[i] * xh + beta[i];
        }
    }
}

// dx = inv_std * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)), dxhat = gamma * dy (ml/layer.hpp:2514-2538)
template <bool VEC>
__global__ void __launch_bounds__(32 * AT_WARPS) layernorm_backward_kernel(float* dx, const float* dy, const float* gamma,
                                                                           const float* xhat, const float* inv_std, int dim, size_t N) {
    pdl_enter();
    const int lane = threadIdx.x;
    const size_t c = size_t(blockIdx.x) * AT_WARPS + threadIdx.y;
    if (c >= N) return;
    const size_t c0 = c * size_t(dim);
    float m1 = 0.0f, m2 = 0.0f;
    if constexpr (VEC) {
        const float4* g4 = reinterpret_cast<const float4*>(gamma);
        const float4* d4p = reinterpret_cast<const float4*>(dy + c0);
        const float4* h4 = reinterpret_cast<const float4*>(xhat + c0);
        float4* o4 = reinterpret_cast<float4*>(dx + c0);
        const int d4 = dim >> 2;
#pragma unroll 4
        for (int i = lane; i < d4; i += 32) {
            const float4 g = g4[i], d = d4p[i], h = h4[i];
            const float t0 = __fmul_rn(g.x, d.x), t1 = __fmul_rn(g.y, d.y), t2 = __fmul_rn(g.z, d.z), t3 = __fmul_rn(g.w, d.w);
            m1 += (t0 + t1) + (t2 + t3);
            m2 += (t0 * h.x + t1 * h.y) + (t2 * h.z + t3 * h.w);
        }
        m1 = warp_sum(m1) / dim;
        m2 = warp_sum(m2) / dim;
        const float inv = inv_std[c];
#pragma unroll 4
        for (int i = lane; i < d4; i += 32) {
            const float4 g = g4[i], d = d4p[i], h = h4[i];
            const float t0 = __fmul_rn(g.x, d.x), t1 = __fmul_rn(g.y, d.y), t2 = __fmul_rn(g.z, d.z), t3 = __fmul_rn(g.w, d.w);
            o4[i] = make_float4(inv * (t0 - m1 - h.x * m2), inv * (t1 - m1 - h.y * m2), inv * (t2 - m1 - h.z * m2), inv * (t3 - m1 - h.w * m2));
        }
        return;
    }
    for (int i = lane; i < dim; i += 32) {
        const float dxh = __fmul_rn(gamma[i], dy[c0 + i]);   // one rounding, the same in both passes
        m1 += dxh;
        m2 += dxh * xhat[c0 + i];
    }
    m1 = warp_sum(m1) / dim;
    m2 = warp_sum(m2) / dim;
    const float inv = inv_std[c];
    for (int i = lane; i < dim; i += 32) {
        const float dxh = __fmul_rn(gamma[i], dy[c0 + i]);
        dx[c0 + i] = inv * (dxh - m1 - xhat[c0 + i] * m2);
    }
}

constexpr size_t kMaxDynSmem = 200 * 1024;

}  // namespace jz

using namespace jz;

extern "C" {

int jz_softmax_rows_batched(float* y, const float* x, size_t seq_len, size_t batch, int causal, float mask_val,
                            jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    if (seq_len == 0 || batch == 0) return JZ_OK;
    if (!y || !x) return fail(JZ_ERR_ARG, "jz_softmax_rows_batched: null pointer");
    if (seq_len >= (size_t(1) << 30) || batch > 65535) return fail(JZ_ERR_UNSUPPORTED, "jz_softmax_rows_batched: shape too large");
    cudaStream_t s = as_stream(stream);
    const dim3 grid((unsigned)ceil_div(seq_len, size_t(32)), (unsigned)batch, 1);
    const size_t smem = seq_len * 32 * sizeof(float);
    const int S = int(seq_len);
    // The tile stays in shared memory while several CTAs still fit on an SM (S <= 512: 64 KB); longer rows use y as the
    // scratch (second and third pass are L2 hits) with 16 warps per CTA, so enough loads are in flight.
    if (smem <= 64 * 1024) {
        static bool attr_done = false;
        if (!attr_done) {
            JZ_CUDA(cudaFuncSetAttribute(softmax_rows_kernel<true, true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            JZ_CUDA(cudaFuncSetAttribute(softmax_rows_kernel<false, true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            attr_done = true;
        }
        const dim3 block(32, 8, 1);
        if (causal) JZ_LAUNCH((softmax_rows_kernel<true, true, 8>), grid, block, smem, s, y, x, S, mask_val);
        else JZ_LAUNCH((softmax_rows_kernel<false, true, 8>), grid, block, smem, s, y, x, S, mask_val);
    } else {
        const dim3 block(32, 16, 1);
        if (causal) JZ_LAUNCH((softmax_rows_kernel<true, false, 16>), grid, block, 0, s, y, x, S, mask_val);
        else JZ_LAUNCH((softmax_rows_kernel<false, false, 16>), grid, block, 0, s, y, x, S, mask_val);
    }
    return JZ_OK;
}

int jz_causal_mask(float* s_inout, size_t seq_len, size_t batch, float mask_val, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    const size_t total = seq_len * seq_len * batch;
    if (total == 0) return JZ_OK;
    if (!s_inout) return fail(JZ_ERR_ARG, "jz_causal_mask: null pointer");
    const size_t cap = size_t(ctx().sm_count) * 8, blocks = ceil_div(total, size_t(256));
    JZ_LAUNCH(causal_mask_kernel, unsigned(blocks < cap ? blocks : cap), 256, 0, as_stream(stream), s_inout, seq_len, total, mask_val);
    return JZ_OK;
}

int jz_softmax_rows_backward(float* dS, const float* A, const float* dAT, size_t seq_len, size_t batch, float scale,
                             jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    if (seq_len == 0 || batch == 0) return JZ_OK;
    if (!dS || !A || !dAT) return fail(JZ_ERR_ARG, "jz_softmax_rows_backward: null pointer");
    const size_t row_blocks = ceil_div(seq_len, size_t(32));
    if (row_blocks * batch >= (size_t(1) << 31) || seq_len >= (size_t(1) << 31)) return fail(JZ_ERR_UNSUPPORTED, "jz_softmax_rows_backward: more than 2^31 row blocks");
    const unsigned grid = unsigned(row_blocks * batch);
    const dim3 block(32, AT_WARPS, 1);
    static bool attr_done = false;
    if (!attr_done) {
        JZ_CUDA(cudaFuncSetAttribute(softmax_rows_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kMaxDynSmem)));
        JZ_CUDA(cudaFuncSetAttribute(softmax_rows_backward_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kMaxDynSmem)));
        attr_done = true;
    }
    const size_t smem = seq_len * 33 * sizeof(float);
    if (smem <= kMaxDynSmem) {
        JZ_LAUNCH(softmax_rows_backward_kernel, grid, block, smem, as_stream(stream), dS, A, dAT, int(seq_len), scale, unsigned(row_blocks));
    } else {   // keys staged 768 at a time (two CTAs per SM stay resident)
        const int chunk = 768;
        JZ_LAUNCH(softmax_rows_backward_long_kernel, grid, block, size_t(chunk) * 33 * sizeof(float), as_stream(stream), dS, A, dAT,
                  int(seq_len), scale, unsigned(row_blocks), chunk);
    }
    return JZ_OK;
}

int jz_layernorm_forward(float* y, float* xhat, float* inv_std, const float* x, const float* gamma, const float* beta,
                         size_t dim, size_t n, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    if (dim == 0 || n == 0) return JZ_OK;
    if (!y || !xhat || !inv_std || !x || !gamma || !beta) return fail(JZ_ERR_ARG, "jz_layernorm_forward: null pointer");
    if (dim >= (size_t(1) << 31)) return fail(JZ_ERR_UNSUPPORTED, "jz_layernorm_forward: dim too large");
    const dim3 grid((unsigned)ceil_div(n, size_t(AT_WARPS)), 1, 1), block(32, AT_WARPS, 1);
    const bool vec = dim % 4 == 0 && aligned16(y) && aligned16(xhat) && aligned16(x) && aligned16(gamma) && aligned16(beta);
    if (vec) JZ_LAUNCH(layernorm_forward_kernel<true>, grid, block, 0, as_stream(stream), y, xhat, inv_std, x, gamma, beta, int(dim), n);
    else JZ_LAUNCH(layernorm_forward_kernel<false>, grid, block, 0, as_stream(stream), y, xhat, inv_std, x, gamma, beta, int(dim), n);
    return JZ_OK;
}

int jz_layernorm_backward(float* dx, const float* dy, const float* gamma, const float* xhat, const float* inv_std,
                          size_t dim, size_t n, jz_stream_t stream) {
    JZ_INIT_OR_RETURN();
    if (dim == 0 || n == 0) return JZ_OK;
    if (!dx || !dy || !gamma || !xhat || !inv_std) return fail(JZ_ERR_ARG, "jz_layernorm_backward: null pointer");
    if (dim >= (size_t(1) << 31)) return fail(JZ_ERR_UNSUPPORTED, "jz_layernorm_backward: dim too large");
    const dim3 grid((unsigned)ceil_div(n, size_t(AT_WARPS)), 1, 1), block(32, AT_WARPS, 1);
    const bool vec = dim % 4 == 0 && aligned16(dx) && aligned16(dy) && aligned16(gamma) && aligned16(xhat);
    if (vec) JZ_LAUNCH(layernorm_backward_kernel<true>, grid, block, 0, as_stream(stream), dx, dy, gamma, xhat, inv_std, int(dim), n);
    else JZ_LAUNCH(layernorm_backward_kernel<false>, grid, block, 0, as_stream(stream), dx, dy, gamma, xhat, inv_std, int(dim), n);
    return JZ_OK;
}

}  // extern "C"
