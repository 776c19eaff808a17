/* include/jz_b200.h -- the drop-in boundary of the B200-native Matrix<CUDAfloat>
 * compute layer: a flat C ABI (plain pointers, sizes, int return codes; no C++ or
 * torch types), in the style of the reference's own flat backend boundaries
 * (cpp/hipbackend.hpp:12-95, cpp/metal/MPSWrapper.h:9-76).
 *
 * Everything the reference's L3 layer (cpp/cumatrix.cu, cpp/cukernels.cu,
 * cpp/launcher.cu, cpp/cumatrix.cuh:67-85) does on the device is reachable through
 * these entry points; the C++ class `Matrix<CUDAfloat>` shipped in
 * juzhen_b200/cpp/cumatrix.cuh is a thin host-side shell over them.
 *
 * Conventions
 *  - All device matrices are fp32, COLUMN-MAJOR with a leading dimension `ld`
 *    (elements), exactly the reference's physical layout (cpp/core.hpp:96-98).  The
 *    reference's lazy `transpose` flag is passed explicitly as `trans` arguments.
 *  - Every compute call is asynchronous on `stream` (a cudaStream_t passed as void*;
 *    NULL = the legacy default stream the reference uses, cpp/cumatrix.cuh:59-62) and
 *    returns JZ_OK or an error code; jz_last_error() gives the text.  Shape errors are
 *    JZ_ERR_SHAPE *before* any launch (the C++ shell turns them into
 *    std::invalid_argument, cpp/cumatrix.cu:181-184,230-232).
 *  - There is NO CPU fallback anywhere behind this header: without a CUDA device every
 *    compute entry point fails with JZ_ERR_CUDA.
 */
#ifndef JZ_B200_H
#define JZ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define JZ_ABI_VERSION 1

typedef void* jz_stream_t; /* cudaStream_t */

enum jz_status {
    JZ_OK = 0,
    JZ_ERR_SHAPE = 1,       /* incompatible dimensions (reference: std::invalid_argument) */
    JZ_ERR_CUDA = 2,        /* CUDA runtime / driver error (reference: LOG_ERROR + exit(1)) */
    JZ_ERR_OOM = 3,         /* device allocation failed (reference: bad_alloc -> exit(1)) */
    JZ_ERR_ARG = 4,         /* null pointer / bad enum */
    JZ_ERR_UNSUPPORTED = 5  /* e.g. tcgen05 GEMM mode requested on a non-sm_100 device */
};

/* unary map selector for jz_unary (ids shared with oracle/) */
enum jz_unary_op {
    JZ_EXP = 0,    /* exp(Matrix)      cpp/cukernels.cu:92-106,156-171 */
    JZ_LOG = 1,    /* log(Matrix)      cpp/cukernels.cu:108-114,173-180 */
    JZ_TANH = 2,   /* tanh             cpp/cukernels.cu:116-130,182-199 */
    JZ_DTANH = 3,  /* 1 - tanh^2       cpp/cukernels.cu:132-146,201-221 */
    JZ_SQUARE = 4, /* x*x              cpp/cukernels.cu:40-54,224-239 */
    JZ_SQRT = 5,   /* sqrt lambda      cpp/juzhen.hpp:73-82 */
    JZ_RELU = 6,   /* relu lambda      ml/util.cuh:72-75 */
    JZ_DRELU = 7,  /* d_relu lambda    ml/util.cuh:77-80 */
    JZ_UNARY_COUNT = 8
};

/* GEMM arithmetic mode (replaces cublasSetMathMode / NVIDIA_TF32, cpp/launcher.cu:70-77) */
enum jz_gemm_mode {
    JZ_GEMM_3XTF32 = 0, /* default: fp32-accuracy emulation, 3 tcgen05 TF32 MMAs per product */
    JZ_GEMM_TF32 = 1,   /* NVIDIA_TF32=1 equivalent: single TF32 MMA */
    JZ_GEMM_FP32_SIMT = 2  /* plain fp32 FMA kernel (no tensor cores) */
    /* no bf16 mode: Matrix<CUDAfloat> stores fp32, and operands rounded to bf16 (8 significant bits) give a relative
       Frobenius error of ~3e-3 on the reference's randn inputs -- outside north_star's 1e-3 bound for the fast mode,
       which single-pass TF32 (7e-4) meets */
};

/* ---- runtime / lifetime (replaces main()'s handle + pool setup, cpp/launcher.cu:44-101) */
int jz_abi_version(void);
int jz_init(int device);             /* idempotent; selects device, reads JZ_GEMM_MODE / NVIDIA_TF32 */
int jz_shutdown(void);               /* releases the pool (== ~Memory<CUDAfloat>, cpp/memory.hpp:103-119) */
const char* jz_last_error(void);
int jz_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);
int jz_sync(jz_stream_t stream);
uint64_t jz_launch_count(void);      /* kernels launched by this library since init (bench gpu_launches) */
int jz_set_gemm_mode(int mode);
int jz_get_gemm_mode(void);

/* ---- stream-ordered pool allocator (replaces Memory<CUDAfloat>, cpp/memory.hpp:50-119 +
 *      cpp/cumatrix.cuh:67-85).  count is in floats; count==0 allocates 1 element like the
 *      reference.  A block freed on `stream` is immediately reusable by later work on the
 *      same stream; reuse from another stream waits on an event. */
int jz_malloc(float** ptr, size_t count, jz_stream_t stream);
int jz_free(float* ptr, jz_stream_t stream);
int jz_pool_trim(void);              /* return all cached (dead) blocks to the driver */
int jz_pool_stats(size_t* live_bytes, size_t* cached_bytes, size_t* n_device_allocs, size_t* n_hits);

/* ---- copies (replaces the sync cudaMemcpy calls, cpp/cumatrix.cu:42,77,108,155) */
int jz_memcpy_h2d(float* dst_dev, const float* src_host, size_t count, jz_stream_t stream);
int jz_memcpy_d2h(float* dst_host, const float* src_dev, size_t count, jz_stream_t stream); /* syncs stream */
/* upload from pageable memory through a pinned staging ring: src_host may be reused on return, and the host does
   not wait for earlier work in the stream (the per-step batch upload of examples/demo_mnist.cu:108-109, which the
   reference does with a synchronous cudaMemcpy, cpp/cumatrix.cu:28-48) */
int jz_upload(float* dst_dev, const float* src_host, size_t count, jz_stream_t stream);
int jz_memcpy_d2d(float* dst_dev, const float* src_dev, size_t count, jz_stream_t stream);

/* ---- flat elementwise maps over n contiguous floats (out may alias in) */
int jz_fill(float* x, size_t n, float value, jz_stream_t stream);                       /* fillKernel cpp/cukernels.cu:32-38,148-153 */
int jz_copy(float* dst, const float* src, size_t n, jz_stream_t stream);                /* copyKernel cpp/cukernels.cu:64-70 */
int jz_affine(float* out, const float* in, size_t n, float s1, float a, jz_stream_t stream); /* s1*x+a: addKernel :24-30, add(a,s1) cumatrix.cu:199-224 */
int jz_eleminv(float* out, const float* in, size_t n, float l, jz_stream_t stream);     /* l/x: divKernel cumatrix.cu:263-303 */
int jz_unary(int op, float* out, const float* in, size_t n, jz_stream_t stream);        /* exp/log/tanh/d_tanh/square/... */
int jz_axpby(float* out, const float* a, const float* b, size_t n, float s1, float s2, jz_stream_t stream); /* same-layout geam, cumatrix.cu:227-260 */
int jz_hadamard(float* out, const float* a, const float* b, size_t n, jz_stream_t stream); /* productKernel cukernels.cu:56-62 */
int jz_div(float* out, const float* a, const float* b, size_t n, jz_stream_t stream);   /* a * (1/b): operators.hpp:270-280 */

/* A fused chain of elementwise steps applied in one pass (what an rvalue operator chain such as
 * log(exp(X)+1)/5 does in 4 launches in the reference).  Each step is
 * {kind, s1, a}: kind < JZ_UNARY_COUNT = that unary map; JZ_STEP_AFFINE = s1*x+a;
 * JZ_STEP_ELEMINV = s1/x.  Results are bit-identical to running the steps one by one. */
#define JZ_STEP_AFFINE 100
#define JZ_STEP_ELEMINV 101
#define JZ_MAX_CHAIN 8
typedef struct jz_step { int kind; float s1; float a; } jz_step;
int jz_chain(float* out, const float* in, size_t n, const jz_step* steps, int nsteps, jz_stream_t stream);

/* ---- transpose-aware 2-D ops.  out is rows x cols (ldo); X_trans != 0 means the logical
 *      rows x cols operand is stored as its transpose (cols x rows, ld = ldx). */
int jz_axpby2d(float* out, size_t ldo, size_t rows, size_t cols,
               const float* a, size_t lda, int a_trans,
               const float* b, size_t ldb, int b_trans,
               float s1, float s2, jz_stream_t stream);                                  /* cublasSgeam, cumatrix.cu:227-260 */
int jz_hadamard2d(float* out, size_t ldo, size_t rows, size_t cols,
                  const float* a, size_t lda, int a_trans,
                  const float* b, size_t ldb, int b_trans, jz_stream_t stream);          /* hadmd mixed flags, cukernels.cu:326-400 */
/* dst(rows x cols, ldd)(i,j) = trans ? src(j,i) : src(i,j); bit-exact.  Covers T()
 * materialisation (geam sites cukernels.cu:289,315), slice get/set (copyKernel :72-90),
 * hstack/vstack placement (:258-323). */
int jz_copy2d(float* dst, size_t ldd, const float* src, size_t lds, size_t rows, size_t cols,
              int trans, jz_stream_t stream);

/* ---- reductions over a physical (non-transposed) col-major rows x cols matrix.
 *      dim 0: one result per column (len cols); dim 1: one result per row (len rows). */
int jz_sum(float* out, const float* a, size_t rows, size_t cols, size_t ld, int dim, jz_stream_t stream); /* sum(): Sgemv+ones, cumatrix.cu:312-337 */
int jz_max(float* out, const float* a, size_t rows, size_t cols, size_t ld, int dim, jz_stream_t stream); /* reduce<max functor>, cumatrix.cuh:334-384 + ml/layer.hpp:254-259 */
/* column softmax with max subtraction: out(:,j) = exp(x - max_j) / sum_j  (ml/layer.hpp:254-262 composite) */
int jz_softmax_cols(float* out, const float* a, size_t rows, size_t cols, size_t ld, jz_stream_t stream);
/* softmax-CE head gradient  G = -(Y - softmax(X)) / nb  (LogisticLayer::grad, ml/layer.hpp:252-264) */
int jz_softmax_ce_grad(float* out, const float* x, const float* y, size_t rows, size_t cols, float nb, jz_stream_t stream);
/* same with the factor given as it reaches the kernel: G = ((-(Y - softmax(X))) + 0) * rnb + 0, rnb = (float)(1.0/nb) */
int jz_softmax_ce_grad_scaled(float* out, const float* x, const float* y, size_t rows, size_t cols, float rnb, jz_stream_t stream);
int jz_nrm2(const float* x, size_t n, float* result_host, jz_stream_t stream);          /* cublasSnrm2, cumatrix.cu:168-175 (syncs) */

/* broadcast idioms the reference spells as rank-1 GEMMs (b*ones(1,N), ones(m,1)*v):
 * out(i,j) = s1*a(i,j) + s2*v[i] (dim 1: v has len rows) or + s2*v[j] (dim 0: len cols) */
int jz_add_bcast(float* out, const float* a, size_t rows, size_t cols, const float* v, int dim,
                 float s1, float s2, jz_stream_t stream);
/* out(i,j) = u[i]*v[j]  (k == 1 GEMM) */
int jz_outer(float* out, size_t ldo, const float* u, size_t rows, const float* v, size_t cols, jz_stream_t stream);

/* ---- GEMM (replaces cublasSgemm behind Matrix::dot, cumatrix.cu:177-197):
 *      C(m x n, ldc) = alpha * op(A)(m x k) * op(B)(k x n) + beta * C, column-major.
 *      mode: one of jz_gemm_mode, or -1 for the process-wide mode. */
int jz_gemm(int transA, int transB, size_t m, size_t n, size_t k, float alpha,
            const float* A, size_t lda, const float* B, size_t ldb, float beta,
            float* C, size_t ldc, int mode, jz_stream_t stream);
/* same product with an elementwise chain fused into the epilogue (C = chain(alpha*A*B)) */
int jz_gemm_chain(int transA, int transB, size_t m, size_t n, size_t k, float alpha,
                  const float* A, size_t lda, const float* B, size_t ldb,
                  float* C, size_t ldc, const jz_step* steps, int nsteps, int mode, jz_stream_t stream);
/* same, with a broadcast stage in front of the steps: C = chain(s1*(alpha*A*B) + s2*bias[i]) (bias_dim 1: one value
 * per row, length m) or + s2*bias[j] (bias_dim 0: per column, length n).  This is W*x + b*ones(1,N) followed by the
 * activation (Layer::eval / Layer::grad, ml/layer.hpp:79,120) as ONE kernel; bit-identical to jz_gemm + jz_add_bcast +
 * jz_chain.  bias == NULL: no broadcast stage. */
int jz_gemm_bias_chain(int transA, int transB, size_t m, size_t n, size_t k, float alpha,
                       const float* A, size_t lda, const float* B, size_t ldb,
                       float* C, size_t ldc, const float* bias, int bias_dim, float s1, float s2,
                       const jz_step* steps, int nsteps, int mode, jz_stream_t stream);
/* Multi-GPU form (column-sharded C = A * B[:, block], then all-gather): the same fused product,
 * whose epilogue ALSO stores every finished tile into `n_peers` more images of C with the same ldc --
 * buffers of peer GPUs mapped into this process (CUDA P2P / symmetric memory over NVLink).  The gather
 * overlaps the math tile by tile; the caller only needs a cross-rank barrier afterwards.  Replaces nothing
 * in the reference (it has no collectives); the NCCL all-gather variant lives in juzhen_b200/mg.py. */
#define JZ_MAX_PEERS 7
int jz_gemm_chain_bcast(int transA, int transB, size_t m, size_t n, size_t k, float alpha,
                        const float* A, size_t lda, const float* B, size_t ldb,
                        float* C, size_t ldc, float* const* peer_C, int n_peers,
                        const jz_step* steps, int nsteps, int mode, jz_stream_t stream);
/* Same, through an NVSwitch MULTICAST mapping of C (mc_C = the multicast address of the same block as C in a buffer
 * bound on every GPU, e.g. torch symmetric memory's multicast_ptr or cuMulticast*): the epilogue issues ONE
 * multimem.st per element and the switch replicates it into every GPU's image, this one included -- 1/N of the
 * NVLink egress of the per-peer form. */
int jz_gemm_chain_mcast(int transA, int transB, size_t m, size_t n, size_t k, float alpha,
                        const float* A, size_t lda, const float* B, size_t ldb,
                        float* C, size_t ldc, float* mc_C,
                        const jz_step* steps, int nsteps, int mode, jz_stream_t stream);
/* strided batch: member i is C + i*strideC = alpha * op(A + i*strideA) * op(B + i*strideB) + beta * (C + i*strideC)
 * (cublasSgemmStridedBatched in TransformerLayer's attention, ml/layer.hpp:2896-2926, 3089-3283) */
int jz_gemm_strided_batched(int transA, int transB, size_t m, size_t n, size_t k, float alpha,
                            const float* A, size_t lda, size_t strideA, const float* B, size_t ldb, size_t strideB,
                            float beta, float* C, size_t ldc, size_t strideC, size_t batch, int mode, jz_stream_t stream);
/* which kernel family the last jz_gemm used: 0 none, 1 tcgen05, 2 simt, 3 outer/gemv special case, 4 small-product */
int jz_gemm_last_path(void);
/* k-splits per tile of the partial last wave in the last tensor-core launch (1 = no split-K units) */
int jz_gemm_last_splits(void);
/* 1 when those k-splits were the CTAs of one thread-block cluster per tile and exchanged their partial tiles through
 * distributed shared memory (no workspace); 0 for the workspace + ticket form or no split */
int jz_gemm_last_cluster_split(void);
/* 1 when the last tensor-core strided batch was WALKED: one CTA (pair) per SM went through the (member, tile) units with its
 * barriers, tensor memory and TMA pipeline alive across units, instead of one CTA per unit */
int jz_gemm_last_walk(void);

/* ---- multi-GPU, one process per GPU (SURVEY 8b jz_mg_*, 8e).  The reference has no collectives; these entry points
 *      give C / C++ callers the sharded forms without torch, NCCL or MPI inside the library: peers' buffers are mapped
 *      with CUDA IPC (P2P over NVLink), and the CALLER carries the opaque handle bytes between its processes (pipe,
 *      MPI, shared memory, torch.distributed ...).  World size <= JZ_MAX_PEERS + 1 = 8 (one NVSwitch domain). */
#define JZ_MG_HANDLE_BYTES 128
/* contiguous share [begin, end) of n items for `rank`; the first n % world ranks get one more */
int jz_mg_block_range(size_t n, int world, int rank, size_t* begin, size_t* end);
/* handle (JZ_MG_HANDLE_BYTES) of a device buffer obtained from jz_malloc / cudaMalloc, for another process to import */
int jz_mg_export(const float* dev_ptr, void* handle);
/* map a peer process' buffer into this process (peer access is enabled lazily); release with jz_mg_release */
int jz_mg_import(const void* handle, float** peer_ptr);
int jz_mg_release(float* peer_ptr);
/* device-side barrier on `stream`: flags[r] = rank r's array of `world` unsigned (zero-initialised once, exported /
 * imported like any buffer; flags[rank] is this rank's own).  epoch must grow by 1 per barrier.  Everything the
 * stream did before the barrier is visible to every rank's work after it. */
int jz_mg_barrier(unsigned* const* flags, int world, int rank, unsigned epoch, jz_stream_t stream);
/* Column-sharded GEMM with the all-gather fused into the tensor-core epilogue: this rank computes
 * C[:, j0:j1] = chain(alpha * op(A) * B_block) (j0, j1 = jz_mg_block_range(n, world, rank); B_block = its k x (j1-j0)
 * column block, ldb) and stores every finished tile into EVERY rank's dense m x n image c_images[r] (c_images[rank]
 * = its own).  Follow with jz_mg_barrier before reading the other ranks' columns. */
int jz_mg_gemm_allgather(int transA, size_t m, size_t n, size_t k, float alpha, const float* A, size_t lda,
                         const float* B_block, size_t ldb, float* const* c_images, int world, int rank,
                         const jz_step* steps, int nsteps, int mode, jz_stream_t stream);
/* out[i] = sum over ranks, in rank order, of partial_images[r][i] (column sums over row-sharded data: local jz_sum into
 * this rank's exported partial vector, jz_mg_barrier, then this; bitwise identical on every rank) */
int jz_mg_allreduce_sum(float* out, float* const* partial_images, size_t n, int world, int rank, jz_stream_t stream);

/* ---- transformer helper kernels (the reference's own __global__ kernels in ml/layer.hpp).
 *      Attention scores: (seq_len, seq_len*batch) column-major, block i = the seq_len x seq_len matrix at
 *      x + i*seq_len^2, element (query a, key b) at a + b*seq_len.  LayerNorm tensors: (dim, n) column-major. */
/* y(a,:) = softmax over keys of x(a,:), per block; causal != 0 first replaces x(a,b), b > a, by mask_val
   (softmax_rows_batched_kernel ml/layer.hpp:2373-2398 with causal_mask_kernel :2400-2412 fused; y may alias x) */
int jz_softmax_rows_batched(float* y, const float* x, size_t seq_len, size_t batch, int causal, float mask_val,
                            jz_stream_t stream);
int jz_causal_mask(float* s_inout, size_t seq_len, size_t batch, float mask_val, jz_stream_t stream); /* :2400-2412 */
/* dS(a,b) = A(a,b) * (dA(a,b) - sum_b' A(a,b') dA(a,b')) * scale, dA(a,b) = dAT[b + a*seq_len] per block
   (softmax_backward_rows_kernel ml/layer.hpp:2418-2445) */
int jz_softmax_rows_backward(float* dS, const float* A, const float* dAT, size_t seq_len, size_t batch, float scale,
                             jz_stream_t stream);
/* y = gamma .* xhat + beta, xhat = (x - mean)/sqrt(var + 1e-5) per column; also stores xhat and 1/sqrt(var + 1e-5)
   (layernorm_forward_kernel ml/layer.hpp:2483-2510) */
int jz_layernorm_forward(float* y, float* xhat, float* inv_std, const float* x, const float* gamma, const float* beta,
                         size_t dim, size_t n, jz_stream_t stream);
/* dx = inv_std .* (dxhat - mean(dxhat) - xhat .* mean(dxhat .* xhat)), dxhat = gamma .* dy
   (layernorm_backward_kernel ml/layer.hpp:2514-2538) */
int jz_layernorm_backward(float* dx, const float* dy, const float* gamma, const float* xhat, const float* inv_std,
                          size_t dim, size_t n, jz_stream_t stream);

/* ---- RNG (replaces cuRAND XORWOW, cumatrix.cu:354-420; counter-based Philox4x32-10).  `offset` selects an independent
 *      STREAM of the generator (it is the upper half of the Philox counter, not an element offset): the same
 *      (seed, offset, n) always yields the same values; two draws that must differ need different offsets. */
int jz_rand_uniform(float* x, size_t n, uint64_t seed, uint64_t offset, jz_stream_t stream);
int jz_rand_normal(float* x, size_t n, uint64_t seed, uint64_t offset, jz_stream_t stream);

/* ---- fused Adam step (adam_update_kernel, ml/util.cuh:152-163) */
int jz_adam_update(float* g, float* m, float* v, size_t n, float alpha, float beta1, float beta2,
                   float eps, float bc1, float bc2, jz_stream_t stream);

/* ---- accuracy self-check: max ulp distance of unary `op` against an fp64 device
 *      evaluation over every fp32 bit pattern in [lo_bits, hi_bits) (test support) */
int jz_unary_ulp_sweep(int op, uint32_t lo_bits, uint32_t hi_bits, uint32_t* max_ulp_host,
                       uint32_t* worst_bits_host, jz_stream_t stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* JZ_B200_H */
