// tune_stream.cu -- design-space probe for the HBM-bound kernels (run on the B200 via gpurun; NOT product code).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tune_stream scripts/tune_stream.cu
// Measures GB/s (algorithmic 8 B/elem) of a 1R+1W map over 2^28 floats for: loads in flight per thread,
// threads per CTA, CTAs per SM, grid-stride vs one-tile-per-CTA, cache hints; and of transposes with
// different tile geometries.  The winners are folded back into jz_elementwise.cu / jz_layout.cu.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

enum Hint { PLAIN = 0, STREAM = 1, NOALLOC = 2 };

template <int H>
__device__ __forceinline__ float4 ld4(const float4* p) {
    float4 v;
    if (H == STREAM) asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    else if (H == NOALLOC) asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    else v = *p;
    return v;
}
template <int H>
__device__ __forceinline__ void st4(float4* p, float4 v) {
    if (H == STREAM) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    else *p = v;
}

__device__ __forceinline__ float f(float x) { return x * 2.0f + 1.0f; }

template <int U, int H>
__global__ void map_gs(float* out, const float* in, size_t n4) {  // grid-stride over tiles of blockDim*U float4
    const float4* in4 = (const float4*)in;
    float4* out4 = (float4*)out;
    const size_t tile = size_t(blockDim.x) * U;
    for (size_t base = size_t(blockIdx.x) * tile; base < n4; base += size_t(gridDim.x) * tile) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t i = base + size_t(u) * blockDim.x + threadIdx.x;
            if (i < n4) v[u] = ld4<H>(in4 + i);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t i = base + size_t(u) * blockDim.x + threadIdx.x;
            if (i < n4) st4<H>(out4 + i, make_float4(f(v[u].x), f(v[u].y), f(v[u].z), f(v[u].w)));
        }
    }
}

// transpose: dst(rows x cols, ld rows) = src^T, src is cols x rows (ld cols); tile TS x TS, float4 both sides
template <int TS, int PAD>
__global__ void transpose_t(float* dst, const float* src, size_t rows, size_t cols) {
    __shared__ float tile[TS][TS + PAD];
    constexpr int Q = TS / 4;            // float4 per tile row
    constexpr int RPP = 256 / Q;         // tile rows per pass
    const int tx = threadIdx.x % Q, ty = threadIdx.x / Q;
    const size_t tiles_i = rows / TS, tiles_j = cols / TS;
    for (size_t t = blockIdx.x; t < tiles_i * tiles_j; t += gridDim.x) {
        const size_t ti = t % tiles_i, tj = t / tiles_i;
        const size_t i0 = ti * TS, j0 = tj * TS;
#pragma unroll
        for (int r = 0; r < TS; r += RPP) {
            const int il = ty + r;
            const float4 v = *(const float4*)(src + (i0 + il) * cols + j0 + 4 * tx);
            tile[il][4 * tx + 0] = v.x; tile[il][4 * tx + 1] = v.y; tile[il][4 * tx + 2] = v.z; tile[il][4 * tx + 3] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < TS; r += RPP) {
            const int jl = ty + r;
            float4 v;
            v.x = tile[4 * tx + 0][jl]; v.y = tile[4 * tx + 1][jl]; v.z = tile[4 * tx + 2][jl]; v.w = tile[4 * tx + 3][jl];
            *(float4*)(dst + (j0 + jl) * rows + i0 + 4 * tx) = v;
        }
        __syncthreads();
    }
}

// transpose variant: 32 x 32 tile, each thread moves a 4 x 4 register block (float4 loads, in-register transpose
// through shared memory written as float4 rows with an XOR swizzle: no padding, conflict-free both ways)
__global__ void transpose_reg4(float* dst, const float* src, size_t rows, size_t cols) {
    __shared__ float4 tile[64][16];  // 64 x 64 floats
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, each 4 x 4 elements
    const size_t tiles_i = rows / 64, tiles_j = cols / 64;
    for (size_t t = blockIdx.x; t < tiles_i * tiles_j; t += gridDim.x) {
        const size_t ti = t % tiles_i, tj = t / tiles_i;
        const size_t i0 = ti * 64, j0 = tj * 64;
        float4 a[4];
#pragma unroll
        for (int r = 0; r < 4; r++) a[r] = *(const float4*)(src + (i0 + 4 * ty + r) * cols + j0 + 4 * tx);
        // a[r] = src rows (i = 4ty+r), cols j = 4tx..4tx+3 ; transposed block: b[c] = {a[0].c, a[1].c, a[2].c, a[3].c}
        float4 b[4];
        b[0] = make_float4(a[0].x, a[1].x, a[2].x, a[3].x);
        b[1] = make_float4(a[0].y, a[1].y, a[2].y, a[3].y);
        b[2] = make_float4(a[0].z, a[1].z, a[2].z, a[3].z);
        b[3] = make_float4(a[0].w, a[1].w, a[2].w, a[3].w);
        // b[c] belongs to dst row j = 4tx+c, floats i = 4ty..4ty+3 -> tile[j][ty], swizzled
#pragma unroll
        for (int c = 0; c < 4; c++) tile[4 * tx + c][ty ^ (tx & 15)] = b[c];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int jl = ty + 16 * r;                    // dst row
            const float4 v = tile[jl][tx ^ ((jl >> 2) & 15)];
            *(float4*)(dst + (j0 + jl) * rows + i0 + 4 * tx) = v;
        }
        __syncthreads();
    }
}

template <class K>
static float time_kernel(K launch, int reps = 20) {
    for (int i = 0; i < 3; i++) launch();
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) launch();
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main() {
    const size_t n = size_t(1) << 28, n4 = n / 4;
    float *a, *b;
    CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4));
    CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(b, 0, n * 4));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double bytes = 8.0 * n;
    {
        float ms = time_kernel([&] { cudaMemcpyAsync(b, a, n * 4, cudaMemcpyDeviceToDevice); });
        printf("cudaMemcpy D2D                          : %8.1f GB/s\n", bytes / ms / 1e6);
    }
#define RUN(U, H, T, CPS)                                                                                   \
    {                                                                                                       \
        size_t tiles = (n4 + size_t(T) * U - 1) / (size_t(T) * U);                                          \
        size_t g = CPS ? size_t(sms) * CPS : tiles;                                                         \
        if (g > tiles) g = tiles;                                                                           \
        float ms = time_kernel([&] { map_gs<U, H><<<(unsigned)g, T>>>(b, a, n4); });                        \
        printf("map U=%d hint=%d threads=%4d ctas/sm=%3d : %8.1f GB/s\n", U, H, T, CPS, bytes / ms / 1e6); \
    }
    RUN(4, 0, 256, 8) RUN(4, 0, 256, 4) RUN(4, 0, 256, 16) RUN(4, 0, 256, 0)
    RUN(8, 0, 256, 8) RUN(8, 0, 256, 4) RUN(8, 0, 256, 0)
    RUN(2, 0, 256, 8) RUN(2, 0, 256, 0) RUN(1, 0, 256, 0)
    RUN(4, 0, 512, 4) RUN(4, 0, 1024, 2) RUN(4, 0, 128, 16) RUN(8, 0, 128, 16)
    RUN(4, 1, 256, 8) RUN(4, 1, 256, 0) RUN(8, 1, 256, 8) RUN(8, 1, 256, 4)
    RUN(4, 2, 256, 8) RUN(8, 2, 256, 8) RUN(8, 2, 256, 4)
    RUN(16, 0, 128, 8) RUN(16, 1, 128, 8) RUN(16, 0, 64, 16)
    const size_t R = 16384, C = 16384;
#define RUNT(NAME, KERN, CPS)                                                                       \
    {                                                                                               \
        size_t g = size_t(sms) * CPS;                                                               \
        float ms = time_kernel([&] { KERN<<<(unsigned)g, 256>>>(b, a, R, C); });                    \
        printf("transpose %-22s ctas/sm=%2d : %8.1f GB/s\n", NAME, CPS, bytes / ms / 1e6);         \
    }
    RUNT("64x64 pad1", (transpose_t<64, 1>), 8) RUNT("64x64 pad1", (transpose_t<64, 1>), 4) RUNT("64x64 pad1", (transpose_t<64, 1>), 16)
    RUNT("32x32 pad1", (transpose_t<32, 1>), 8) RUNT("32x32 pad1", (transpose_t<32, 1>), 16)
    
    RUNT("64x64 reg4 swizzle", transpose_reg4, 8) RUNT("64x64 reg4 swizzle", transpose_reg4, 4) RUNT("64x64 reg4 swizzle", transpose_reg4, 12)
    {   // one tile per CTA (no stride loop)
        const size_t nt = (R / 64) * (C / 64);
        float ms = time_kernel([&] { transpose_t<64, 1><<<(unsigned)nt, 256>>>(b, a, R, C); });
        printf("transpose 64x64 pad1 one tile per CTA    : %8.1f GB/s\n", bytes / ms / 1e6);
        ms = time_kernel([&] { transpose_reg4<<<(unsigned)nt, 256>>>(b, a, R, C); });
        printf("transpose reg4 swizzle one tile per CTA  : %8.1f GB/s\n", bytes / ms / 1e6);
        for (int cps : {2, 3, 5, 6}) {
            ms = time_kernel([&] { transpose_t<64, 1><<<(unsigned)(sms * cps), 256>>>(b, a, R, C); });
            printf("transpose 64x64 pad1 ctas/sm=%d           : %8.1f GB/s\n", cps, bytes / ms / 1e6);
        }
    }
    // verify reg4 transpose on a pattern
    {
        std::vector<float> h(1024 * 1024);
        for (size_t i = 0; i < h.size(); i++) h[i] = float(i % 65521);
        CK(cudaMemcpy(a, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        transpose_reg4<<<148 * 8, 256>>>(b, a, 1024, 1024);
        std::vector<float> o(h.size());
        CK(cudaMemcpy(o.data(), b, o.size() * 4, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (size_t i = 0; i < 1024; i++)
            for (size_t j = 0; j < 1024; j++)
                if (o[j * 1024 + i] != h[i * 1024 + j]) bad++;
        printf("reg4 transpose mismatches: %zu\n", bad);
    }
    return 0;
}
