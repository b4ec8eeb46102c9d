// Developer microbenchmark (B200): shared-memory wavefront cost of LDS.128 / LDS.64 broadcast patterns and FFMA / FFMA2 issue rate.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_patterns lds_patterns.cu && ./lds_patterns
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ int chunk_of(int pat, int lane) {
    switch (pat) {
        case 0: return 0;            // all lanes one chunk
        case 1: return lane / 8;     // 4 distinct, quarter-uniform
        case 2: return lane % 8;     // 8 distinct, quarters identical
        case 3: return lane;         // 32 distinct
        case 4: return lane / 16;    // 2 distinct, half-uniform
        case 5: return lane % 4;     // 4 distinct inside a quarter
        case 6: return lane % 2;
        case 7: return lane / 4;     // 8 distinct, 4 adjacent lanes each
        case 8: return lane % 16;    // 16 distinct, halves identical
        case 9: return lane / 2;     // 16 distinct, lane pairs
        case 10: return (lane % 8) / 2 + 4 * (lane / 16);   // mixed
        case 11: return (lane / 8) * 8 + lane % 8;           // = lane (sanity)
        case 12: return (lane >> 1) & 7;                             // 8 distinct, lane pairs, halves identical
        case 13: return (lane & 1) | ((lane >> 4) << 1);            // 4 distinct: bit0 and bit4
        case 14: return 2 * ((lane & 1) | ((lane >> 4) << 1));      // same with a 32-byte stride
        case 15: return (lane >> 2) & 3;                             // 4 distinct, quads, halves identical
        case 16: return (lane & 3) >> 1 | ((lane >> 4) << 1);       // bit1 and bit4
        default: return 0;
    }
}

template <int VEC>
__global__ void lds_kernel(int pat, int iters, long long* out, float* sink) {
    __shared__ __align__(16) float sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const float* base = sm + chunk_of(pat, lane) * VEC;
    float acc = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float* p = base + ((j * 32 * VEC + it * 4) & 4095);
            const unsigned a = (unsigned)__cvta_generic_to_shared(p);
            if (VEC == 4) { float x, y, z, w; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a)); acc += (x + y) + (z + w); }
            else if (VEC == 2) { float x, y; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(a)); acc += x + y; }
            else { float x; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a)); acc += x; }
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>   // 0 = FFMA, 1 = FFMA2
__global__ void fma_kernel(int iters, long long* out, float* sink) {
    float2 a[8]; float2 b = make_float2(1.0001f, 0.9999f), c = make_float2(1e-3f, -1e-3f);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 1) a[i] = __ffma2_rn(a[i], b, c);
                else { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
            }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    long long* d_out; float* d_sink; cudaMalloc(&d_out, 8); cudaMalloc(&d_sink, 4096 * 4);
    long long h;
    const int iters = 2000;
    for (int warps : {8}) {
        printf("== %d warp(s) per SM (one CTA), cycles per warp-level LDS instruction (SM-wide cycles / instructions issued by ALL warps)\n", warps);
        for (int pat = 0; pat < 17; ++pat) {
            double r[3];
            lds_kernel<4><<<1, 32 * warps>>>(pat, iters, d_out, d_sink); cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost); r[0] = (double)h / (iters * 16.0 * warps);
            lds_kernel<2><<<1, 32 * warps>>>(pat, iters, d_out, d_sink); cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost); r[1] = (double)h / (iters * 16.0 * warps);
            lds_kernel<1><<<1, 32 * warps>>>(pat, iters, d_out, d_sink); cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost); r[2] = (double)h / (iters * 16.0 * warps);
            printf("pattern %2d: LDS.128 %.2f  LDS.64 %.2f  LDS.32 %.2f\n", pat, r[0], r[1], r[2]);
        }
    }
    for (int warps : {1, 2, 4, 8, 16}) {
        fma_kernel<0><<<1, 32 * warps>>>(iters, d_out, d_sink); cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
        double c0 = (double)h / (iters * 64.0);     // cycles per (2 FFMA = 64 lane-FMA per warp) pair
        fma_kernel<1><<<1, 32 * warps>>>(iters, d_out, d_sink); cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
        double c1 = (double)h / (iters * 32.0);     // cycles per FFMA2 (= 64 lane-FMA per warp)
        printf("%2d warps: cycles per 64 lane-FMAs of one warp: FFMA x2 = %.2f   FFMA2 = %.2f   => SM FMA/clk: %.1f vs %.1f\n", warps, c0, c1,
               64.0 * warps / c0, 64.0 * warps / c1);
    }
    cudaError_t e = cudaDeviceSynchronize(); printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
