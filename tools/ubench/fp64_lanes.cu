// Micro-benchmark (B200): issue cost of double-precision FMAs for ONE warp as a function of the number of active lanes and of the
// instruction-level parallelism.  Question behind it: the collector's owner warp (32 envs, one per lane) is FP64-throughput bound at ~4 cycles
// per warp instruction; would four owner warps with 8 active lanes each (on the four SM sub-partitions) issue faster?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lanes fp64_lanes.cu && ./fp64_lanes
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double* out, long long* cyc, int active, int iters, double a, double b) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a + i + lane;
    __syncthreads();
    long long t0 = 0, t1 = 0;
    if (lane < active) {
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], b, a);
        }
        t1 = clock64();
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[threadIdx.x] = s;
    if (lane == 0) cyc[warp] = t1 - t0;
}

template <int ILP> void run(int warps, int active) {
    double* d; long long* c; cudaMalloc(&d, 1024 * 8); cudaMalloc(&c, 32 * 8);
    const int iters = 4096;
    k<ILP><<<1, warps * 32>>>(d, c, active, iters, 1.0000001, 0.9999999);
    k<ILP><<<1, warps * 32>>>(d, c, active, iters, 1.0000001, 0.9999999);
    long long h[32]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
    double mx = 0; for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
    printf("warps %d  active lanes %2d  ILP %d : %.2f cycles per DFMA warp-instruction (slowest warp)\n", warps, active, ILP, mx / ((double)iters * ILP));
    cudaFree(d); cudaFree(c);
}

int main() {
    for (int active : {32, 16, 8, 4}) { run<1>(1, active); run<8>(1, active); }
    for (int active : {32, 16, 8}) { run<8>(4, active); run<8>(8, active); }
    return 0;
}
