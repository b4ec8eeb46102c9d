// Developer microbenchmark (B200, sm_100a): the hidden layer of the policy MLP -- D[64 units][32 envs] = W[64][64] . H[32][64]^T --
// on the 5th-generation tensor cores with 3xTF32 (hi/lo split, three tcgen05.mma passes into one TMEM accumulator), checked
// against a float64 reference and timed with clock64.  Groundwork for a tensor-core variant of rollout_kernel's layer 2
// (DESIGN.md section 7): it answers "how accurate" and "how many cycles per 32-env tile".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mlp_tc mlp_tc.cu && ./mlp_tc
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 64, N = 32, K = 64;          // units, envs, inputs
constexpr int KC = K / 4;                      // 16-byte chunks along K (4 tf32 each)
constexpr int LBO_A = M * 16;                  // bytes between K-chunks of A (all 64 rows of a chunk are contiguous)
constexpr int LBO_B = (N + 1) * 16;            // padded by one 16-byte row: conflict-free stores from the epilogue threads
constexpr int SBO = 128;                       // bytes between 8-row groups (a core matrix = 8 rows x 16 bytes, contiguous)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    // cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout_type=0 (no swizzle) [61,64)
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__device__ __forceinline__ float tf32_hi(float x) { uint32_t u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x)); return __uint_as_float(u); }

__global__ void __launch_bounds__(128) mlp_tc_kernel(const float* __restrict__ W, const float* __restrict__ H, float* __restrict__ D,
                                                     long long* __restrict__ cyc, int iters) {
    extern __shared__ __align__(128) float dyn_smem[];
    float (*sA)[KC * M * 4] = reinterpret_cast<float (*)[KC * M * 4]>(dyn_smem);   // hi, lo: [chunk][row][4]
    __shared__ __align__(128) float sB[2][KC * (N + 1) * 4];      // hi, lo: [chunk][row (padded)][4]
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < M * K; i += 128) {                      // W[u][k] (nn.Linear layout) -> K-major canonical layout
        const int u = i / K, k = i % K; const float x = W[i], hi = tf32_hi(x);
        sA[0][((k >> 2) * M + u) * 4 + (k & 3)] = hi; sA[1][((k >> 2) * M + u) * 4 + (k & 3)] = x - hi;
    }
    for (int i = tid; i < N * K; i += 128) {
        const int e = i / K, k = i % K; const float x = H[i], hi = tf32_hi(x);
        sB[0][((k >> 2) * (N + 1) + e) * 4 + (k & 3)] = hi; sB[1][((k >> 2) * (N + 1) + e) * 4 + (k & 3)] = x - hi;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {                                              // one warp allocates 32 TMEM columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=TF32 [7,10)=2, B=TF32 [10,13)=2, K-major both, N>>3 [17,23), M>>4 [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t a_hi = smem_u32(sA[0]), a_lo = smem_u32(sA[1]), b_hi = smem_u32(sB[0]), b_lo = smem_u32(sB[1]);
    long long t_mma = 0, t_ld = 0;
    float v[32];
    for (int it = 0; it < iters; ++it) {
        const long long c0 = clock64();
        if (tid == 0) {
            int first = 1;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t ab = pass == 2 ? a_lo : a_hi, bb = pass == 1 ? b_lo : b_hi;
#pragma unroll
                for (int ks = 0; ks < K / 8; ++ks) {              // one MMA = K 8 = two 16-byte chunks
                    const uint64_t da = make_desc(ab + ks * 2 * LBO_A, LBO_A, SBO), db = make_desc(bb + ks * 2 * LBO_B, LBO_B, SBO);
                    const uint32_t acc = first ? 0u : 1u; first = 0;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
        }
        {   // everybody waits for the accumulator (bounded spin: a wrong descriptor must not hang the GPU)
            const uint32_t parity = it & 1;
            uint32_t done = 0; int spins = 0;
            while (!done && spins < (1 << 22)) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&s_bar)), "r"(parity) : "memory");
                ++spins;
            }
            if (!done) { if (tid == 0) cyc[2] = it + 1; break; }
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const long long c1 = clock64();
        // M = 64 accumulators live in lanes 0-15 of each 32-lane quadrant: row = 16 * quadrant + lane
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        uint32_t r[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                       "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                       "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                          // TMEM reads done before the next iteration overwrites D
        const long long c2 = clock64();
        t_mma += c1 - c0; t_ld += c2 - c1;
    }
    if (lane < 16)
        for (int j = 0; j < 32; ++j) D[(warp * 16 + lane) * N + j] = v[j];
    if (tid == 0) { cyc[0] = t_mma; cyc[1] = t_ld; }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

int main() {
    std::vector<float> W(M * K), H(N * K), D(M * N);
    srand(1);
    for (auto& x : W) x = (rand() / (float)RAND_MAX - 0.5f) * 0.25f;      // nn.Linear-like scale 1/sqrt(64)
    for (auto& x : H) x = (rand() / (float)RAND_MAX - 0.5f) * 2.0f;       // tanh outputs
    float *dW, *dH, *dD; long long* dC;
    cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dH, H.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dC, 32); cudaMemset(dC, 0, 32);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dH, H.data(), H.size() * 4, cudaMemcpyHostToDevice);
    const int iters = 2000;
    const int dyn = 2 * KC * M * 4 * (int)sizeof(float);
    cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    mlp_tc_kernel<<<1, 128, dyn>>>(dW, dH, dD, dC, iters);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    long long c[3]; cudaMemcpy(c, dC, 24, cudaMemcpyDeviceToHost);
    if (c[2]) printf("TIMEOUT waiting for the MMA commit at iteration %lld\n", c[2] - 1); cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double max_abs = 0, max_ref = 0, max_f32 = 0;
    for (int u = 0; u < M; ++u)
        for (int n = 0; n < N; ++n) {
            double ref = 0; float f = 0.f;
            for (int k = 0; k < K; ++k) { ref += (double)W[u * K + k] * (double)H[n * K + k]; f = fmaf(W[u * K + k], H[n * K + k], f); }
            max_abs = fmax(max_abs, fabs(D[u * N + n] - ref)); max_ref = fmax(max_ref, fabs(ref)); max_f32 = fmax(max_f32, fabs((double)f - ref));
        }
    printf("3xTF32 tcgen05: max |err| = %.3e (fp32 FMA chain: %.3e), max |ref| = %.3f\n", max_abs, max_f32, max_ref);
    printf("cycles per 64x32x64 tile: 24 MMAs + commit + wait = %.1f, tcgen05.ld 32x32b.x32 + barrier = %.1f\n", (double)c[0] / iters, (double)c[1] / iters);
    return 0;
}
