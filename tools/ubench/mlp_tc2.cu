// Developer microbenchmark #2 (B200, sm_100a): how fast can the tensor cores run the policy MLP's GEMMs with 3xTF32 when the MMAs are
// organised as SEVERAL INDEPENDENT accumulator chains (actor / critic, the three hi-lo passes, a K split) instead of one dependent
// chain?  Round 1 measured 56 cycles per dependent tcgen05.mma (M64 N32 K8): latency, not throughput.  This program times arbitrary
// "schedules" (lists of MMAs: operand pass, k-step, accumulator column block) for the tile shapes the rollout kernel would use and
// checks the summed accumulators against a float64 reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mlp_tc2 mlp_tc2.cu && ./mlp_tc2
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__device__ __forceinline__ float tf32_hi(float x) { uint32_t u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x)); return __uint_as_float(u); }

struct Op { int a_pass, b_pass, ks, dcol, acc; };   // a_pass / b_pass: 0 = hi, 1 = lo; dcol: accumulator column block (units of N columns)
#define MAX_OPS 192
__constant__ Op c_ops[MAX_OPS];

// D[M][N] (+)= A[M][K] . B[N][K]^T, A and B K-major canonical no-swizzle layouts, hi/lo copies of each.
// a_from_tmem: the LO copy of A is read from tensor memory instead of shared memory (M = 128 only).
template <int M, int N, int K, int NETS, int MODE>
__global__ void __launch_bounds__(128) tc_kernel(const float* __restrict__ W, const float* __restrict__ H, float* __restrict__ D, int n_ops, int n_dblk,
                                                 long long* __restrict__ cyc, int iters, int a_from_tmem) {
    constexpr int KC = K / 4;
    constexpr int LBO_A = M * 16, LBO_B = (N + 1) * 16, SBO = 128;
    extern __shared__ __align__(128) float dyn[];
    float* sA0 = dyn; float* sA1 = dyn + KC * M * 4; float* sB0 = sA1 + KC * M * 4; float* sB1 = sB0 + KC * (N + 1) * 4;
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < M * K; i += 128) {
        const int u = i / K, k = i % K; const float x = W[i], hi = tf32_hi(x);
        sA0[((k >> 2) * M + u) * 4 + (k & 3)] = hi; sA1[((k >> 2) * M + u) * 4 + (k & 3)] = x - hi;
    }
    for (int i = tid; i < N * K; i += 128) {
        const int e = i / K, k = i % K; const float x = H[i], hi = tf32_hi(x);
        sB0[((k >> 2) * (N + 1) + e) * 4 + (k & 3)] = hi; sB1[((k >> 2) * (N + 1) + e) * 4 + (k & 3)] = x - hi;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    constexpr int TCOLS = 512;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t tmem_alo = tmem + 256;            // columns 256.. hold A_lo[M][K] when a_from_tmem (K <= 128)
    if (a_from_tmem && M == 128) {
        // thread = row (lane of TMEM), writes its K lo values, 8 columns at a time
        const int row = tid;
        for (int k0 = 0; k0 < K; k0 += 8) {
            uint32_t r[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float x = W[row * K + k0 + j]; r[j] = __float_as_uint(x - tf32_hi(x)); }
            const uint32_t ta = tmem_alo + ((uint32_t)(warp * 32) << 16) + k0;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                         "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t a_base[2] = {smem_u32(sA0), smem_u32(sA1)}, b_base[2] = {smem_u32(sB0), smem_u32(sB1)};
    long long t_mma = 0, t_ld = 0;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
    for (int it = 0; it < iters; ++it) {
        const long long c0 = clock64();
        if (tid == 0) {
            // compile-time schedule: base descriptors in (uniform) registers, constant adds per MMA
            const uint64_t dA[2] = {make_desc(a_base[0], LBO_A, SBO), make_desc(a_base[1], LBO_A, SBO)};
            const uint64_t dB[2] = {make_desc(b_base[0], LBO_B, SBO), make_desc(b_base[1], LBO_B, SBO)};
            auto mma = [&](int ap, int bp, int ks, int dblk, uint32_t acc) {
                const uint64_t db = dB[bp] + (uint64_t)((ks * 2 * LBO_B) >> 4);
                const uint32_t d = tmem + dblk * N;
                if (a_from_tmem && ap == 1) {
                    const uint32_t ta = tmem_alo + ks * 8;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(d), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
                } else {
                    const uint64_t da = dA[ap] + (uint64_t)((ks * 2 * LBO_A) >> 4);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                }
            };
            constexpr int KS = K / 8;
            if (MODE == 0) {
#pragma unroll
                for (int p = 0; p < 3; ++p)
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                        for (int net = 0; net < NETS; ++net) mma(p == 2, p == 1, ks, net, !(p == 0 && ks == 0));
            } else if (MODE == 1) {
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                    for (int net = 0; net < NETS; ++net)
#pragma unroll
                        for (int p = 0; p < 3; ++p) mma(p == 2, p == 1, ks, net * 3 + p, ks > 0);
            } else {
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                    for (int net = 0; net < NETS; ++net)
#pragma unroll
                        for (int p = 0; p < 3; ++p) mma(p == 2, p == 1, ks, (net * 3 + p) * 2 + (ks & 1), ks > 1);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
        }
        {
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
        // sum the accumulator blocks: block b at columns [b*N, (b+1)*N)
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
        for (int b = 0; b < n_dblk; ++b) {
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + b * N;
            uint32_t r[32];
            if (N == 32) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                               "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                               "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                             : "r"(taddr));
            } else {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                               "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                             : "r"(taddr));
#pragma unroll
                for (int j = 16; j < 32; ++j) r[j] = 0;
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(r[j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        const long long c2 = clock64();
        t_mma += c1 - c0; t_ld += c2 - c1;
    }
    // M = 128: row = 32 * warp + lane.  M = 64: rows live in lanes 0-15 of each quadrant, row = 16 * warp + lane
    if (M == 128) { for (int j = 0; j < N; ++j) D[(warp * 32 + lane) * N + j] = v[j]; }
    else if (lane < 16) { for (int j = 0; j < N; ++j) D[(warp * 16 + lane) * N + j] = v[j]; }
    if (tid == 0) { cyc[0] = t_mma; cyc[1] = t_ld; }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS) : "memory");
}

// schedule builders: 3xTF32 = (hi,hi) + (hi,lo) + (lo,hi).  mode: 0 = one chain, 1 = chain per pass (3), 2 = pass x K-split-2 (6)
static std::vector<Op> build(int K, int nets, int mode) {
    const int ksteps = K / 8;
    const int passes[3][2] = {{0, 0}, {0, 1}, {1, 0}};
    std::vector<std::vector<Op>> chains;
    for (int n = 0; n < nets; ++n) {
        if (mode == 0) {
            std::vector<Op> c;
            for (int p = 0; p < 3; ++p) for (int ks = 0; ks < ksteps; ++ks) c.push_back({passes[p][0], passes[p][1], ks, 0, 1});
            chains.push_back(c);
        } else if (mode == 1) {
            for (int p = 0; p < 3; ++p) { std::vector<Op> c; for (int ks = 0; ks < ksteps; ++ks) c.push_back({passes[p][0], passes[p][1], ks, 0, 1}); chains.push_back(c); }
        } else {
            for (int p = 0; p < 3; ++p)
                for (int h = 0; h < 2; ++h) { std::vector<Op> c; for (int ks = h; ks < ksteps; ks += 2) c.push_back({passes[p][0], passes[p][1], ks, 0, 1}); chains.push_back(c); }
        }
    }
    for (size_t c = 0; c < chains.size(); ++c) { for (auto& o : chains[c]) o.dcol = (int)c; chains[c][0].acc = 0; }
    std::vector<Op> out;            // round-robin over the chains
    for (size_t i = 0;; ++i) {
        bool any = false;
        for (auto& c : chains) if (i < c.size()) { out.push_back(c[i]); any = true; }
        if (!any) break;
    }
    return out;
}

template <int M, int N, int K, int NETS, int MODE>
static void run(const char* what, int a_from_tmem) {
    const int nets = NETS, mode = MODE;
    std::vector<float> W(M * K), H(N * K), D(M * N);
    srand(1);
    for (auto& x : W) x = (rand() / (float)RAND_MAX - 0.5f) * 0.25f;
    for (auto& x : H) x = (rand() / (float)RAND_MAX - 0.5f) * 2.0f;
    float *dW, *dH, *dD; long long* dC;
    cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dH, H.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dC, 32); cudaMemset(dC, 0, 32);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dH, H.data(), H.size() * 4, cudaMemcpyHostToDevice);
    std::vector<Op> ops = build(K, nets, mode);
    int nblk = 0; for (auto& o : ops) nblk = o.dcol + 1 > nblk ? o.dcol + 1 : nblk;
    if (nblk * N > 256) { printf("%-44s skipped (schedule too large)\n", what); return; }
    const int iters = 1000;
    const int dyn = (2 * (K / 4) * M * 4 + 2 * (K / 4) * (N + 1) * 4) * (int)sizeof(float);
    cudaFuncSetAttribute(tc_kernel<M, N, K, NETS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    tc_kernel<M, N, K, NETS, MODE><<<1, 128, dyn>>>(dW, dH, dD, (int)ops.size(), nblk, dC, iters, a_from_tmem);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s CUDA error: %s\n", what, cudaGetErrorString(e)); exit(1); }
    long long c[3]; cudaMemcpy(c, dC, 24, cudaMemcpyDeviceToHost);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double max_abs = 0;
    for (int u = 0; u < M; ++u)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)W[u * K + k] * (double)H[n * K + k];
            max_abs = fmax(max_abs, fabs(D[u * N + n] - nets * ref));     // every net computes the same product here
        }
    printf("%-44s M%3d N%2d K%3d: %3d MMAs in %2d chains: %7.1f cycles (%.1f / MMA), ld+sum %6.1f, max|err| %.2e%s\n", what, M, N, K, (int)ops.size(), nblk,
           (double)c[0] / iters, (double)c[0] / iters / ops.size(), (double)c[1] / iters, max_abs, c[2] ? "  TIMEOUT" : "");
    cudaFree(dW); cudaFree(dH); cudaFree(dD); cudaFree(dC);
}

int main() {
    // hidden 64, layer 2: one net = 24 MMAs
    run<64, 32, 64, 1, 0>("h64 L2, 1 net, 1 chain (round-1 baseline)", 0);
    run<64, 32, 64, 2, 0>("h64 L2, 2 nets, chain per net", 0);
    run<64, 32, 64, 2, 1>("h64 L2, 2 nets, chain per pass (6)", 0);
    run<64, 32, 64, 1, 2>("h64 L2, 1 net, pass x K-split (6)", 0);
    run<64, 16, 64, 2, 1>("h64 L2, N16, 2 nets, chain per pass (6)", 0);
    // layer 1 (obs padded to K = 16) as ONE M = 128 tile (actor rows 0-63, critic rows 64-127)
    run<128, 32, 16, 1, 0>("h64 L1 fused M128, 1 chain", 0);
    run<128, 32, 16, 1, 1>("h64 L1 fused M128, chain per pass (3)", 0);
    run<128, 32, 32, 1, 1>("L1 obs 24 -> K32, M128, chain per pass (3)", 0);
    // hidden 128, layer 2: one net = 48 MMAs of M128
    run<128, 32, 128, 1, 0>("h128 L2, 1 net, 1 chain", 0);
    run<128, 32, 128, 1, 1>("h128 L2, 1 net, chain per pass (3)", 0);
    run<128, 32, 128, 1, 2>("h128 L2, 1 net, pass x K-split (6)", 0);
    run<128, 32, 128, 2, 1>("h128 L2, 2 nets, chain per pass (6)", 0);
    run<128, 32, 128, 1, 1>("h128 L2, 1 net, pass chains, A_lo in TMEM", 1);
    run<128, 32, 128, 2, 1>("h128 L2, 2 nets, pass chains, A_lo in TMEM", 1);
    return 0;
}
