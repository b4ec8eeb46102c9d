// scg_rollout_tc.cuh -- tensor-core variant of the fused rollout collector (sm_100a: tcgen05.mma + TMEM).
//
// Same collector as rollout_kernel (scg_kernels.cuh; PPO.train_step collector loop, safe_control_gym/controllers/ppo/ppo.py:259-297;
// MLPActorCritic.step ppo_utils.py:224-231), same env-step schedule (owner warp, helper warps, record helpers): what changes is the
// policy forward.  The two hidden layers of actor and critic run on the 5th-generation tensor cores as 3xTF32 products
// (x = hi + lo, D = Whi.Hhi + Whi.Hlo + Wlo.Hhi, FP32 accumulation in tensor memory: max |err| ~6e-7 against float64, measured
// with tools/ubench/mlp_tc2.cu) -- accurate enough for the 1e-5 policy parity, which bf16 / plain TF32 are not:
//
//   layer 1  D1[128 x 32] = [W1_actor; W1_critic][128 x OBSP] . obs^T            one M128 N32 tile (hidden 64), K = obs_dim padded to 8
//   layer 2  D2_net[64 x 32] = W2_net[64 x 64] . H1_net^T                          two M64 N32 tiles
//   every (net, hi/lo pass) is its OWN accumulator (independent MMA chains keep the tensor pipe busy; the three passes are summed
//   in FP32 by the epilogue), A operands (weights, hi and lo copies) are resident in shared memory in the K-major canonical
//   no-swizzle layout, B operands (observation tile / first hidden activations of the CTA's 32 envs) are written by the CUDA cores
//   in the same layout, one elected thread issues the MMAs, `tcgen05.commit` -> mbarrier hands the accumulators to the epilogue:
//   all 8 warps read TMEM with `tcgen05.ld`, add the bias, apply tanh (packed FFMA2 rational), and either write the hi/lo split of
//   the activations as the next B operand (layer 1) or contract with the output rows (layer 3, fused: 16-lane reduce-scatter with
//   shuffles, per-warp partial sums in shared memory exactly like the FFMA2 kernel).
//
// The weights arrive in the SAME packed blob as for rollout_kernel (one TMA bulk copy); each CTA converts them once per launch.
#pragma once
#include <cuda_fp16.h>
#include <type_traits>
#include "scg_kernels.cuh"

#ifdef SCG_TC_PROF
static __device__ long long tc_ticks[16];
#define TC_TICK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long c_ = clock64(); tc_ticks[i] += c_ - tk_; tk_ = c_; } } while (0)
#define TC_TICK0() long long tk_ = clock64()
#define TC_OWN(i, a, b) do { if (threadIdx.x == tc::NT - 32 && blockIdx.x == 0) tc_ticks[i] += (b) - (a); } while (0)
#else
#define TC_TICK(i) do { } while (0)
#define TC_TICK0() do { } while (0)
#define TC_OWN(i, a, b) do { } while (0)
#endif

namespace tc {

constexpr int E = 32;                    // envs per CTA = N of every MMA
constexpr int NT = 256;
constexpr int LBO_B = (E + 1) * 16;      // bytes between K-chunks (4 tf32) of a B tile: one 16-byte row of padding => conflict-free epilogue stores
constexpr int SBO = 128;                 // bytes between 8-row groups (core matrix = 8 rows x 16 bytes, contiguous)
constexpr int BCH = LBO_B / 4;           // floats per K-chunk of a B tile

__host__ __device__ constexpr int pad8(int x) { return (x + 7) & ~7; }
// float index of element (row n, k) in a B tile / of element (row m, k) in an A tile with M rows
__device__ __forceinline__ int b_idx(int k, int n) { return (k >> 2) * BCH + n * 4 + (k & 3); }
__host__ __device__ constexpr int b_tile_floats(int K) { return (K / 4) * BCH; }
__device__ __forceinline__ int a_idx(int m, int k, int M) { return ((k >> 2) * M + m) * 4 + (k & 3); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    // cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout_type=0 (no swizzle) [61,64)
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=TF32 [7,10)=2, B=TF32 [10,13)=2, K-major both, N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ float tf32_hi(float x) { uint32_t u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x)); return __uint_as_float(u); }

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// bounded wait: a wrong descriptor must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0; int spins = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && ++spins > (1 << 24)) asm volatile("trap;");
    }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int TU> __device__ __forceinline__ void act16(float (&x)[16], int act) {
    if (act == SCG_ACT_TANH) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) tanh_fast4(x[j], x[j + 1], x[j + 2], x[j + 3]);
    } else if (act == SCG_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = fmaxf(x[j], 0.0f);
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = x[j] > 0.0f ? x[j] : 0.01f * x[j];
    }
}

// observation sink of the tensor-core kernel: element (k, env) of a B tile as its hi / lo pair
struct TcObsSink {
    float* hi; float* lo; int e;
    template <typename R> __device__ __forceinline__ void operator()(int k, R v) const {
        const float x = (float)v, h = tf32_hi(x);
        const int i = b_idx(k, e);
        hi[i] = h; lo[i] = x - h;
    }
};

// shared-memory plan (float offsets; every section a multiple of 4 floats = 16 bytes)
template <int HID> struct Smem {
    int blob, small, w1, w2, h1, obs, tobs, cobs, out, xg, total_floats;
    int OBSP, w1_floats, w2_floats, h1_floats, obs_floats, small_floats, slot_floats, tiles_floats;
    __host__ __device__ Smem(int obs_dim, int act_dim, int xg_floats) {
        const BlobLayout L = blob_layout(obs_dim, act_dim, HID);
        OBSP = pad8(obs_dim);
        w1_floats = OBSP * 2 * HID;                  // one copy (hi or lo) of the fused layer-1 A tile [2H rows][OBSP]
        w2_floats = HID * HID;                       // one copy of one net's layer-2 A tile
        h1_floats = b_tile_floats(HID);              // one copy of one net's H1 B tile
        obs_floats = b_tile_floats(OBSP);            // one copy of an observation B tile
        small_floats = pad4i(2 * HID) * 2 + pad4i((act_dim + 1) * HID) + pad4i(act_dim + 1) + pad4i(act_dim);   // b1, b2, W3, b3, logstd
        // the staged blob is dead once converted: the activation tiles alias it
        const int act_tiles = 4 * h1_floats + 6 * obs_floats;       // H1 {net 0,1} x {hi,lo}; obs {cur,next,terminal} x {hi,lo}
        blob = 0;
        const int region0 = L.total > act_tiles ? L.total : act_tiles;
        h1 = 0; obs = 4 * h1_floats; tobs = obs + 4 * obs_floats; slot_floats = 2 * obs_floats; tiles_floats = act_tiles;
        small = pad4i(region0);
        w1 = small + pad4i(small_floats);
        w2 = w1 + 2 * w1_floats;
        cobs = w2 + 4 * w2_floats;
        out = cobs + 2 * obs_dim * E;
        xg = out + pad4i(4 * (act_dim + 1) * E);
        total_floats = xg + pad4i(xg_floats);
    }
};
struct SmallLayout { int b1, b2, w3, b3, logstd; };
__host__ __device__ inline SmallLayout small_layout(int act_dim, int hid) {
    SmallLayout s; s.b1 = 0; s.b2 = pad4i(2 * hid); s.w3 = 2 * pad4i(2 * hid); s.b3 = s.w3 + pad4i((act_dim + 1) * hid); s.logstd = s.b3 + pad4i(act_dim + 1);
    return s;
}

// TMEM columns: D1 (three pass accumulators of the fused layer-1 tile) and D2 (net x pass)
constexpr int D1_COL = 0;          // + pass * 32
constexpr int D2_COL = 128;        // + (net * 3 + pass) * 32
constexpr int TMEM_COLS = 512;

struct Ctx {
    uint32_t tmem;
    uint32_t w1_s[2], w2_s[2][2], h1_s[2][2];   // shared-memory byte addresses: [hi/lo], [net][hi/lo]
    float* h1_p[2][2];
    const float* small; SmallLayout SL;
    float* s_out;
    uint64_t* bar;                               // two mbarriers
    int obsp;
};

// One forward of actor + critic for the CTA's 32 envs on the tensor cores.  obs_s: shared byte address of the observation B tile
// (hi copy; the lo copy follows at +obs_lo_bytes).  `fwd` counts forwards (mbarrier parity).  Ends with a __syncthreads(): the
// per-warp partial sums of the output layer are then readable with mlp_out<E, 4, A>(ctx.s_out, ...).
template <int HID, int A>
__device__ __noinline__ void forward(const Ctx& c /* in shared memory */, uint32_t obs_s, uint32_t obs_lo_bytes, int act, uint32_t& fwd) {
    static_assert(HID == 64, "3xTF32 forward of the hidden-64 collector (hidden 128: forward128)");
    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    const uint32_t parity = fwd & 1u;
    TC_TICK0();
    // ---- layer 1: one elected thread issues the MMAs (3 independent pass chains, K = obsp)
    if (w == 0) {
      if (l == 0) {
        tc_fence_after();
        constexpr uint32_t I1 = idesc_tf32(128, E);
        constexpr uint32_t LBO_A1 = 128 * 16;
        const int ksteps = c.obsp >> 3;
        const uint32_t tm = c.tmem;
        // base descriptors once; per MMA only the start-address field (bytes >> 4) advances by two K-chunks
        const uint64_t dA[2] = {make_desc(c.w1_s[0], LBO_A1, SBO), make_desc(c.w1_s[1], LBO_A1, SBO)};
        const uint64_t dB[2] = {make_desc(obs_s, LBO_B, SBO), make_desc(obs_s + obs_lo_bytes, LBO_B, SBO)};
#pragma unroll 1
        for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t ka = (uint64_t)((ks * 2 * LBO_A1) >> 4), kb = (uint64_t)((ks * 2 * LBO_B) >> 4);
#pragma unroll
            for (int p = 0; p < 3; ++p)        // (hi,hi) (hi,lo) (lo,hi)
                mma_tf32(tm + D1_COL + p * 32, dA[p == 2] + ka, dB[p == 1] + kb, I1, ks > 0);
        }
        mma_commit(c.bar);
      }
      __syncwarp();       // the other lanes of the issuing warp wait here instead of competing with the elected lane for issue slots
    }
    TC_TICK(0);
    mbar_wait_bounded(c.bar, parity);
    tc_fence_after();
    TC_TICK(1);
    {   // ---- epilogue 1: row u = 32 (w & 3) + l of the fused tile (actor units 0..63, critic 64..127), envs [16 (w >> 2), +16)
        const int u = 32 * (w & 3) + l, h = w >> 2, net = u >> 6, unit = u & 63;
        const uint32_t ta = c.tmem + ((uint32_t)(32 * (w & 3)) << 16) + D1_COL + 16 * h;
        float x[16], y[16], z[16];
        tmem_ld16(ta, x); tmem_ld16(ta + 32, y); tmem_ld16(ta + 64, z);
        tmem_ld_wait();
        const float b = c.small[c.SL.b1 + u];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = (x[j] + y[j]) + (z[j] + b);
        act16<0>(x, act);
        float* hi = c.h1_p[net][0]; float* lo = c.h1_p[net][1];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float hv = tf32_hi(x[j]);
            const int i = b_idx(unit, 16 * h + j);
            hi[i] = hv; lo[i] = x[j] - hv;
        }
    }
    TC_TICK(2);
    fence_proxy_async();          // the activation tile was written through the generic proxy, the tensor core reads it through the async proxy
    tc_fence_before();
    __syncthreads();
    TC_TICK(3);
    // ---- layer 2: 2 nets x 3 passes = 6 independent chains, K = 64
    if (w == 0) {
      if (l == 0) {
        tc_fence_after();
        constexpr uint32_t I2 = idesc_tf32(HID, E);
        constexpr uint32_t LBO_A2 = HID * 16;
        const uint32_t tm = c.tmem;
        uint64_t dA[2][2], dB[2][2];
#pragma unroll
        for (int net = 0; net < 2; ++net)
#pragma unroll
            for (int p = 0; p < 2; ++p) { dA[net][p] = make_desc(c.w2_s[net][p], LBO_A2, SBO); dB[net][p] = make_desc(c.h1_s[net][p], LBO_B, SBO); }
#pragma unroll
        for (int ks = 0; ks < HID / 8; ++ks) {
            constexpr uint64_t one = 1;
            const uint64_t ka = one * ((ks * 2 * LBO_A2) >> 4), kb = one * ((ks * 2 * LBO_B) >> 4);
#pragma unroll
            for (int net = 0; net < 2; ++net)
#pragma unroll
                for (int p = 0; p < 3; ++p)
                    mma_tf32(tm + D2_COL + (net * 3 + p) * 32, dA[net][p == 2] + ka, dB[net][p == 1] + kb, I2, ks > 0);
        }
        mma_commit(c.bar + 1);
      }
      __syncwarp();
    }
    TC_TICK(4);
    mbar_wait_bounded(c.bar + 1, parity);
    tc_fence_after();
    TC_TICK(5);
    {   // ---- epilogue 2 (+ fused output layer): warps 0-3 actor, 4-7 critic; M = 64 accumulators live in lanes 0-15 of each TMEM
        //      quadrant (row = 16 q + lane): lane l < 16 takes envs 0..15 of unit 16 q + l, lane l + 16 envs 16..31 of the same unit
        const int net = w >> 2, q = w & 3, unit = 16 * q + (l & 15);
        const uint32_t ta = c.tmem + ((uint32_t)(32 * q) << 16) + D2_COL + net * 96;
        float x[16], y[16], z[16], lo16[16];
        tmem_ld16(ta, x); tmem_ld16(ta + 32, y); tmem_ld16(ta + 64, z);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) lo16[j] = (x[j] + y[j]) + z[j];
        tmem_ld16(ta + 16, x); tmem_ld16(ta + 48, y); tmem_ld16(ta + 80, z);
        tmem_ld_wait();
        const float b = c.small[c.SL.b2 + net * HID + unit];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float up = __shfl_sync(0xffffffffu, (x[j] + y[j]) + z[j], l & 15);     // the upper half-warp fetches envs 16..31 from lane l - 16
            x[j] = ((l < 16) ? lo16[j] : up) + b;
        }
        act16<0>(x, act);
        // output rows of this net: reduce-scatter over the 16 units of the quadrant (lane bits 3..0); lane l ends with env l
        const int n_out = net ? 1 : A, row0 = net ? A : 0;
#pragma unroll
        for (int o = 0; o < A; ++o) {
            if (o < n_out) {                   // warp-uniform
                const float w3 = c.small[c.SL.w3 + (row0 + o) * HID + unit];
                float p8[8], p4[4], p2[2];
                {
                    const bool up = (l & 8) != 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float a0 = x[j] * w3, a1 = x[j + 8] * w3;
                        const float r = __shfl_xor_sync(0xffffffffu, up ? a0 : a1, 8);
                        p8[j] = (up ? a1 : a0) + r;
                    }
                }
                {
                    const bool up = (l & 4) != 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const float r = __shfl_xor_sync(0xffffffffu, up ? p8[j] : p8[j + 4], 4); p4[j] = (up ? p8[j + 4] : p8[j]) + r; }
                }
                {
                    const bool up = (l & 2) != 0;
#pragma unroll
                    for (int j = 0; j < 2; ++j) { const float r = __shfl_xor_sync(0xffffffffu, up ? p4[j] : p4[j + 2], 2); p2[j] = (up ? p4[j + 2] : p4[j]) + r; }
                }
                const bool up = (l & 1) != 0;
                const float r = __shfl_xor_sync(0xffffffffu, up ? p2[0] : p2[1], 1);
                c.s_out[(q * (A + 1) + row0 + o) * E + l] = (up ? p2[1] : p2[0]) + r;
            }
        }
    }
    TC_TICK(6);
    tc_fence_before();
    __syncthreads();
    TC_TICK(7);
    fwd += 1;
}

// =============================================================================================
// hidden 128 (the width of every quadrotor checkpoint the reference ships): FP16 operands with a SCALED low part.
//   x = hi + lo' * 2^-11,  hi = fp16(x), lo' = fp16((x - hi) * 2^11)   (22 significant bits, like the TF32 hi/lo pair, and lo' stays
//   a normal fp16 number); D = acc0 + acc1 * 2^-11 with acc0 = Whi.Hhi and acc1 = Whi.Hlo' + Wlo'.Hhi in two tensor-memory accumulators.
//   `kind::f16` MMAs contract K = 16 per instruction -- half the instructions of the TF32 form (the tensor pipe needs ~65 cycles
//   per tcgen05.mma at N = 32 whatever its size, tools/ubench/mlp_tc2.cu) -- and both nets' weights (2 x 128 x 128, hi and lo') fit
//   in shared memory as 128 KB, which the TF32 copies (256 KB) would not.
//   layer 1  D1_net[128 x 32] = W1_net[128 x OBSP16] . obs^T     layer 2  D2_net[128 x 32] = W2_net[128 x 128] . H1_net^T    (M128 N32 K16)
// =============================================================================================
constexpr float LO_SCALE = 2048.0f, LO_UNSCALE = 1.0f / 2048.0f;
__host__ __device__ constexpr int pad16(int x) { return (x + 15) & ~15; }
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {     // D=F32 [4,6)=1, A=F16 [7,10)=0, B=F16 [10,13)=0, K-major both
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// half index of element (row, k) in a B tile (33 rows of 16 bytes per K-chunk of 8 halves) / in an A tile with M rows
__device__ __forceinline__ int bh_idx(int k, int n) { return (k >> 3) * (LBO_B / 2) + n * 8 + (k & 7); }
__device__ __forceinline__ int ah_idx(int m, int k, int M) { return ((k >> 3) * M + m) * 8 + (k & 7); }
__host__ __device__ constexpr int bh_tile_floats(int K) { return (K / 8) * (LBO_B / 4); }     // size of one fp16 B tile in float units
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn((x - __half2float(hi)) * LO_SCALE);
}

struct Smem128 {
    int small, w1, w2, h1, obs, tobs, cobs, out, total_floats;
    int OBSP, w1_floats, w2_floats, h1_floats, obs_tile_floats, slot_floats, small_floats, tiles_floats;
    __host__ __device__ Smem128(int obs_dim, int act_dim, int /*xg_floats: X_GOAL stays in global memory*/ = 0) {
        constexpr int HID = 128;
        OBSP = pad16(obs_dim);
        w1_floats = 128 * OBSP / 2;                  // one fp16 copy of one net's layer-1 A tile [128 rows][OBSP]
        w2_floats = HID * HID / 2;
        h1_floats = bh_tile_floats(HID);
        obs_tile_floats = bh_tile_floats(OBSP);
        slot_floats = pad4i(obs_dim * E) + 2 * obs_tile_floats;       // plain float32 copy [k][E] + hi + lo' operand tiles
        small_floats = pad4i(2 * HID) * 2 + pad4i((act_dim + 1) * HID) + pad4i(act_dim + 1) + pad4i(act_dim);
        h1 = 0; obs = 4 * h1_floats; tobs = obs + 2 * slot_floats; tiles_floats = obs + 3 * slot_floats;
        small = tiles_floats;
        w1 = small + pad4i(small_floats);
        w2 = w1 + 4 * w1_floats;
        cobs = w2 + 4 * w2_floats;
        out = cobs + 2 * obs_dim * E;
        total_floats = out + pad4i(4 * (act_dim + 1) * E);
    }
};

struct Sink128 {       // observation element (k, env): exact float32 copy + fp16 hi / scaled lo operands
    float* plain; __half* hi; __half* lo; int e;
    template <typename R> __device__ __forceinline__ void operator()(int k, R v) const {
        const float x = (float)v;
        plain[k * E + e] = x;
        __half a, b; split_f16(x, a, b);
        const int i = bh_idx(k, e);
        hi[i] = a; lo[i] = b;
    }
};

constexpr int D1H_COL = 0;         // + net * 64 + acc * 32
constexpr int D2H_COL = 128;

struct Ctx128 {
    uint32_t tmem;
    uint32_t w1_s[2][2], w2_s[2][2], h1_s[2][2];    // [net][hi/lo]
    __half* h1_p[2][2];
    const float* small; SmallLayout SL;
    float* s_out;
    uint64_t* bar;
    int obsp, od;
};

// NETS = 2: hidden 128, actor and critic are two M128 tiles.  NETS = 1: hidden 64, actor and critic FUSED into one width-128 net --
// layer 1 [W1_actor; W1_critic] (they share the observation), layer 2 the block-diagonal diag(W2_actor, W2_critic) over the stacked
// hidden vector (same number of MMAs as two M64 tiles, half the MMAs of the TF32 form); rows 0..63 actor units, 64..127 critic units.
template <int A, int NETS>
__device__ __noinline__ void forward_f16(const Ctx128& c /* in shared memory */, uint32_t obs_hi_s, uint32_t obs_lo_s, int act, uint32_t& fwd) {
    constexpr int HID = 128;
    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    const uint32_t parity = fwd & 1u;
    constexpr uint32_t I = idesc_f16(128, E);
    constexpr uint32_t LBO_A = 128 * 16;
    TC_TICK0();
    if (w == 0) {
      if (l == 0) {
        tc_fence_after();
        const int ksteps = c.obsp >> 4;
        const uint32_t tm = c.tmem;
        const uint64_t dB[2] = {make_desc(obs_hi_s, LBO_B, SBO), make_desc(obs_lo_s, LBO_B, SBO)};
#pragma unroll
        for (int net = 0; net < NETS; ++net) {
            const uint64_t dA[2] = {make_desc(c.w1_s[net][0], LBO_A, SBO), make_desc(c.w1_s[net][1], LBO_A, SBO)};
#pragma unroll 1
            for (int ks = 0; ks < ksteps; ++ks) {
                const uint64_t ka = (uint64_t)((ks * 2 * LBO_A) >> 4), kb = (uint64_t)((ks * 2 * LBO_B) >> 4);
                mma_f16(tm + D1H_COL + net * 64, dA[0] + ka, dB[0] + kb, I, ks > 0);             // hi . hi
                mma_f16(tm + D1H_COL + net * 64 + 32, dA[0] + ka, dB[1] + kb, I, ks > 0);        // hi . lo'
                mma_f16(tm + D1H_COL + net * 64 + 32, dA[1] + ka, dB[0] + kb, I, 1u);            // lo' . hi
            }
        }
        mma_commit(c.bar);
      }
      __syncwarp();
    }
    TC_TICK(0);
    mbar_wait_bounded(c.bar, parity);
    tc_fence_after();
    TC_TICK(1);
    // NETS = 2: warp w serves net w >> 2, all 32 envs of its unit.  NETS = 1: both warps of a quadrant serve the fused net, 16 envs each.
    const int net = NETS == 2 ? (w >> 2) : 0, q = w & 3, u = 32 * q + l;          // this thread's unit (a TMEM lane)
    constexpr int NH = NETS == 2 ? 2 : 1;                                           // halves of 16 envs per thread
    const int h0 = NETS == 2 ? 0 : (w >> 2);
    {   // ---- epilogue 1
        const float b = c.small[c.SL.b1 + net * HID + u];
        __half* hi = c.h1_p[net][0]; __half* lo = c.h1_p[net][1];
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
            const int eh = h0 + hh;
            const uint32_t ta = c.tmem + ((uint32_t)(32 * q) << 16) + D1H_COL + net * 64 + 16 * eh;
            float x[16], y[16];
            tmem_ld16(ta, x); tmem_ld16(ta + 32, y);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = fmaf(y[j], LO_UNSCALE, x[j]) + b;
            act16<0>(x, act);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                __half a, bb; split_f16(x[j], a, bb);
                const int i = bh_idx(u, 16 * eh + j);
                hi[i] = a; lo[i] = bb;
            }
        }
    }
    TC_TICK(2);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    TC_TICK(3);
    if (w == 0) {
      if (l == 0) {
        tc_fence_after();
        const uint32_t tm = c.tmem;
        uint64_t dA[2][2], dB[2][2];
#pragma unroll
        for (int n2 = 0; n2 < NETS; ++n2)
#pragma unroll
            for (int p = 0; p < 2; ++p) { dA[n2][p] = make_desc(c.w2_s[n2][p], LBO_A, SBO); dB[n2][p] = make_desc(c.h1_s[n2][p], LBO_B, SBO); }
#pragma unroll
        for (int ks = 0; ks < HID / 16; ++ks) {
            constexpr uint64_t one = 1;
            const uint64_t ka = one * ((ks * 2 * LBO_A) >> 4), kb = one * ((ks * 2 * LBO_B) >> 4);
#pragma unroll
            for (int n2 = 0; n2 < NETS; ++n2) {
                mma_f16(tm + D2H_COL + n2 * 64, dA[n2][0] + ka, dB[n2][0] + kb, I, ks > 0);
                mma_f16(tm + D2H_COL + n2 * 64 + 32, dA[n2][0] + ka, dB[n2][1] + kb, I, ks > 0);
                mma_f16(tm + D2H_COL + n2 * 64 + 32, dA[n2][1] + ka, dB[n2][0] + kb, I, 1u);
            }
        }
        mma_commit(c.bar + 1);
      }
      __syncwarp();
    }
    TC_TICK(4);
    mbar_wait_bounded(c.bar + 1, parity);
    tc_fence_after();
    TC_TICK(5);
    if constexpr (NETS == 2) {   // ---- epilogue 2 + fused output layer: reduce-scatter over the 32 units of the quadrant (all five lane bits); lane l ends with env l
        const float b = c.small[c.SL.b2 + net * HID + u];
        float x[32];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const uint32_t ta = c.tmem + ((uint32_t)(32 * q) << 16) + D2H_COL + net * 64 + 16 * hh;
            float a0[16], a1[16];
            tmem_ld16(ta, a0); tmem_ld16(ta + 32, a1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) a0[j] = fmaf(a1[j], LO_UNSCALE, a0[j]) + b;
            act16<0>(a0, act);
#pragma unroll
            for (int j = 0; j < 16; ++j) x[16 * hh + j] = a0[j];
        }
        const int n_out = net ? 1 : A, row0 = net ? A : 0;
#pragma unroll
        for (int o = 0; o < A; ++o) {
            if (o < n_out) {
                const float w3 = c.small[c.SL.w3 + (row0 + o) * HID + u];
                float p16[16], p8[8], p4[4], p2[2];
                {
                    const bool up = (l & 16) != 0;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float a0 = x[j] * w3, a1 = x[j + 16] * w3;
                        const float r = __shfl_xor_sync(0xffffffffu, up ? a0 : a1, 16);
                        p16[j] = (up ? a1 : a0) + r;
                    }
                }
                {
                    const bool up = (l & 8) != 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const float r = __shfl_xor_sync(0xffffffffu, up ? p16[j] : p16[j + 8], 8); p8[j] = (up ? p16[j + 8] : p16[j]) + r; }
                }
                {
                    const bool up = (l & 4) != 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const float r = __shfl_xor_sync(0xffffffffu, up ? p8[j] : p8[j + 4], 4); p4[j] = (up ? p8[j + 4] : p8[j]) + r; }
                }
                {
                    const bool up = (l & 2) != 0;
#pragma unroll
                    for (int j = 0; j < 2; ++j) { const float r = __shfl_xor_sync(0xffffffffu, up ? p4[j] : p4[j + 2], 2); p2[j] = (up ? p4[j + 2] : p4[j]) + r; }
                }
                const bool up = (l & 1) != 0;
                const float r = __shfl_xor_sync(0xffffffffu, up ? p2[0] : p2[1], 1);
                c.s_out[(q * (A + 1) + row0 + o) * E + l] = (up ? p2[1] : p2[0]) + r;
            }
        }
    } else {   // ---- fused width-128 net (hidden 64): quadrants 0, 1 hold actor units, 2, 3 critic units; 16 envs per thread
        constexpr int H = 64;
        const int eh = w >> 2, cr = q >> 1, unit = u & 63;             // env half, critic?, unit inside its net
        const uint32_t ta = c.tmem + ((uint32_t)(32 * q) << 16) + D2H_COL + 16 * eh;
        float x[16], a1[16];
        tmem_ld16(ta, x); tmem_ld16(ta + 32, a1);
        tmem_ld_wait();
        const float b = c.small[c.SL.b2 + u];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = fmaf(a1[j], LO_UNSCALE, x[j]) + b;
        act16<0>(x, act);
        // the quadrant's nets' output rows; every quadrant also writes zeros for the other net's outputs (mlp_out sums all four)
#pragma unroll
        for (int o = 0; o < A + 1; ++o) {
            const bool mine = (o == A) == (cr != 0);                   // warp-uniform
            float res = 0.f;
            if (mine) {
                const float w3 = c.small[c.SL.w3 + o * H + unit];
                float p8[8], p4[4], p2[2];
                {
                    const bool up = (l & 16) != 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const float r = __shfl_xor_sync(0xffffffffu, (up ? x[j] : x[j + 8]) * w3, 16); p8[j] = (up ? x[j + 8] : x[j]) * w3 + r; }
                }
                {
                    const bool up = (l & 8) != 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const float r = __shfl_xor_sync(0xffffffffu, up ? p8[j] : p8[j + 4], 8); p4[j] = (up ? p8[j + 4] : p8[j]) + r; }
                }
                {
                    const bool up = (l & 4) != 0;
#pragma unroll
                    for (int j = 0; j < 2; ++j) { const float r = __shfl_xor_sync(0xffffffffu, up ? p4[j] : p4[j + 2], 4); p2[j] = (up ? p4[j + 2] : p4[j]) + r; }
                }
                const bool up = (l & 2) != 0;
                const float r = __shfl_xor_sync(0xffffffffu, up ? p2[0] : p2[1], 2);
                res = (up ? p2[1] : p2[0]) + r;
                res += __shfl_xor_sync(0xffffffffu, res, 1);           // lanes 2e, 2e + 1 hold the two halves of env e's sum
            }
            if ((l & 1) == 0) c.s_out[(q * (A + 1) + o) * E + 16 * eh + (l >> 1)] = res;
        }
    }
    TC_TICK(6);
    tc_fence_before();
    __syncthreads();
    TC_TICK(7);
    fwd += 1;
}

// ---- per-width dispatch used by the collector body (forward declaration of the out-of-line normaliser pass below) --------------------------------------------------------------------------
template <int HID> struct Variant;
template <> struct Variant<64> {
    using Plan = Smem<64>; using CtxT = Ctx;
    static __device__ __forceinline__ TcObsSink sink(float* slot, const Plan& S, int, int e) { return TcObsSink{slot, slot + S.obs_floats, e}; }
    static __device__ __forceinline__ float val(const float* slot, const Plan& S, int k, int en) { const int i = b_idx(k, en); return slot[i] + slot[S.obs_floats + i]; }
    template <int A> static __device__ __forceinline__ void fwd(const CtxT& c, float* slot, const Plan& S, int act, uint32_t& f) {
        forward<64, A>(c, smem_u32(slot), (uint32_t)S.obs_floats * 4u, act, f);
    }
};
constexpr int HID64F = 1064;       // variant tag: hidden 64 as ONE fused width-128 net with fp16 operands (forward_f16<A, 1>)
struct Smem64F {
    int small, w1, w2, h1, obs, tobs, cobs, out, xg, total_floats;
    int OBSP, w1_floats, w2_floats, h1_floats, obs_tile_floats, slot_floats, small_floats, tiles_floats;
    __host__ __device__ Smem64F(int obs_dim, int act_dim, int xg_floats) {
        OBSP = pad16(obs_dim);
        w1_floats = 128 * OBSP / 2;                  // one fp16 copy of the fused layer-1 A tile [128 rows][OBSP]
        w2_floats = 128 * 128 / 2;                   // one fp16 copy of diag(W2_actor, W2_critic)
        h1_floats = bh_tile_floats(128);
        obs_tile_floats = bh_tile_floats(OBSP);
        slot_floats = pad4i(obs_dim * E) + 2 * obs_tile_floats;
        small_floats = pad4i(2 * 64) * 2 + pad4i((act_dim + 1) * 64) + pad4i(act_dim + 1) + pad4i(act_dim);
        h1 = 0; obs = 2 * h1_floats; tobs = obs + 2 * slot_floats; tiles_floats = obs + 3 * slot_floats;
        small = tiles_floats;
        w1 = small + pad4i(small_floats);
        w2 = w1 + 2 * w1_floats;
        cobs = w2 + 2 * w2_floats;
        out = cobs + 2 * obs_dim * E;
        xg = out + pad4i(4 * (act_dim + 1) * E);
        total_floats = xg + pad4i(xg_floats);
    }
};
template <> struct Variant<HID64F> {
    using Plan = Smem64F; using CtxT = Ctx128;
    static __device__ __forceinline__ Sink128 sink(float* slot, const Plan& S, int od, int e) {
        __half* hi = reinterpret_cast<__half*>(slot + pad4i(od * E));
        return Sink128{slot, hi, hi + 2 * S.obs_tile_floats, e};
    }
    static __device__ __forceinline__ float val(const float* slot, const Plan&, int k, int en) { return slot[k * E + en]; }
    template <int A> static __device__ __forceinline__ void fwd(const CtxT& c, float* slot, const Plan& S, int act, uint32_t& f) {
        const uint32_t hi = smem_u32(slot) + (uint32_t)pad4i(c.od * E) * 4u;
        forward_f16<A, 1>(c, hi, hi + (uint32_t)S.obs_tile_floats * 4u, act, f);
    }
};
template <> struct Variant<128> {
    using Plan = Smem128; using CtxT = Ctx128;
    static __device__ __forceinline__ Sink128 sink(float* slot, const Plan& S, int od, int e) {
        __half* hi = reinterpret_cast<__half*>(slot + pad4i(od * E));
        return Sink128{slot, hi, hi + 2 * S.obs_tile_floats, e};
    }
    static __device__ __forceinline__ float val(const float* slot, const Plan&, int k, int en) { return slot[k * E + en]; }
    template <int A> static __device__ __forceinline__ void fwd(const CtxT& c, float* slot, const Plan& S, int act, uint32_t& f) {
        const uint32_t hi = smem_u32(slot) + (uint32_t)pad4i(c.od * E) * 4u;
        forward_f16<A, 2>(c, hi, hi + (uint32_t)S.obs_tile_floats * 4u, act, f);
    }
};

// out of line for the same reason as norm_obs_pass<E> of the FFMA2 collector
template <int HID> static __device__ __noinline__ void norm_obs_pass(float* slot, const typename Variant<HID>::Plan& S, int od, int ol, float* raw,
                                                                     const NormState* s, bool have_norm) {
    const bool on = have_norm && s->obs_on;
    const auto sink = Variant<HID>::sink(slot, S, od, ol);
    for (int k = 0; k < od; ++k) {
        const float x = Variant<HID>::val(slot, S, k, ol);
        if (raw) raw[k] = x;
        if (on) sink(k, norm_obs_value(s, k, x));
    }
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// the collector.  Body = rollout_kernel<SYS, R, 64, 32, 256> with the tensor-core forward and B-tile observation tiles.
// ---------------------------------------------------------------------------------------------
template <int SYS, typename R, int HID>
__global__ void __launch_bounds__(tc::NT, 1) rollout_tc_kernel(DevTables T, RolloutParams P) {
    constexpr int E = tc::E, NT = tc::NT;
    constexpr int NX = SysDims<SYS>::NX, A = SysDims<SYS>::NU;
    constexpr int NZ = NX + 2 * A;
    constexpr int NW = 4;
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t s_bar;          // TMA staging of the blob
    __shared__ __align__(8) uint64_t s_mbar[2];      // layer-1 / layer-2 accumulators ready
    __shared__ uint32_t s_tmem;
    using V = tc::Variant<HID>;
    constexpr int HIDN = (HID == tc::HID64F) ? 64 : HID;      // network width (HID is the variant tag)
    using XG = std::conditional_t<HID == 128, double, R>;      // X_GOAL: shared copy in the integrator's precision (hidden 64) / global table (hidden 128)
    __shared__ typename V::CtxT s_ctx;
    __shared__ HotCfg<R> s_hot;
    __shared__ double s_z[2 * NZ * E];
    __shared__ float s_eps[2 * A * E];
    __shared__ uint32_t s_tsteps0[E];
    __shared__ int s_cstep[2 * E];
    __shared__ int s_flag[2 * E];
    constexpr int PS = SysDims<SYS>::PS;
    constexpr int NC = SCG_MAX_INERTIAL + SCG_NINV + PS;
    __shared__ double s_cand[2 * NC * E];
    __shared__ int s_cand_ep[2 * E];
    __shared__ int s_cand_raw[2 * E];          // draws of that episode sit in the slot (first half of the preparation, see rollout_kernel)
    __shared__ int s_epi[E];
    __shared__ float s_rec[2 * (A + 2) * E];
    __shared__ float s_stat[6 * E];
    __shared__ int s_epc[E];
    __shared__ NormState s_norm;

    const int od = T.obs_dim, N = T.N, nc = T.n_con;
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * E;
    const typename V::Plan S(od, A, T.L * NX * (int)(sizeof(R) / 4));
    const BlobLayout L = blob_layout(od, A, HIDN);
    const tc::SmallLayout SL = tc::small_layout(A, HIDN);
    float* s_small = smem + S.small;
    float* s_cobs = smem + S.cobs;
    float* s_out = smem + S.out;
    const XG* s_xg;
    // observation slots: [cur | next] ping-pong + terminal; a slot holds the operand tiles (hi / lo) of one observation of the 32 envs
    const int SLOT = S.slot_floats;
    float* s_obs = smem + S.obs;
    float* s_tobs = smem + S.tobs;

    // ---- stage the packed policy (hidden 64: TMA bulk copy into shared memory), allocate tensor memory, load the hot constants
    if (tid == 0) { mbar_init(&s_bar, 1); mbar_init(&s_mbar[0], 1); mbar_init(&s_mbar[1], 1); fence_mbar_init(); }
    if (tid == 32) hot_load<SYS, R>(T, &s_hot);
    const HotCfg<R>* h = &s_hot;
    __syncthreads();
    if constexpr (HID == 64) {
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)L.total * 4u;
            mbar_expect_tx(&s_bar, bytes);
            tma_bulk_g2s(smem + S.blob, P.policy, bytes, &s_bar);
        }
    }
    const int warp = tid >> 5, lane = tid & 31;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(tc::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    const int role = (NT / 32 - 1) - warp;          // 0 = owner warp, 1..3 = helper warps
    const int ol = tid - (NT - 32);
    const bool owner = ol >= 0 && ol < E;
    const int n = n0 + ol;
    const bool live = owner && n < N;
    EnvRegs<SYS, R> e;
    if (live) env_load<SYS, R>(T, n, e);
    if (owner) { s_tsteps0[ol] = live ? e.total_steps : 0u; s_epi[ol] = live ? e.episode : 0; s_cand_ep[ol] = -1; s_cand_ep[E + ol] = -1; s_cand_raw[ol] = -1; s_cand_raw[E + ol] = -1; }
    const bool use_cand = h->n_off_draws == 0;
    float ep_ret_sum = 0.f, ep_len_sum = 0.f, ep_cv_sum = 0.f; int ep_count = 0;
    if constexpr (HID == 64) {
        R* xg = reinterpret_cast<R*>(smem + S.xg);
        for (int i = tid; i < T.L * NX; i += NT) xg[i] = (R)T.x_goal[i];
        s_xg = xg;
        const float* s_blob = smem + S.blob;
        mbar_wait(&s_bar, 0);
        __syncthreads();
        // ---- convert the blob: biases / output rows / logstd as they are, W1 and W2 as TF32 hi / lo A tiles (K-major canonical layout)
        for (int i = tid; i < 2 * HID; i += NT) { s_small[SL.b1 + i] = s_blob[L.b1 + i]; s_small[SL.b2 + i] = s_blob[L.b2 + i]; }
        for (int i = tid; i < (A + 1) * HID; i += NT) s_small[SL.w3 + i] = s_blob[L.w3 + i];
        if (tid < 4) { s_small[SL.b3 + tid] = tid < A + 1 ? s_blob[L.b3 + tid] : 0.f; s_small[SL.logstd + tid] = tid < A ? s_blob[L.logstd + tid] : 0.f; }
        if (tid == 4 && A + 1 > 4) s_small[SL.b3 + 4] = s_blob[L.b3 + 4];
        float* w1h = smem + S.w1; float* w1l = w1h + S.w1_floats;
        for (int i = tid; i < 2 * HID * S.OBSP; i += NT) {
            const int m = i % (2 * HID), k = i / (2 * HID);
            const float x = k < od ? s_blob[L.w1t + ((k >> 1) * 2 * HID + m) * 2 + (k & 1)] : 0.f;
            const float hv = tc::tf32_hi(x);
            const int j = tc::a_idx(m, k, 2 * HID);
            w1h[j] = hv; w1l[j] = x - hv;
        }
        for (int net = 0; net < 2; ++net) {
            float* w2h = smem + S.w2 + net * 2 * S.w2_floats; float* w2l = w2h + S.w2_floats;
            for (int i = tid; i < HID * HID; i += NT) {
                const int u = i % HID, k = i / HID;
                const float x = s_blob[L.w2t + net * HID * HID + ((k >> 1) * HID + u) * 2 + (k & 1)];
                const float hv = tc::tf32_hi(x);
                const int j = tc::a_idx(u, k, HID);
                w2h[j] = hv; w2l[j] = x - hv;
            }
        }
        __syncthreads();      // the blob is dead from here on: the activation tiles alias it
    } else if constexpr (HID == tc::HID64F) {
        // hidden 64, fp16 operands, actor + critic fused into one width-128 net: [W1_actor; W1_critic] and diag(W2_actor, W2_critic)
        R* xg = reinterpret_cast<R*>(smem + S.xg);
        for (int i = tid; i < T.L * NX; i += NT) xg[i] = (R)T.x_goal[i];
        s_xg = xg;
        const float* __restrict__ g = P.policy;
        for (int i = tid; i < 2 * 64; i += NT) { s_small[SL.b1 + i] = __ldg(g + L.b1 + i); s_small[SL.b2 + i] = __ldg(g + L.b2 + i); }
        for (int i = tid; i < (A + 1) * 64; i += NT) s_small[SL.w3 + i] = __ldg(g + L.w3 + i);
        if (tid < 4) { s_small[SL.b3 + tid] = tid < A + 1 ? __ldg(g + L.b3 + tid) : 0.f; s_small[SL.logstd + tid] = tid < A ? __ldg(g + L.logstd + tid) : 0.f; }
        if (tid == 4 && A + 1 > 4) s_small[SL.b3 + 4] = __ldg(g + L.b3 + 4);
        for (int i = tid; i < 2 * (S.w1_floats + S.w2_floats); i += NT) smem[S.w1 + i] = 0.f;     // padded K columns and the off-diagonal blocks
        __syncthreads();
        __half* w1h = reinterpret_cast<__half*>(smem + S.w1); __half* w1l = w1h + 2 * S.w1_floats;
        for (int i = tid; i < 128 * od; i += NT) {             // W1P [k/2][2H][2]: consecutive threads read consecutive floats
            const int par = i & 1, m = (i >> 1) % 128, k = 2 * ((i >> 1) / 128) + par;
            __half a, b; tc::split_f16(__ldg(g + L.w1t + i), a, b);
            const int j = tc::ah_idx(m, k, 128);
            w1h[j] = a; w1l[j] = b;
        }
        __half* w2h = reinterpret_cast<__half*>(smem + S.w2); __half* w2l = w2h + 2 * S.w2_floats;
        for (int i = tid; i < 2 * 64 * 64; i += NT) {          // W2P [net][k/2][H][2]
            const int net = i / (64 * 64), r = i % (64 * 64), par = r & 1, u = (r >> 1) % 64, k = 2 * ((r >> 1) / 64) + par;
            __half a, b; tc::split_f16(__ldg(g + L.w2t + i), a, b);
            const int j = tc::ah_idx(net * 64 + u, net * 64 + k, 128);
            w2h[j] = a; w2l[j] = b;
        }
    } else {
        // hidden 128: the packed blob (147 KB) does not fit next to its converted copy -- convert straight from global memory (L2)
        s_xg = T.x_goal;
        const float* __restrict__ g = P.policy;
        for (int i = tid; i < 2 * HID; i += NT) { s_small[SL.b1 + i] = __ldg(g + L.b1 + i); s_small[SL.b2 + i] = __ldg(g + L.b2 + i); }
        for (int i = tid; i < (A + 1) * HID; i += NT) s_small[SL.w3 + i] = __ldg(g + L.w3 + i);
        if (tid < 4) { s_small[SL.b3 + tid] = tid < A + 1 ? __ldg(g + L.b3 + tid) : 0.f; s_small[SL.logstd + tid] = tid < A ? __ldg(g + L.logstd + tid) : 0.f; }
        if (tid == 4 && A + 1 > 4) s_small[SL.b3 + 4] = __ldg(g + L.b3 + 4);
        for (int net = 0; net < 2; ++net) {
            __half* w1h = reinterpret_cast<__half*>(smem + S.w1 + net * 2 * S.w1_floats); __half* w1l = w1h + 2 * S.w1_floats;
            for (int i = tid; i < HID * S.OBSP; i += NT) {
                const int m = i % HID, k = i / HID;
                const float x = k < od ? __ldg(g + L.w1t + ((k >> 1) * 2 * HID + net * HID + m) * 2 + (k & 1)) : 0.f;
                __half a, b; tc::split_f16(x, a, b);
                const int j = tc::ah_idx(m, k, HID);
                w1h[j] = a; w1l[j] = b;
            }
            __half* w2h = reinterpret_cast<__half*>(smem + S.w2 + net * 2 * S.w2_floats); __half* w2l = w2h + 2 * S.w2_floats;
            for (int i = tid; i < HID * HID; i += NT) {          // consecutive threads read consecutive floats of W2P [k/2][H][2]
                const int par = i & 1, u = (i >> 1) % HID, kp = (i >> 1) / HID, k = 2 * kp + par;
                const float x = __ldg(g + L.w2t + net * HID * HID + i);
                __half a, b; tc::split_f16(x, a, b);
                const int j = tc::ah_idx(u, k, HID);
                w2h[j] = a; w2l[j] = b;
            }
        }
    }
    for (int i = tid; i < S.tiles_floats; i += NT) smem[S.h1 + i] = 0.f;     // zero: padded K columns and dead env rows stay 0
    __syncthreads();
    const float* obs_src = P.obs0 ? P.obs0 : T.st.cur_obs;
    for (int i = tid; i < E * od; i += NT) {
        const int ee = i / od, k = i % od;
        const float x = (n0 + ee < N) ? obs_src[(size_t)(n0 + ee) * od + k] : 0.f;
        V::sink(s_obs, S, od, ee)(k, x);
    }
    if (P.norm) norm_load(&s_norm, P.norm, od, tid, NT);
    float sd[A], lsd[A], inv2var[A];
#pragma unroll
    for (int k = 0; k < A; ++k) { lsd[k] = s_small[SL.logstd + k]; sd[k] = expf(lsd[k]); inv2var[k] = 1.0f / (2.f * sd[k] * sd[k]); }
    const float log_sqrt_2pi = 0.91893853320467267f;
    const bool philox_eps = (P.eps == nullptr);
    const bool obs_noise = h->has_dist[SCG_MODE_OBSERVATION] != 0;
    const bool fast = use_cand && !obs_noise && !h->done_on_violation && !h->use_penalty;
    const bool relay = fast && (P.n_peers > 0 || P.mc_off != 0);
    const int OBS_TILE = od * E;
    const int ee = lane;
    const bool hl = role >= 1 && role <= 3 && ee < E && n0 + ee < N;
    float h_a = 0.f, h_b = 0.f, h_sa = 0.f, h_sb = 0.f; int h_cnt = 0;
    if (hl && fast) {
        if (role == 2) { h_a = T.st.ep_acc[n0 + ee]; h_b = T.st.ep_acc[N + n0 + ee]; }
        if (role == 1) h_a = T.st.ep_acc[2 * N + n0 + ee];
    }
    fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (tid == 0) {
        typename V::CtxT& cx_tc = s_ctx;
        cx_tc.tmem = s_tmem;
        const uint32_t w1s = smem_u32(smem + S.w1), w2s = smem_u32(smem + S.w2), h1s = smem_u32(smem + S.h1);
        if constexpr (HID == 64) {
            cx_tc.w1_s[0] = w1s; cx_tc.w1_s[1] = w1s + S.w1_floats * 4;
        } else {
            for (int net = 0; net < 2; ++net)
                for (int p = 0; p < 2; ++p) cx_tc.w1_s[net][p] = w1s + (net * 2 + p) * S.w1_floats * 4;
            cx_tc.od = od;
        }
        // (the fused hidden-64 variant only uses net 0 of the [net][hi/lo] tables: its single tiles sit at the same offsets)
        for (int net = 0; net < 2; ++net)
            for (int p = 0; p < 2; ++p) {
                cx_tc.w2_s[net][p] = w2s + (net * 2 + p) * S.w2_floats * 4;
                cx_tc.h1_s[net][p] = h1s + (net * 2 + p) * S.h1_floats * 4;
                cx_tc.h1_p[net][p] = reinterpret_cast<std::remove_reference_t<decltype(cx_tc.h1_p[0][0])>>(smem + S.h1 + (net * 2 + p) * S.h1_floats);
            }
        cx_tc.small = s_small; cx_tc.SL = SL; cx_tc.s_out = s_out; cx_tc.bar = s_mbar; cx_tc.obsp = S.OBSP;
    }
    const typename V::CtxT& cx_tc = s_ctx;     // published by the barrier below (before the first forward)
    uint32_t fwd = 0;
    auto tile_val = [&](const float* slot, int k, int en) { return V::val(slot, S, k, en); };   // the float32 observation, exactly

    auto store_policy_record = [&](size_t tn, const float* act, float val, float logp, int done, int truncated) {
        if (A == 2) pstore(P, reinterpret_cast<float2*>(P.out.d_act + tn * A), make_float2(act[0], act[1 % A]));
        else if (A == 4) pstore(P, reinterpret_cast<float4*>(P.out.d_act + tn * A), make_float4(act[0], act[1 % A], act[2 % A], act[3 % A]));
        else pstore(P, P.out.d_act + tn * A, act[0]);
        pstore(P, P.out.d_v + tn, val);
        pstore(P, P.out.d_logp + tn, logp);
        pstore(P, P.out.d_mask + tn, done ? 0.f : 1.f);
        if (!truncated) pstore(P, P.out.d_terminal_v + tn, 0.f);
    };
    auto process_record = [&](int tp) {
        const int bsel = tp & 1;
        const double* zb = s_z + bsel * NZ * E + ee;
        const size_t tnp = (size_t)tp * N + n0 + ee;
        const int fl = s_flag[bsel * E + ee], dn = fl & 1;
        if (role == 1) {
            if (relay) {
                const float* rb = s_rec + bsel * (A + 2) * E + ee;
                float act[A];
#pragma unroll
                for (int k = 0; k < A; ++k) act[k] = rb[k * E];
                store_policy_record(tnp, act, rb[A * E], rb[(A + 1) * E], dn, fl >> 1);
            }
            int cv = 0;
            if (nc > 0) cv = env_constraints(T, h->rows, nc, h->con_fast, zb, E, nullptr, P.out.d_cvals ? P.out.d_cvals + tnp * nc : nullptr);
            if (P.out.d_cviol) P.out.d_cviol[tnp] = (float)cv;
            h_a += (float)cv;
            if (dn) { h_sa += h_a; h_a = 0.f; }
        } else {
            R st[NX], ny[A], cl[A], rw, ms;
#pragma unroll
            for (int k = 0; k < NX; ++k) st[k] = (R)zb[k * E];
#pragma unroll
            for (int k = 0; k < A; ++k) { ny[k] = (R)zb[(NX + k) * E]; cl[k] = (R)zb[(NX + A + k) * E]; }
            env_reward_mse<SYS, R, XG>(h, s_xg, st, ny, cl, s_cstep[bsel * E + ee], rw, ms);
            if (P.out.d_rew_raw) P.out.d_rew_raw[tnp] = (float)rw;
            pstore(P, P.out.d_rew + tnp, P.norm ? norm_rew_apply(&s_norm, (float)rw) : (float)rw);
            if (P.out.d_mse) P.out.d_mse[tnp] = (float)ms;
            h_a += (float)rw; h_b += 1.f;
            if (dn) { h_cnt += 1; h_sa += h_a; h_sb += h_b; h_a = 0.f; h_b = 0.f; }
        }
    };
    // episode-start preparation in two halves on two warps (rollout_kernel: draw_candidates / finish_candidates)
    auto draw_candidates_coop = [&]() {              // Philox mode, all 32 lanes: lane j makes draw slot j of one env's episode (rollout_kernel)
        const bool valid = ee < E && n0 + ee < N;
        const int want = valid ? *(volatile int*)&s_epi[ee] : 0;
#pragma unroll 1
        for (int d = 0; d < 2; ++d) {
            const int ep = want + d, slot = ep & 1;
            const bool need = valid && *(volatile int*)&s_cand_ep[slot * E + ee] != ep && *(volatile int*)&s_cand_raw[slot * E + ee] != ep;
            unsigned m = __ballot_sync(0xffffffffu, need);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const int ep_s = __shfl_sync(0xffffffffu, ep, src);
                reset_draw_slot<SYS, R>(h, n0 + src, ep_s, lane, s_cand + (ep_s & 1) * NC * E + src, E, SCG_MAX_INERTIAL + SCG_NINV);
                __syncwarp();
                if (lane == src) sts_release(&s_cand_raw[(ep_s & 1) * E + src], ep_s);
            }
        }
    };
    auto draw_candidates = [&]() {
        const int want = *(volatile int*)&s_epi[ee];
#pragma unroll 1
        for (int d = 0; d < 2; ++d) {
            const int ep = want + d, slot = ep & 1;
            if (*(volatile int*)&s_cand_ep[slot * E + ee] != ep && *(volatile int*)&s_cand_raw[slot * E + ee] != ep) {
                reset_draw<SYS, R>(T, h, n0 + ee, ep, s_cand + slot * NC * E + ee, E, SCG_MAX_INERTIAL + SCG_NINV);
                sts_release(&s_cand_raw[slot * E + ee], ep);
            }
        }
    };
    auto finish_candidates = [&]() {
        const int want = *(volatile int*)&s_epi[ee];
#pragma unroll 1
        for (int d = 0; d < 2; ++d) {
            const int ep = want + d, slot = ep & 1;
            if (*(volatile int*)&s_cand_ep[slot * E + ee] != ep && lds_acquire(&s_cand_raw[slot * E + ee]) == ep) {
                double* cd = s_cand + slot * NC * E + ee;
                EnvRegs<SYS, R> tmp; R initr[NX], st0[NX];
#pragma unroll
                for (int j = 0; j < SCG_MAX_INERTIAL; ++j) tmp.inertial[j] = (R)cd[j * E];
#pragma unroll
                for (int j = 0; j < NX; ++j) initr[j] = (R)cd[(SCG_MAX_INERTIAL + SCG_NINV + j) * E];
                env_derived<SYS, R>(tmp);
                Physics<SYS, R>::set_init(tmp, initr);
                Physics<SYS, R>::state(tmp, st0);
#pragma unroll
                for (int j = 0; j < SCG_NINV; ++j) cd[(SCG_MAX_INERTIAL + j) * E] = (double)tmp.inv[j];
#pragma unroll
                for (int j = 0; j < PS; ++j) cd[(SCG_MAX_INERTIAL + SCG_NINV + j) * E] = (double)tmp.ps[j];
                NoiseSrc<R> nn; nn.row = nullptr; nn.philox = false; nn.slot = 0; nn.stream = 0; nn.idx = 0; nn.key.k0 = nn.key.k1 = nn.key.genv = 0;
                make_obs<SYS, R, XG>(T, h, s_xg, st0, n0 + ee, 0, true, nn, PlainObsSink{s_cobs + slot * OBS_TILE + ee, E});
                sts_release(&s_cand_ep[slot * E + ee], ep);
            }
        }
    };
    constexpr int NIDLE = NT - 128;
    const bool gen_eps = !P.deterministic && philox_eps;
    auto make_eps = [&](int tt) {
        if (tid < E && n0 + tid < N) {
            const uint32_t idx = s_tsteps0[tid] + (uint32_t)tt;
            float* dst = s_eps + (tt & 1) * A * E + tid;
#pragma unroll
            for (int k = 0; k < A; k += 2) {
                uint32_t x[4]; philox4x32_10((uint32_t)(k >> 1), idx, h->genv0 + (uint32_t)(n0 + tid), SCG_STREAM_POLICY, h->key0, h->key1, x);
                dst[k * E] = normal_from_words<float>(x[0], x[1]);
                if (k + 1 < A) dst[(k + 1) * E] = normal_from_words<float>(x[2], x[3]);
            }
        }
    };
    auto write_obs_record = [&](int tt, const float* tile, int t0, int nthr) {
        float* dst = P.out.d_obs + ((size_t)tt * N + n0) * od;
        const int lim = (N - n0 < E ? N - n0 : E) * od;
        int k = t0 % od, en = t0 / od;
        const int dk = nthr % od, de = nthr / od;
        for (int i = t0; i < lim; i += nthr) {
            pstore(P, dst + i, tile_val(tile, k, en));
            k += dk; en += de;
            if (k >= od) { k -= od; en += 1; }
        }
    };
    if (gen_eps) make_eps(0);
    fence_proxy_async();      // the first observation tile was written through the generic proxy
    __syncthreads();

    int cur = 0;
    for (int t = 0; t < P.T; ++t) {
        const size_t tn0 = (size_t)t * N + n0;
        float* s_oc = s_obs + cur * SLOT;
        float* s_on = s_obs + (cur ^ 1) * SLOT;
        float z[A];
#pragma unroll
        for (int k = 0; k < A; ++k) z[k] = 0.f;
        if (!P.deterministic && !philox_eps && live) {
#pragma unroll
            for (int k = 0; k < A; ++k) z[k] = __ldg(P.eps + (tn0 + ol) * A + k);
        }
        V::template fwd<A>(cx_tc, s_oc, S, P.activation, fwd);

        int trunc = 0;
#ifdef SCG_TC_PROF
        const long long own0 = clock64();
#endif
        if (role == 0) {
            float act[A], logp = 0.f, val = 0.f;
            if (live) {
#pragma unroll
                for (int k = 0; k < A; ++k) {
                    const float zk = gen_eps ? s_eps[(t & 1) * A * E + k * E + ol] : z[k];
                    const float mu = mlp_out<E, NW, A>(s_out, s_small, BlobLayout{0, 0, 0, 0, 0, SL.b3, 0, 0}, k, ol);
                    act[k] = fmaf(sd[k], zk, mu);
                    const float d = act[k] - mu;
                    logp += -(d * d) * inv2var[k] - lsd[k] - log_sqrt_2pi;
                }
                val = mlp_out<E, NW, A>(s_out, s_small, BlobLayout{0, 0, 0, 0, 0, SL.b3, 0, 0}, A, ol);
                R action[A], state[NX];
#pragma unroll
                for (int k = 0; k < A; ++k) action[k] = (R)act[k];
                const double* nrow = (!h->philox && T.noise_table && T.D > 0)
                                         ? T.noise_table + ((size_t)((P.noise_cursor + t) % T.Tn) * N + n) * T.D : nullptr;
                StepCtx<SYS, R> cx; NoiseSrc<R> ns;
                env_pre<SYS, R>(T, h, n, e, action, nrow, cx, ns);
                env_physics<SYS, R>(h, e, cx, state);
                double* zb = s_z + (fast ? (t & 1) * NZ * E : 0) + ol;
#pragma unroll
                for (int k = 0; k < NX; ++k) zb[k * E] = (double)state[k];
#pragma unroll
                for (int k = 0; k < A; ++k) { zb[(NX + k) * E] = (double)cx.noisy[k]; zb[(NX + A + k) * E] = (double)cx.clipped[k]; }
                const int step_before = e.ctrl_step;
                const size_t tn = tn0 + ol;
                int goal_reached = 0, oob = 0, done = env_done<SYS, R, XG>(h, s_xg, state, goal_reached, oob), truncated = 0;
                const auto sink_next = V::sink(s_on, S, od, ol), sink_term = V::sink(s_tobs, S, od, ol);
                if (fast) {
                    const int fl = (step_before + 1 >= h->ctrl_steps) ? (done ? 1 : 3) : done;
                    s_cstep[(t & 1) * E + ol] = step_before;
                    s_flag[(t & 1) * E + ol] = fl;
                    done = fl & 1; truncated = fl >> 1;
                    e.total_steps += 1;
                    if (fl != 1) { if (truncated) make_obs<SYS, R, XG>(T, h, s_xg, state, n, step_before, false, ns, sink_term);
                                   else make_obs<SYS, R, XG>(T, h, s_xg, state, n, step_before, false, ns, sink_next); }
                    if (done) {
                        const int slot = e.episode & 1;
                        if (lds_acquire(&s_cand_ep[slot * E + ol]) == e.episode) {
                            const double* cd = s_cand + slot * NC * E + ol;
#pragma unroll
                            for (int j = 0; j < SCG_MAX_INERTIAL; ++j) e.inertial[j] = (R)cd[j * E];
#pragma unroll
                            for (int j = 0; j < SCG_NINV; ++j) e.inv[j] = (R)cd[(SCG_MAX_INERTIAL + j) * E];
#pragma unroll
                            for (int j = 0; j < PS; ++j) e.ps[j] = (R)cd[(SCG_MAX_INERTIAL + SCG_NINV + j) * E];
                            const float* co = s_cobs + slot * OBS_TILE + ol;
                            for (int k = 0; k < od; k += 4) {
                                float cv[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) cv[j] = (k + j < od) ? co[(k + j) * E] : 0.f;
#pragma unroll
                                for (int j = 0; j < 4; ++j) if (k + j < od) sink_next(k + j, cv[j]);
                            }
                            e.ctrl_step = 0; e.episode += 1;
                        } else {
                            env_reset<SYS, R>(T, h, n, e, state);
                            make_obs<SYS, R, XG>(T, h, s_xg, state, n, 0, true, ns, sink_next);
                        }
                        *(volatile int*)&s_epi[ol] = e.episode;
                    } else e.ctrl_step += 1;
                } else {
                    R rew, mse; int cviol = 0;
                    env_reward_mse<SYS, R, XG>(h, s_xg, state, cx.noisy, cx.clipped, step_before, rew, mse);
                    if (nc > 0) cviol = env_constraints(T, h->rows, nc, h->con_fast, zb, E, nullptr, P.out.d_cvals ? P.out.d_cvals + tn * nc : nullptr);
                    e.ctrl_step += 1; e.total_steps += 1;
                    env_finalize<R>(h, e.ctrl_step, cviol, rew, done, truncated);
                    e.ep_ret += (float)rew; e.ep_len += 1.f; e.ep_cv += (float)cviol;
                    if (done) {
                        if (truncated) make_obs<SYS, R, XG>(T, h, s_xg, state, n, step_before, false, ns, sink_term);
                        ep_count += 1; ep_ret_sum += e.ep_ret; ep_len_sum += e.ep_len; ep_cv_sum += e.ep_cv;
                        e.ep_ret = 0.f; e.ep_len = 0.f; e.ep_cv = 0.f;
                        env_reset<SYS, R>(T, h, n, e, state);
                        ns = reset_noise<R>(T, h, n, e.episode);
                        make_obs<SYS, R, XG>(T, h, s_xg, state, n, 0, true, ns, sink_next);
                    } else {
                        make_obs<SYS, R, XG>(T, h, s_xg, state, n, step_before, false, ns, sink_next);
                    }
                    if (P.out.d_rew_raw) P.out.d_rew_raw[tn] = (float)rew;
                    pstore(P, P.out.d_rew + tn, P.norm ? norm_rew_apply(&s_norm, (float)rew) : (float)rew);
                    if (P.out.d_cviol) P.out.d_cviol[tn] = (float)cviol;
                    if (P.out.d_mse) P.out.d_mse[tn] = (float)mse;
                }
                if (P.norm || P.out.d_obs_raw)           // the observation env.step returned: raw copy for the statistics update, normalised in place
                    tc::norm_obs_pass<HID>(s_on, S, od, ol, P.out.d_obs_raw ? P.out.d_obs_raw + tn * od : nullptr, &s_norm, P.norm != nullptr);
                trunc = truncated;
                if (relay) {
                    float* rb = s_rec + (t & 1) * (A + 2) * E + ol;
#pragma unroll
                    for (int k = 0; k < A; ++k) rb[k * E] = act[k];
                    rb[A * E] = val; rb[(A + 1) * E] = logp;
                } else {
                    store_policy_record(tn, act, val, logp, done, truncated);
                }
            }
            fence_proxy_async();          // next / terminal observation tiles: generic-proxy writes -> visible to the tensor core
        } else if (role <= 3) {
            if (fast && role == 3 && h->philox) draw_candidates_coop();
            else if (fast && hl) {
                if (role == 3) draw_candidates();
                else if (t > 0) process_record(t - 1);
            }
        } else {
            write_obs_record(t, s_oc, tid, NIDLE);
            if (gen_eps && t + 1 < P.T) make_eps(t + 1);
            if (fast && warp == 1 && ee < E && n0 + ee < N) finish_candidates();
        }
#ifdef SCG_TC_PROF
        const long long own1 = clock64();
        TC_OWN(8, own0, own1);
#endif
        const int any_trunc = __syncthreads_or(trunc);
#ifdef SCG_TC_PROF
        TC_OWN(9, own1, clock64());
        if (threadIdx.x == 0 && blockIdx.x == 0) tc_ticks[10] += clock64() - own0;     // whole env phase seen by warp 0
#endif
        if (any_trunc) {
            V::template fwd<A>(cx_tc, s_tobs, S, P.activation, fwd);
            if (live && trunc) pstore(P, P.out.d_terminal_v + tn0 + ol, mlp_out<E, NW, A>(s_out, s_small, BlobLayout{0, 0, 0, 0, 0, SL.b3, 0, 0}, A, ol));
            __syncthreads();
        }
        cur ^= 1;
    }
    float* s_ol = s_obs + cur * SLOT;
    if (fast && hl && role != 3) {
        process_record(P.T - 1);
        if (role == 2) { s_stat[ee] = h_a; s_stat[E + ee] = h_b; s_stat[3 * E + ee] = h_sa; s_stat[4 * E + ee] = h_sb; s_epc[ee] = h_cnt; }
        else { s_stat[2 * E + ee] = h_a; s_stat[5 * E + ee] = h_sa; }
    }
    V::template fwd<A>(cx_tc, s_ol, S, P.activation, fwd);
    if (live) {
        if (fast) {
            e.ep_ret = s_stat[ol]; e.ep_len = s_stat[E + ol]; e.ep_cv = s_stat[2 * E + ol];
            ep_ret_sum = s_stat[3 * E + ol]; ep_len_sum = s_stat[4 * E + ol]; ep_cv_sum = s_stat[5 * E + ol]; ep_count = s_epc[ol];
        }
        env_store<SYS, R>(T, n, e);
        for (int k = 0; k < od; ++k) {
            const float v = tile_val(s_ol, k, ol);
            T.st.cur_obs[(size_t)n * od + k] = v;
            if (P.out.d_last_obs) P.out.d_last_obs[(size_t)n * od + k] = v;
        }
        const float lval = mlp_out<E, NW, A>(s_out, s_small, BlobLayout{0, 0, 0, 0, 0, SL.b3, 0, 0}, A, ol);
        if (P.out.d_last_val) P.out.d_last_val[n] = lval;
        if (P.out.d_ep_count) { P.out.d_ep_count[n] = ep_count; P.out.d_ep_return_sum[n] = ep_ret_sum; P.out.d_ep_length_sum[n] = ep_len_sum; P.out.d_ep_cviol_sum[n] = ep_cv_sum; }
        if (P.out.d_ret && P.out.d_adv) {
            float ret = lval, vnext = lval; double adv = 0.0;
            constexpr int CH = 8;
            for (int t1 = P.T; t1 > 0; t1 -= CH) {
                const int cnt = t1 < CH ? t1 : CH;
                float rr[CH], tv[CH], mm[CH], vv[CH];
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    if (j < cnt) {
                        const size_t tn = (size_t)(t1 - 1 - j) * N + n;
                        rr[j] = P.out.d_rew[tn]; tv[j] = P.out.d_terminal_v[tn]; mm[j] = P.out.d_mask[tn]; vv[j] = P.out.d_v[tn];
                    }
                }
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    if (j < cnt) {
                        const size_t tn = (size_t)(t1 - 1 - j) * N + n;
                        const float r = rr[j] + P.gamma * tv[j];
                        if (P.mc_off) pstore(P, P.out.d_rew + tn, r); else P.out.d_rew[tn] = r;
                        const float m = mm[j], v = vv[j];
                        ret = r + P.gamma * m * ret;
                        if (!P.use_gae) adv = (double)(ret - v);
                        else { const float td = r + P.gamma * m * vnext - v; adv = adv * P.lam_d * P.gamma_d * (double)m + (double)td; }
                        if (P.mc_off) { pstore(P, P.out.d_ret + tn, ret); pstore(P, P.out.d_adv + tn, (float)adv); }
                        else { P.out.d_ret[tn] = ret; P.out.d_adv[tn] = (float)adv; }
                        vnext = v;
                    }
                }
            }
        }
    }
    if (P.n_peers > 0 && !P.mc_off && P.out.d_ret && P.out.d_adv) {
        __syncthreads();
        const int cols = N - n0 < E ? N - n0 : E;
        for (int i = tid; i < P.T * E; i += NT) {
            const int tt = i / E, c = i % E;
            if (c < cols) {
                const size_t tn = (size_t)tt * N + n0 + c;
                const float r = P.out.d_rew[tn], q = P.out.d_ret[tn], a = P.out.d_adv[tn];
                for (int g = 0; g < P.n_peers; ++g) {
                    *reinterpret_cast<float*>(reinterpret_cast<char*>(P.out.d_rew + tn) + P.peer_off[g]) = r;
                    *reinterpret_cast<float*>(reinterpret_cast<char*>(P.out.d_ret + tn) + P.peer_off[g]) = q;
                    *reinterpret_cast<float*>(reinterpret_cast<char*>(P.out.d_adv + tn) + P.peer_off[g]) = a;
                }
            }
        }
    }
    // ---- release the tensor memory
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(tc::TMEM_COLS) : "memory");
}
