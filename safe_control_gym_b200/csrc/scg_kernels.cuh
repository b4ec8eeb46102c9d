// scg_kernels.cuh -- sm_100a kernels of the rollout engine.
//
//   reset_kernel   : lane per env, (re-)initialises masked envs            (Quadrotor.reset / CartPole.reset)
//   step_kernel    : lane per env, one control step with caller actions    (env.step / DummyVecEnv.step_wait)
//   rollout_kernel : one CTA owns E envs for the whole horizon.  Per step: fused actor+critic MLP as
//                    register-tiled FP32 GEMMs out of shared memory, Gaussian sampling, env step
//                    (physics sub-steps, reward, done, constraints, disturbances), auto-reset, one coalesced
//                    write of the PPOBuffer record, terminal-value pass on truncation, and after the horizon
//                    the critic bootstrap + reverse GAE scan.  No host round trip inside the horizon.
//                    (PPO.train_step collector, safe_control_gym/controllers/ppo/ppo.py:259-297)
#pragma once
#include "scg_device.cuh"

// ---------------------------------------------------------------------------------------------
// policy blob layout (floats), produced by scg_policy_pack() / the python packer:
//   W1P [OBS/2][2H][2] layer-1 weights, input-PAIR major: (w[k][u], w[k+1][u]) adjacent; actor units 0..H-1, critic H..2H-1
//   b1  [2H]
//   W2P [2][H/2][H][2] layer-2 weights per net, input-pair major
//   b2  [2H]
//   W3  [A+1][H]    A actor output rows then the critic row
//   b3  [pad4(A+1)]
//   logstd [pad4(A)]
// every section is a multiple of 4 floats so that the whole blob can be staged with one TMA bulk copy.
// The pair-major order feeds Blackwell's packed FP32 pipe: one FFMA2 multiplies an (input k, k+1) activation pair
// with the matching weight pair and accumulates the even / odd partial sums of a unit.
// ---------------------------------------------------------------------------------------------
struct BlobLayout {
    int w1t, b1, w2t, b2, w3, b3, logstd, total;
};
__host__ __device__ inline int pad4i(int x) { return (x + 3) & ~3; }
__host__ __device__ inline BlobLayout blob_layout(int obs, int act, int hid) {
    BlobLayout L;
    L.w1t = 0;
    L.b1 = L.w1t + obs * 2 * hid;
    L.w2t = L.b1 + 2 * hid;
    L.b2 = L.w2t + 2 * hid * hid;
    L.w3 = L.b2 + 2 * hid;
    L.b3 = L.w3 + (act + 1) * hid;
    L.logstd = L.b3 + pad4i(act + 1);
    L.total = L.logstd + pad4i(act);
    return L;
}

// ---------------------------------------------------------------------------------------------
// small PTX helpers: mbarrier + TMA bulk copy (global -> shared), sm_90+/sm_100a
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// activations.  tanh: branch-free, |rel err| < 5e-7 (odd polynomial below 0.3, 1 - 2/(e^{2|x|}+1) above)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float tanh_fast(float x) {
    const float ax = fabsf(x);
    const float x2 = x * x;
    float p = fmaf(x2, 0.021869488536155203f, -0.053968253968253971f);   // 62/2835, -17/315
    p = fmaf(x2, p, 0.13333333333333333f);                                  // 2/15
    p = fmaf(x2, p, -0.33333333333333333f);
    p = fmaf(x2 * x, p, x);
    // e^{2|x|} = 2^{2|x| log2 e}
    float e = exp2f(ax * 2.8853900817779268f);
    float t = 1.0f - __fdividef(2.0f, e + 1.0f);
    t = copysignf(t, x);
    return ax < 0.3f ? p : t;
}
// four tanh at once: 1/(e_i + 1) for four values from ONE reciprocal of their product (MUFU.RCP is the scarce pipe here:
// 2 MUFU per tanh -> 1.25).  |x| is clamped to 10 so that the product of four (e^{2|x|} + 1) stays below 2^127.
__device__ __forceinline__ void tanh_fast4(float& x0, float& x1, float& x2, float& x3) {
    float x[4] = {x0, x1, x2, x3}, q[4], p[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float ax = fminf(fabsf(x[i]), 10.0f);
        const float x2_ = x[i] * x[i];
        float pp = fmaf(x2_, 0.021869488536155203f, -0.053968253968253971f);
        pp = fmaf(x2_, pp, 0.13333333333333333f);
        pp = fmaf(x2_, pp, -0.33333333333333333f);
        p[i] = fmaf(x2_ * x[i], pp, x[i]);
        q[i] = exp2f(ax * 2.8853900817779268f) + 1.0f;
    }
    const float q01 = q[0] * q[1], q23 = q[2] * q[3];
    const float r = __fdividef(1.0f, q01 * q23);           // 1 / (q0 q1 q2 q3)
    const float r01 = r * q23, r23 = r * q01;              // 1/(q0 q1), 1/(q2 q3)
    const float inv[4] = {r01 * q[1], r01 * q[0], r23 * q[3], r23 * q[2]};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float t = copysignf(fmaf(-2.0f, inv[i], 1.0f), x[i]);
        x[i] = fabsf(x[i]) < 0.3f ? p[i] : t;
    }
    x0 = x[0]; x1 = x[1]; x2 = x[2]; x3 = x[3];
}
template <int ACT> __device__ __forceinline__ float hidden_act(float x) {
    if (ACT == SCG_ACT_TANH) return tanh_fast(x);
    if (ACT == SCG_ACT_RELU) return fmaxf(x, 0.0f);
    return x > 0.0f ? x : 0.01f * x;
}
// hidden activation over a register tile; the (CTA-uniform) kind is switched once per tile, not per element
template <int TU> __device__ __forceinline__ void apply_hidden_act(float (&acc)[4][TU], int act) {
    if (act == SCG_ACT_TANH) {
#pragma unroll
        for (int j = 0; j < TU; ++j) tanh_fast4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
    } else if (act == SCG_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < TU; ++j) acc[i][j] = fmaxf(acc[i][j], 0.0f);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < TU; ++j) acc[i][j] = acc[i][j] > 0.0f ? acc[i][j] : 0.01f * acc[i][j];
    }
}

// Activation tiles (observations and hidden features) are stored "pair-swizzled": feature pair kp = k/2, env group eg = e/4,
// half = (e/2)&1; the 16-byte chunk ((kp*2+half)*EG+eg) holds (x[e][k], x[e][k+1], x[e+1][k], x[e+1][k+1]) for the even env e
// of the half.  A thread's 4 envs x 2 features are two conflict-free LDS.128 and arrive as ready-made float2 operands.
template <int E> __device__ __forceinline__ int tile_idx(int k, int e) {
    return ((((k >> 1) * 2 + ((e >> 1) & 1)) * (E / 4) + (e >> 2)) << 2) + ((e & 1) << 1) + (k & 1);
}

// ---------------------------------------------------------------------------------------------
// observation sinks
// ---------------------------------------------------------------------------------------------
struct GlobalObsSink {      // row `n` of up to three [N][obs_dim] global buffers
    float* a; float* b; double* c; size_t base;
    template <typename R> __device__ __forceinline__ void operator()(int k, R v) const {
        if (a) a[base + k] = (float)v;
        if (b) b[base + k] = (float)v;
        if (c) c[base + k] = (double)v;
    }
};
template <int E> struct TileObsSink {   // env `e` of the CTA's pair-swizzled observation tile (see tile_idx)
    float* tile; int e;
    template <typename R> __device__ __forceinline__ void operator()(int k, R v) const { tile[tile_idx<E>(k, e)] = (float)v; }
};

// ---------------------------------------------------------------------------------------------
// lane-per-env kernels (gym-style API)
// ---------------------------------------------------------------------------------------------
#define SCG_LANE_BLOCK 128

struct ResetArgs {
    const uint8_t* mask; float* obs; double* obs_f64; double* state_f64;
};

template <int SYS, typename R>
__global__ void __launch_bounds__(SCG_LANE_BLOCK) reset_kernel(DevTables T, ResetArgs A) {
    constexpr int NX = SysDims<SYS>::NX;
    __shared__ HotCfg<R> s_hot;
    if (threadIdx.x == 0) hot_load<SYS, R>(T, &s_hot);
    __syncthreads();
    const HotCfg<R>* h = &s_hot;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= T.N) return;
    const int od = T.obs_dim;
    if (A.mask && !A.mask[n]) {
        // untouched env: still report its current observation
        if (A.obs) for (int k = 0; k < od; ++k) A.obs[(size_t)n * od + k] = T.st.cur_obs[(size_t)n * od + k];
        return;
    }
    EnvRegs<SYS, R> e;
    env_load<SYS, R>(T, n, e);
    R state[NX];
    env_reset<SYS, R>(T, h, n, e, state);
    e.ep_ret = 0.f; e.ep_len = 0.f; e.ep_cv = 0.f;
    T.st.initial_reset[n] = 1;
    env_store<SYS, R>(T, n, e);
    NoiseSrc<R> ns = reset_noise<R>(T, h, n, e.episode);
    GlobalObsSink sink{T.st.cur_obs, A.obs, A.obs_f64, (size_t)n * od};
    make_obs<SYS, R, double>(T, h, T.x_goal, state, n, 0, true, ns, sink);
    if (A.state_f64) for (int k = 0; k < NX; ++k) A.state_f64[(size_t)n * NX + k] = (double)state[k];
}

struct StepArgs {
    const float* act; const double* act_f64; scg_step_out out; long long noise_cursor;
};

template <int SYS, typename R>
__global__ void __launch_bounds__(SCG_LANE_BLOCK) step_kernel(DevTables T, StepArgs A) {
    constexpr int NX = SysDims<SYS>::NX, NU = SysDims<SYS>::NU;
    __shared__ HotCfg<R> s_hot;
    __shared__ double s_z[(NX + NU) * SCG_LANE_BLOCK];
    if (threadIdx.x == 0) hot_load<SYS, R>(T, &s_hot);
    __syncthreads();
    const HotCfg<R>* h = &s_hot;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= T.N) return;
    const scg_config* c = T.cfg;
    const int od = T.obs_dim, nc = T.n_con;
    EnvRegs<SYS, R> e;
    env_load<SYS, R>(T, n, e);
    R action[NU];
#pragma unroll
    for (int k = 0; k < NU; ++k) action[k] = A.act_f64 ? (R)A.act_f64[(size_t)n * NU + k] : (R)A.act[(size_t)n * NU + k];
    const double* nrow = (!h->philox && T.noise_table && T.D > 0) ? T.noise_table + ((size_t)(A.noise_cursor % T.Tn) * T.N + n) * T.D : nullptr;
    R state[NX], rew, mse; StepInfo info; NoiseSrc<R> ns;
    const scg_step_out& o = A.out;
    double* cv = o.d_cvals ? o.d_cvals + (size_t)n * nc : nullptr;
    env_step<SYS, R, double>(T, h, T.x_goal, n, e, action, nrow, s_z + threadIdx.x, SCG_LANE_BLOCK, state, rew, mse, info, cv, nullptr, ns);
    e.ep_ret += (float)rew; e.ep_len += 1.f; e.ep_cv += (float)info.cviol;
    const bool reset_now = info.done && c->auto_reset;
    {   // observation of the transition: the returned obs unless the env auto-resets, always the terminal obs
        GlobalObsSink term{o.d_terminal_obs, nullptr, o.d_terminal_obs_f64, (size_t)n * od};
        GlobalObsSink cur{reset_now ? nullptr : T.st.cur_obs, reset_now ? nullptr : o.d_obs, reset_now ? nullptr : o.d_obs_f64, (size_t)n * od};
        auto both = [&](int k, R v) { term(k, v); cur(k, v); };
        make_obs<SYS, R, double>(T, h, T.x_goal, state, n, e.ctrl_step - 1, false, ns, both);
    }
    if (o.d_terminal_state_f64) for (int k = 0; k < NX; ++k) o.d_terminal_state_f64[(size_t)n * NX + k] = (double)state[k];
    if (reset_now) {   // dummy_vec_env.py:33-38
        env_reset<SYS, R>(T, h, n, e, state);
        e.ep_ret = 0.f; e.ep_len = 0.f; e.ep_cv = 0.f;
        NoiseSrc<R> rs = reset_noise<R>(T, h, n, e.episode);
        GlobalObsSink sink{T.st.cur_obs, o.d_obs, o.d_obs_f64, (size_t)n * od};
        make_obs<SYS, R, double>(T, h, T.x_goal, state, n, 0, true, rs, sink);
    }
    env_store<SYS, R>(T, n, e);
    if (o.d_state_f64) for (int k = 0; k < NX; ++k) o.d_state_f64[(size_t)n * NX + k] = (double)state[k];
    if (o.d_rew) o.d_rew[n] = (float)rew;
    if (o.d_rew_f64) o.d_rew_f64[n] = (double)rew;
    if (o.d_done) o.d_done[n] = (uint8_t)info.done;
    if (o.d_truncated) o.d_truncated[n] = (uint8_t)info.truncated;
    if (o.d_cviol) o.d_cviol[n] = (float)info.cviol;
    if (o.d_mse) o.d_mse[n] = (double)mse;
    if (o.d_out_of_bounds) o.d_out_of_bounds[n] = (uint8_t)info.out_of_bounds;
    if (o.d_goal_reached) o.d_goal_reached[n] = (uint8_t)info.goal_reached;
    if (o.d_current_step) o.d_current_step[n] = info.current_step;
}

// ---------------------------------------------------------------------------------------------
// fused rollout
// ---------------------------------------------------------------------------------------------
struct RolloutParams {
    int T;               // horizon
    int activation;
    int use_gae, deterministic;
    float gamma, lam;
    double gamma_d, lam_d;
    const float* policy; // packed blob (device)
    const float* eps;    // [T][N][A] or null
    long long noise_cursor;
    scg_rollout_out out;
    long long* prof;     // SCG_PROF builds: per-phase cycle counters of CTA 0 (nullable)
};

// shared memory carve-up (floats), all offsets multiples of 4
template <int HID, int E> struct RolloutSmem {
    int blob, obs, tobs, h1, h2, out, xg, total_floats;
    // xg_floats: size of the shared X_GOAL copy in float units (L * nx * sizeof(R) / 4)
    __host__ __device__ RolloutSmem(int obs_dim, int act_dim, int xg_floats) {
        BlobLayout L = blob_layout(obs_dim, act_dim, HID);
        blob = 0;
        obs = blob + L.total;
        h1 = obs + obs_dim * E;
        h2 = h1 + 2 * HID * E;
        tobs = h2;                       // aliases h2: see rollout_kernel (c)
        out = h2 + (2 * HID * E > obs_dim * E ? 2 * HID * E : obs_dim * E);
        xg = out + pad4i((act_dim + 1) * E);
        total_floats = xg + pad4i(xg_floats);
    }
};

// One fused actor+critic forward for the CTA's E envs.
//   in  : s_in  pair-swizzled observation tile (tile_idx)
//   out : s_out [A+1][E]       (A action means, then the value)
// Layers 1 and 2 are register-tiled FP32 GEMMs: each thread owns a (4 envs x TU units) tile, weights and
// activations stream from shared memory as 128-bit loads (weights broadcast across the env groups of a warp,
// activations broadcast across its unit groups).
#ifdef SCG_PROF
__device__ long long g_mlp_prof[8];
#define MLP_TICK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_mlp_prof[i] += clock64() - mt; mt = clock64(); } while (0)
#else
#define MLP_TICK(i) do { } while (0)
#endif

template <int HID, int E, int NT, int A>
__device__ __noinline__ void mlp_forward(const float* __restrict__ s_blob, const BlobLayout L, const float* __restrict__ s_in, float* __restrict__ s_h1,
                                         float* __restrict__ s_h2, float* __restrict__ s_out, int obs_dim, int act) {
    constexpr int EG = E / 4;              // env groups
    constexpr int UG = NT / EG;            // unit groups
    constexpr int TU = (2 * HID) / UG;     // units per thread
    static_assert(TU == 2 || TU == 4, "unsupported tile");
    static_assert(HID % TU == 0 && (2 * HID) % UG == 0, "tile must not straddle nets");
    const int tid = threadIdx.x;
    const int eg = tid % EG, ug = tid / EG;
    const int u0 = ug * TU;
    const int net = u0 / HID, ul0 = u0 % HID;

#ifdef SCG_PROF
    long long mt = clock64();
#endif
    float2 acc[4][TU];   // (even-k, odd-k) partial sums of env i x unit j
    // operands of one input pair: 4 envs x (k, k+1) activations, TU units x (k, k+1) weights
    struct Ops { float4 a01, a23; float4 w[TU / 2]; };
    auto load_ops = [&](Ops& o, const float* __restrict__ a_chunk, const float* __restrict__ w_chunk) {
        o.a01 = *reinterpret_cast<const float4*>(a_chunk);                 // envs 0,1 of the group
        o.a23 = *reinterpret_cast<const float4*>(a_chunk + EG * 4);        // envs 2,3 (the other half)
#pragma unroll
        for (int j = 0; j < TU / 2; ++j) o.w[j] = *reinterpret_cast<const float4*>(w_chunk + 4 * j);
    };
    auto fma_ops = [&](const Ops& o) {                                     // 4*TU packed FMAs (FFMA2)
        const float2 av[4] = {make_float2(o.a01.x, o.a01.y), make_float2(o.a01.z, o.a01.w), make_float2(o.a23.x, o.a23.y), make_float2(o.a23.z, o.a23.w)};
#pragma unroll
        for (int j = 0; j < TU / 2; ++j) {
            const float2 w0 = make_float2(o.w[j].x, o.w[j].y), w1 = make_float2(o.w[j].z, o.w[j].w);
#pragma unroll
            for (int i = 0; i < 4; ++i) { acc[i][2 * j] = __ffma2_rn(av[i], w0, acc[i][2 * j]); acc[i][2 * j + 1] = __ffma2_rn(av[i], w1, acc[i][2 * j + 1]); }
        }
    };
    // two operand sets ping-pong so that the shared-memory latency of pair kp+1 hides behind the FMAs of pair kp
    auto gemm = [&](const float* __restrict__ a, int a_stride, const float* __restrict__ w, int w_stride, int npair) {
        Ops o0, o1;
        load_ops(o0, a, w);
#pragma unroll 2
        for (int kp = 0; kp < npair; kp += 2) {
            const bool has1 = kp + 1 < npair;
            if (has1) load_ops(o1, a + (kp + 1) * a_stride, w + (kp + 1) * w_stride);
            fma_ops(o0);
            if (kp + 2 < npair) load_ops(o0, a + (kp + 2) * a_stride, w + (kp + 2) * w_stride);
            if (has1) fma_ops(o1);
        }
    };
    auto finish = [&](float* __restrict__ dst, int dst_unit0) {   // reduce pairs, activate, store pair-swizzled as the next layer's input
        float h[4][TU];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < TU; ++j) h[i][j] = acc[i][j].x + acc[i][j].y;
        apply_hidden_act<TU>(h, act);
#pragma unroll
        for (int j = 0; j < TU; j += 2) {
            float* d = dst + ((((dst_unit0 + j) >> 1) * 2) * EG + eg) * 4;
            *reinterpret_cast<float4*>(d) = make_float4(h[0][j], h[0][j + 1], h[1][j], h[1][j + 1]);
            *reinterpret_cast<float4*>(d + EG * 4) = make_float4(h[2][j], h[2][j + 1], h[3][j], h[3][j + 1]);
        }
    };
    // ---- layer 1: [E x obs] . [obs x 2H]
    {
        const float* b1 = s_blob + L.b1 + u0;
#pragma unroll
        for (int j = 0; j < TU; ++j) { const float b = b1[j]; acc[0][j] = make_float2(b, 0.f); acc[1][j] = acc[0][j]; acc[2][j] = acc[0][j]; acc[3][j] = acc[0][j]; }
        const float* w = s_blob + L.w1t + u0 * 2;
        const float* a = s_in + eg * 4;
        gemm(a, 2 * EG * 4, w, 2 * HID * 2, obs_dim >> 1);
        MLP_TICK(0);
        finish(s_h1 + net * (HID / 2) * (2 * EG * 4), ul0);
        MLP_TICK(1);
    }
    __syncthreads();
    MLP_TICK(2);
    // ---- layer 2: per net [E x H] . [H x H]
    {
        const float* b2 = s_blob + L.b2 + u0;
#pragma unroll
        for (int j = 0; j < TU; ++j) { const float b = b2[j]; acc[0][j] = make_float2(b, 0.f); acc[1][j] = acc[0][j]; acc[2][j] = acc[0][j]; acc[3][j] = acc[0][j]; }
        const float* w = s_blob + L.w2t + net * HID * HID + ul0 * 2;
        const float* a = s_h1 + net * (HID / 2) * (2 * EG * 4) + eg * 4;
        gemm(a, 2 * EG * 4, w, HID * 2, HID / 2);
        MLP_TICK(3);
        finish(s_h2 + net * (HID / 2) * (2 * EG * 4), ul0);
        MLP_TICK(4);
    }
    __syncthreads();
    MLP_TICK(5);
    // ---- layer 3: A action means from the actor features, 1 value from the critic features
    {
        const int g = tid / E, e = tid % E;
        if (g < A + 1) {
            const float* w = s_blob + L.w3 + g * HID;
            const float* hh = s_h2 + (g < A ? 0 : (HID / 2) * (2 * EG * 4)) + ((((e >> 1) & 1) * EG + (e >> 2)) << 2) + ((e & 1) << 1);
            float2 s0 = make_float2(s_blob[L.b3 + g], 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll 4
            for (int kp = 0; kp < HID / 2; kp += 2) {
                const float4 wv = *reinterpret_cast<const float4*>(w + 2 * kp);
                const float2 h0 = *reinterpret_cast<const float2*>(hh + kp * (2 * EG * 4));
                const float2 h1 = *reinterpret_cast<const float2*>(hh + (kp + 1) * (2 * EG * 4));
                s0 = __ffma2_rn(h0, make_float2(wv.x, wv.y), s0);
                s1 = __ffma2_rn(h1, make_float2(wv.z, wv.w), s1);
            }
            s_out[g * E + e] = (s0.x + s0.y) + (s1.x + s1.y);
        }
    }
    MLP_TICK(6);
    __syncthreads();
    MLP_TICK(7);
}

// named barriers (ids 1, 2) between the owner warp and the helper warps of the env-step section
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

#ifdef SCG_PROF
#define PROF_T(v) long long v = clock64()
#else
#define PROF_T(v) do { } while (0)
#endif

template <int SYS, typename R, int HID, int E, int NT>
__global__ void __launch_bounds__(NT) rollout_kernel(DevTables T, RolloutParams P) {
    constexpr int NX = SysDims<SYS>::NX, A = SysDims<SYS>::NU;
    constexpr int NZ = NX + 2 * A;             // z scratch rows: state, noisy action, clipped action
    // With >= 4 warps the env-step section is split: warp 0 = owners (sampling, action map, physics, done, finalize, reset),
    // warp 1 = constraints, warp 2 = reward + mse, warp 3 = next observation.  Smaller CTAs keep everything on the owner lanes.
    constexpr bool SPLIT = (NT >= 128) && (E <= 32);
    static_assert(NT / E >= A + 1, "layer 3 needs A+1 lane groups");
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ HotCfg<R> s_hot;
    __shared__ double s_z[NZ * E];
    __shared__ float s_eps[A * E];
    __shared__ uint32_t s_tsteps0[E];
    __shared__ int s_cstep[E];                 // ctrl_step BEFORE the increment of the step being processed
    __shared__ int s_cviol[E];
    __shared__ R s_rew[E], s_mse[E];

    const int od = T.obs_dim, N = T.N, nc = T.n_con;
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * E;
    const RolloutSmem<HID, E> S(od, A, T.L * NX * (int)(sizeof(R) / 4));
    const BlobLayout L = blob_layout(od, A, HID);
    float* s_blob = smem + S.blob;
    float* s_obs = smem + S.obs;
    float* s_tobs = smem + S.tobs;
    float* s_h1 = smem + S.h1;
    float* s_h2 = smem + S.h2;
    float* s_out = smem + S.out;
    R* s_xg = reinterpret_cast<R*>(smem + S.xg);      // X_GOAL [L][nx] in the integrator's precision

    // ---- stage the packed policy into shared memory with one TMA bulk copy
    if (tid == 0) { mbar_init(&s_bar, 1); fence_mbar_init(); }
    if (tid == 32) hot_load<SYS, R>(T, &s_hot);
    const HotCfg<R>* h = &s_hot;
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)L.total * 4u;
        mbar_expect_tx(&s_bar, bytes);
        tma_bulk_g2s(s_blob, P.policy, bytes, &s_bar);
    }
    // ---- meanwhile: owner lanes pull their env into registers, everyone fills the observation tile
    const bool owner = tid < E;
    const int n = n0 + tid;
    const bool live = owner && n < N;
    const int warp = tid >> 5, lane = tid & 31;
    EnvRegs<SYS, R> e;
    if (live) env_load<SYS, R>(T, n, e);
    if (owner) s_tsteps0[tid] = live ? e.total_steps : 0u;
    for (int i = tid; i < T.L * NX; i += NT) s_xg[i] = (R)T.x_goal[i];
    for (int i = tid; i < E * od; i += NT) {
        const int ee = i / od, k = i % od;
        s_obs[tile_idx<E>(k, ee)] = (n0 + ee < N) ? T.st.cur_obs[(size_t)(n0 + ee) * od + k] : 0.f;
    }
    float ep_ret_sum = 0.f, ep_len_sum = 0.f, ep_cv_sum = 0.f; int ep_count = 0;
    mbar_wait(&s_bar, 0);
    __syncthreads();
    float sd[A], lsd[A], inv2var[A];
#pragma unroll
    for (int k = 0; k < A; ++k) { lsd[k] = s_blob[L.logstd + k]; sd[k] = expf(lsd[k]); inv2var[k] = 1.0f / (2.f * sd[k] * sd[k]); }
    const float log_sqrt_2pi = 0.91893853320467267f;
    const bool philox_eps = (P.eps == nullptr);
    const bool obs_noise = h->has_dist[SCG_MODE_OBSERVATION] != 0;
    // obs write-out: element i = tid + j*NT of the [E][od] record maps to (env i / od, dim i % od); advance incrementally
    const int ow_k = tid % od, ow_e = tid / od, ow_dk = NT % od, ow_de = NT / od;
#ifdef SCG_PROF
    long long pacc[8]; for (int i = 0; i < 8; ++i) pacc[i] = 0;
#define PROF_ON (P.prof != nullptr && tid == 0)
#endif

    for (int t = 0; t < P.T; ++t) {
        const size_t tn0 = (size_t)t * N + n0;
        PROF_T(k0);
        // standard-normal draws of this step: table mode -> owner lanes issue the loads now (latency hidden by the MLP);
        // Philox mode -> the LAST lane group (idle during layer 3) generates them into shared memory, off the owners' critical path
        float z[A];
#pragma unroll
        for (int k = 0; k < A; ++k) z[k] = 0.f;
        if (!P.deterministic) {
            if (!philox_eps) {
                if (live) {
#pragma unroll
                    for (int k = 0; k < A; ++k) z[k] = __ldg(P.eps + (tn0 + tid) * A + k);
                }
            } else if (tid >= NT - E) {
                const int ee = tid - (NT - E);
                if (n0 + ee < N) {
                    const uint32_t idx = s_tsteps0[ee] + (uint32_t)t;     // == that env's total_steps
#pragma unroll
                    for (int k = 0; k < A; k += 2) {                      // one Philox block serves two action dims
                        uint32_t x[4]; philox4x32_10((uint32_t)(k >> 1), idx, h->genv0 + (uint32_t)(n0 + ee), SCG_STREAM_POLICY, h->key0, h->key1, x);
                        s_eps[k * E + ee] = normal_from_words<float>(x[0], x[1]);
                        if (k + 1 < A) s_eps[(k + 1) * E + ee] = normal_from_words<float>(x[2], x[3]);
                    }
                }
            }
        }
        // (a) PPOBuffer.obs[t] <- observation tile (coalesced), then the policy forward
        {
            float* dst = P.out.d_obs + tn0 * od;
            const int lim = (N - n0 < E ? N - n0 : E) * od;
            int k = ow_k, ee = ow_e;
            for (int i = tid; i < lim; i += NT) {
                dst[i] = s_obs[tile_idx<E>(k, ee)];
                k += ow_dk; ee += ow_de;
                if (k >= od) { k -= od; ee += 1; }
            }
        }
        PROF_T(k1);
        mlp_forward<HID, E, NT, A>(s_blob, L, s_obs, s_h1, s_h2, s_out, od, P.activation);
        PROF_T(k2);

        // (b) env step.  Owner lanes: sample, action map, physics; then reward / constraints / next observation run on helper
        // warps (SPLIT) while the owners evaluate termination; owners finalize, emit the record and auto-reset.
        int trunc = 0;
        if (warp == 0) {
            float act[A], logp = 0.f, val = 0.f;
            R state[NX], rew = (R)0, mse = (R)0;
            StepCtx<SYS, R> cx; NoiseSrc<R> ns; StepInfo info;
            int done = 0, goal_reached = 0, oob = 0, cviol = 0, truncated = 0;
            if (live) {
#pragma unroll
                for (int k = 0; k < A; ++k) {
                    const float zk = (philox_eps && !P.deterministic) ? s_eps[k * E + tid] : z[k];
                    const float mu = s_out[k * E + tid];
                    act[k] = fmaf(sd[k], zk, mu);                         // Normal.sample (ppo_utils.py:228)
                    const float d = act[k] - mu;
                    logp += -(d * d) * inv2var[k] - lsd[k] - log_sqrt_2pi;   // distributions.py:12-21
                }
                val = s_out[A * E + tid];
                R action[A];
#pragma unroll
                for (int k = 0; k < A; ++k) action[k] = (R)act[k];
                const double* nrow = (!h->philox && T.noise_table && T.D > 0)
                                         ? T.noise_table + ((size_t)((P.noise_cursor + t) % T.Tn) * N + n) * T.D : nullptr;
                env_pre<SYS, R>(T, h, n, e, action, nrow, cx, ns);
                env_physics<SYS, R>(h, e, cx, state);
                // publish z = [state; noisy action; clipped action] and the step counter for the helpers / the constraint rows
#pragma unroll
                for (int k = 0; k < NX; ++k) s_z[k * E + tid] = (double)state[k];
#pragma unroll
                for (int k = 0; k < A; ++k) { s_z[(NX + k) * E + tid] = (double)cx.noisy[k]; s_z[(NX + A + k) * E + tid] = (double)cx.clipped[k]; }
                s_cstep[tid] = e.ctrl_step;
            }
            PROF_T(k3);
            if (SPLIT) {
                named_bar_arrive(1, 128);                                  // helpers may start
                if (live) done = env_done<SYS, R, R>(h, s_xg, state, goal_reached, oob);
                named_bar_sync(2, 128);                                    // helpers are done
                if (live) { rew = s_rew[tid]; mse = s_mse[tid]; cviol = s_cviol[tid]; }
            } else if (live) {
                env_reward_mse<SYS, R, R>(h, s_xg, state, cx.noisy, cx.clipped, e.ctrl_step, rew, mse);
                done = env_done<SYS, R, R>(h, s_xg, state, goal_reached, oob);
                if (nc > 0) cviol = env_constraints(T, h->rows, nc, h->con_fast, s_z + tid, E, nullptr, P.out.d_cvals ? P.out.d_cvals + (tn0 + tid) * nc : nullptr);
            }
            PROF_T(k4);
            if (live) {
                const int step_before = e.ctrl_step;
                e.ctrl_step += 1; e.total_steps += 1;
                env_finalize<R>(h, e.ctrl_step, cviol, rew, done, truncated);
                info.done = done; info.truncated = truncated; info.cviol = cviol;
                e.ep_ret += (float)rew; e.ep_len += 1.f; e.ep_cv += (float)cviol;
                const bool helper_obs = SPLIT && !obs_noise;               // the helper already wrote the continuing observation
                if (done) {
                    if (truncated) {                                       // ppo.py:274-283: critic on the terminal observation
                        trunc = 1;
                        if (helper_obs) { for (int k = 0; k < od; ++k) s_tobs[tile_idx<E>(k, tid)] = s_obs[tile_idx<E>(k, tid)]; }
                        else make_obs<SYS, R, R>(T, h, s_xg, state, n, step_before, false, ns, TileObsSink<E>{s_tobs, tid});
                    }
                    ep_count += 1; ep_ret_sum += e.ep_ret; ep_len_sum += e.ep_len; ep_cv_sum += e.ep_cv;
                    e.ep_ret = 0.f; e.ep_len = 0.f; e.ep_cv = 0.f;
                    env_reset<SYS, R>(T, h, n, e, state);                  // dummy_vec_env.py:33-38
                    ns = reset_noise<R>(T, h, n, e.episode);
                    make_obs<SYS, R, R>(T, h, s_xg, state, n, 0, true, ns, TileObsSink<E>{s_obs, tid});
                } else if (!helper_obs) {
                    make_obs<SYS, R, R>(T, h, s_xg, state, n, step_before, false, ns, TileObsSink<E>{s_obs, tid});
                }
                const size_t tn = tn0 + tid;
                if (A == 2) *reinterpret_cast<float2*>(P.out.d_act + tn * A) = make_float2(act[0], act[1 % A]);
                else if (A == 4) *reinterpret_cast<float4*>(P.out.d_act + tn * A) = make_float4(act[0], act[1 % A], act[2 % A], act[3 % A]);
                else P.out.d_act[tn * A] = act[0];
                P.out.d_rew[tn] = (float)rew;
                P.out.d_mask[tn] = done ? 0.f : 1.f;
                P.out.d_v[tn] = val;
                P.out.d_logp[tn] = logp;
                if (P.out.d_cviol) P.out.d_cviol[tn] = (float)cviol;
                if (P.out.d_mse) P.out.d_mse[tn] = (float)mse;
                if (!trunc) P.out.d_terminal_v[tn] = 0.f;
            }
#ifdef SCG_PROF
            if (PROF_ON) { long long k5 = clock64(); pacc[0] += k1 - k0; pacc[1] += k2 - k1; pacc[2] += k3 - k2; pacc[3] += k4 - k3; pacc[4] += k5 - k4; }
#endif
        } else if (SPLIT && warp <= 3) {
            named_bar_sync(1, 128);                                        // z, ctrl_step published by the owners
            const int ee = lane;
            const bool hl = ee < E && n0 + ee < N;
            if (hl) {
                if (warp == 1) {                                           // constraints (benchmark_env.py:472-485)
                    int cv = 0;
                    if (nc > 0) cv = env_constraints(T, h->rows, nc, h->con_fast, s_z + ee, E, nullptr, P.out.d_cvals ? P.out.d_cvals + (tn0 + ee) * nc : nullptr);
                    s_cviol[ee] = cv;
                } else {
                    R st[NX];
#pragma unroll
                    for (int k = 0; k < NX; ++k) st[k] = (R)s_z[k * E + ee];
                    if (warp == 2) {                                       // reward + info['mse']
                        R ny[A], cl[A], rw, ms;
#pragma unroll
                        for (int k = 0; k < A; ++k) { ny[k] = (R)s_z[(NX + k) * E + ee]; cl[k] = (R)s_z[(NX + A + k) * E + ee]; }
                        env_reward_mse<SYS, R, R>(h, s_xg, st, ny, cl, s_cstep[ee], rw, ms);
                        s_rew[ee] = rw; s_mse[ee] = ms;
                    } else if (!obs_noise) {                               // next observation of the continuing episode
                        NoiseSrc<R> nn; nn.row = nullptr; nn.philox = false; nn.slot = 0; nn.stream = 0; nn.idx = 0; nn.key.k0 = nn.key.k1 = nn.key.genv = 0;
                        make_obs<SYS, R, R>(T, h, s_xg, st, n0 + ee, s_cstep[ee], false, nn, TileObsSink<E>{s_obs, ee});
                    }
                }
            }
            named_bar_arrive(2, 128);
        }
        PROF_T(k6);
        // (c) critic on the terminal observation of time-limit truncated envs (rare, CTA-uniform branch).
        //     s_tobs aliases s_h2: written after layer 3 consumed h2, read by layer 1 of this pass, overwritten by its layer 2.
        const int any_trunc = __syncthreads_or(trunc);
#ifdef SCG_PROF
        if (PROF_ON) pacc[5] += clock64() - k6;
#endif
        if (any_trunc) {
            mlp_forward<HID, E, NT, A>(s_blob, L, s_tobs, s_h1, s_h2, s_out, od, P.activation);
            if (live && trunc) P.out.d_terminal_v[tn0 + tid] = s_out[A * E + tid];
            __syncthreads();
        }
    }
#ifdef SCG_PROF
    if (PROF_ON && blockIdx.x == 0) for (int i = 0; i < 8; ++i) P.prof[i] = pacc[i];
    long long g0 = clock64();
#endif
    // ---- bootstrap value of the last observation, state write-back, GAE
    mlp_forward<HID, E, NT, A>(s_blob, L, s_obs, s_h1, s_h2, s_out, od, P.activation);
    if (live) {
        env_store<SYS, R>(T, n, e);
        for (int k = 0; k < od; ++k) {
            const float v = s_obs[tile_idx<E>(k, tid)];
            T.st.cur_obs[(size_t)n * od + k] = v;
            if (P.out.d_last_obs) P.out.d_last_obs[(size_t)n * od + k] = v;
        }
        const float lval = s_out[A * E + tid];
        if (P.out.d_last_val) P.out.d_last_val[n] = lval;
        if (P.out.d_ep_count) { P.out.d_ep_count[n] = ep_count; P.out.d_ep_return_sum[n] = ep_ret_sum; P.out.d_ep_length_sum[n] = ep_len_sum; P.out.d_ep_cviol_sum[n] = ep_cv_sum; }
        if (P.out.d_ret && P.out.d_adv) {
            // compute_returns_and_advantages (ppo_utils.py:374-400): float32 ret / td recursions, float64 GAE accumulator
            float ret = lval, vnext = lval; double adv = 0.0;
            constexpr int CH = 8;            // loads of CH steps are issued together: the scan itself is a serial chain
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
                        P.out.d_rew[tn] = r;                               // rews += gamma * terminal_v, in place (ppo_utils.py:389)
                        const float m = mm[j], v = vv[j];
                        ret = r + P.gamma * m * ret;
                        if (!P.use_gae) adv = (double)(ret - v);
                        else { const float td = r + P.gamma * m * vnext - v; adv = adv * P.lam_d * P.gamma_d * (double)m + (double)td; }
                        P.out.d_ret[tn] = ret; P.out.d_adv[tn] = (float)adv; vnext = v;
                    }
                }
            }
        }
    }
#ifdef SCG_PROF
    if (PROF_ON && blockIdx.x == 0) P.prof[7] = clock64() - g0;
#endif
}

// Philox known-answer hook
__global__ void philox_selftest_kernel(const uint32_t* ctr, const uint32_t* key, uint32_t* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t o[4];
    philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], key[2 * i], key[2 * i + 1], o);
    out[4 * i] = o[0]; out[4 * i + 1] = o[1]; out[4 * i + 2] = o[2]; out[4 * i + 3] = o[3];
}
