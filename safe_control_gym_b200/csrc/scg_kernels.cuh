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
// two tanh at once: odd rational minimax x P(x^2) / Q(x^2) (degrees 13 / 6) on |x| <= 7.905 (tanh rounds to +-1 in float beyond),
// |rel err| < 5e-7.  Horner steps are packed FP32 ops (FFMA2 / FMUL2: one instruction per step for both values), one MUFU.RCP per
// value, no exponential: 11 instructions per value against ~28 for the exp-based form.
__device__ __forceinline__ float2 tanh_rat2(float2 v) {
    const float lim = 7.90531110763549805f;
    const float2 x = make_float2(fminf(fmaxf(v.x, -lim), lim), fminf(fmaxf(v.y, -lim), lim));
    const float2 x2 = __fmul2_rn(x, x);
#define SCG_C2(c) make_float2(c, c)
    float2 p = __ffma2_rn(x2, SCG_C2(-2.76076847742355e-16f), SCG_C2(2.00018790482477e-13f));
    p = __ffma2_rn(x2, p, SCG_C2(-8.60467152213735e-11f));
    p = __ffma2_rn(x2, p, SCG_C2(5.12229709037114e-08f));
    p = __ffma2_rn(x2, p, SCG_C2(1.48572235717979e-05f));
    p = __ffma2_rn(x2, p, SCG_C2(6.37261928875436e-04f));
    p = __ffma2_rn(x2, p, SCG_C2(4.89352455891786e-03f));
    p = __fmul2_rn(p, x);
    float2 q = __ffma2_rn(x2, SCG_C2(1.19825839466702e-06f), SCG_C2(1.18534705686654e-04f));
    q = __ffma2_rn(x2, q, SCG_C2(2.26843463243900e-03f));
    q = __ffma2_rn(x2, q, SCG_C2(4.89352518554385e-03f));
#undef SCG_C2
    return make_float2(__fdividef(p.x, q.x), __fdividef(p.y, q.y));
}
__device__ __forceinline__ void tanh_fast4(float& x0, float& x1, float& x2, float& x3) {
    const float2 a = tanh_rat2(make_float2(x0, x1)), b = tanh_rat2(make_float2(x2, x3));
    x0 = a.x; x1 = a.y; x2 = b.x; x3 = b.y;
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
    int blob, obs, tobs, h1, out, xg, total_floats;
    // xg_floats: size of the shared X_GOAL copy in float units (L * nx * sizeof(R) / 4); nw: GEMM warps per network
    __host__ __device__ RolloutSmem(int obs_dim, int act_dim, int xg_floats, int nw) {
        BlobLayout L = blob_layout(obs_dim, act_dim, HID);
        blob = 0;
        obs = blob + L.total;
        h1 = obs + 2 * obs_dim * E;          // two observation tiles (ping-pong): tile t+1 is built while tile t is still the policy input
        tobs = h1 + 2 * HID * E;             // terminal observations of time-limit truncated envs
        out = tobs + obs_dim * E;
        xg = out + pad4i(nw * (act_dim + 1) * E);   // per-warp partial sums of the output layer
        total_floats = xg + pad4i(xg_floats);
    }
};

// named barriers: 1, 2 = owner warp <-> helper warps of the env-step section
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// Forward of ONE network (net 0 = actor with n_out = A outputs, net 1 = critic with 1 output) for the CTA's E envs by a group
// of NTN threads (tl = index inside the group); both groups run in lockstep (__syncthreads).
//   s_in  : pair-swizzled observation tile (tile_idx)          s_h1n : this net's hidden tile
//   s_outp: [NTN/32 warps][A+1 outputs][E] per-warp partial sums of the output layer (see mlp_out)
// Layers 1 and 2 are register-tiled FP32 GEMMs (4 envs x TU units per thread) out of shared memory with packed FFMA2.
// Layer 3 is fused into the epilogue of layer 2: every thread multiplies its 4 x TU activations with the matching slice of the
// output rows, the unit groups of a warp are summed with shuffles and the per-warp partial sums go to shared memory -- the second
// hidden tile is never stored and the whole forward needs two barriers.
template <int HID, int E, int NTN, int AMAX>
__device__ __forceinline__ void net_forward(int tl, int net, int n_out, int w3_row0, const float* __restrict__ s_blob, const BlobLayout L,
                                            const float* __restrict__ s_in, float* __restrict__ s_h1n, float* __restrict__ s_outp,
                                            int obs_dim, int act) {
    constexpr int EG = E / 4;              // env groups
    constexpr int UG = NTN / EG;           // unit groups
    constexpr int TU = HID / UG;           // units per thread
    static_assert(TU == 2 || TU == 4, "unsupported tile");
    // lane -> (env group, unit group).  Measured on B200 (tools/ubench/lds_patterns.cu): an LDS.128 takes 2 shared-memory wavefronts
    // when every quad of 4 adjacent lanes touches at most 2 distinct 16-byte chunks and 4 otherwise.  Lane bit 0 therefore selects
    // the unit group, bits 1.. the env group, the remaining high bits the unit group again: both the activation and the weight
    // loads of the GEMM loops are on the 2-wavefront path.
    constexpr int EGB = (EG == 16) ? 4 : (EG == 8) ? 3 : (EG == 4) ? 2 : (EG == 2) ? 1 : 0;
    static_assert((1 << EGB) == EG && EG <= 16, "E must be 4, 8, 16, 32 or 64");
    const int ln = tl & 31;
    const int eg = (ln >> 1) & (EG - 1);
    const int ug = (tl >> 5) * (32 / EG) + ((ln & 1) | ((ln >> (1 + EGB)) << 1));
    const int ul0 = ug * TU;

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
    auto init_acc = [&](const float* __restrict__ bias) {
#pragma unroll
        for (int j = 0; j < TU; ++j) { const float b = bias[j]; acc[0][j] = make_float2(b, 0.f); acc[1][j] = acc[0][j]; acc[2][j] = acc[0][j]; acc[3][j] = acc[0][j]; }
    };
    float h[4][TU];
    auto activate = [&]() {                        // reduce the (even, odd) pairs and apply the hidden activation
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < TU; ++j) h[i][j] = acc[i][j].x + acc[i][j].y;
        apply_hidden_act<TU>(h, act);
    };
    // ---- layer 1: [E x obs] . [obs x H]      (W1P [kp][2H][2], this net's units start at net*H)
    SCG_TK(14);
    init_acc(s_blob + L.b1 + net * HID + ul0);
    gemm(s_in + eg * 4, 2 * EG * 4, s_blob + L.w1t + (net * HID + ul0) * 2, 2 * HID * 2, obs_dim >> 1);
    SCG_TK(15);
    activate();
#pragma unroll
    for (int j = 0; j < TU; j += 2) {              // store pair-swizzled as the input of layer 2
        float* d = s_h1n + ((((ul0 + j) >> 1) * 2) * EG + eg) * 4;
        *reinterpret_cast<float4*>(d) = make_float4(h[0][j], h[0][j + 1], h[1][j], h[1][j + 1]);
        *reinterpret_cast<float4*>(d + EG * 4) = make_float4(h[2][j], h[2][j + 1], h[3][j], h[3][j + 1]);
    }
    SCG_TK(16);
    __syncthreads();
    SCG_TK(17);
    // ---- layer 2: [E x H] . [H x H]          (W2P [net][kp][H][2])
    init_acc(s_blob + L.b2 + net * HID + ul0);
    gemm(s_h1n + eg * 4, 2 * EG * 4, s_blob + L.w2t + net * HID * HID + ul0 * 2, HID * 2, HID / 2);
    SCG_TK(18);
    activate();
    SCG_TK(19);
    // ---- layer 3 on the register tile: partial dot products over this thread's TU units, summed over the warp's unit groups
    {
        float part[AMAX][4];
#pragma unroll
        for (int o = 0; o < AMAX; ++o) {
            if (o < n_out) {                       // warp-uniform
                const float* w3 = s_blob + L.w3 + (w3_row0 + o) * HID + ul0;
                float w[TU];
                if (TU == 4) { const float4 t = *reinterpret_cast<const float4*>(w3); w[0] = t.x; w[1] = t.y; w[2 % TU] = t.z; w[3 % TU] = t.w; }
                else { const float2 t = *reinterpret_cast<const float2*>(w3); w[0] = t.x; w[1] = t.y; }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float v = h[i][0] * w[0];
#pragma unroll
                    for (int j = 1; j < TU; ++j) v = fmaf(h[i][j], w[j], v);
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
#pragma unroll
                    for (int b = 1 + EGB; b < 5; ++b) v += __shfl_xor_sync(0xffffffffu, v, 1 << b);
                    part[o][i] = v;
                }
            }
        }
        if ((ln & 1) == 0 && (ln >> (1 + EGB)) == 0) {     // one lane per env group holds the warp's sums
#pragma unroll
            for (int o = 0; o < AMAX; ++o)
                if (o < n_out)
                    *reinterpret_cast<float4*>(s_outp + ((tl >> 5) * (AMAX + 1) + w3_row0 + o) * E + eg * 4) = make_float4(part[o][0], part[o][1], part[o][2], part[o][3]);
        }
    }
    SCG_TK(21);
    __syncthreads();
    SCG_TK(22);
}

// The lower half of the CTA computes the actor, the upper half the critic, in lockstep.
//   s_outp : [NT/64 warps][A+1][E] partial sums; read the outputs with mlp_out()
template <int HID, int E, int NT, int A>
__device__ __noinline__ void mlp_forward(const float* __restrict__ s_blob, const BlobLayout L, const float* __restrict__ s_in, float* __restrict__ s_h1,
                                         float* __restrict__ s_outp, int obs_dim, int act) {
    constexpr int NTN = NT / 2;
    constexpr int NETSZ = (HID / 2) * (2 * (E / 4) * 4);       // floats of one net's hidden tile
    const int net = threadIdx.x / NTN, tl = threadIdx.x % NTN;
    net_forward<HID, E, NTN, A>(tl, net, net ? 1 : A, net ? A : 0, s_blob, L, s_in, s_h1 + net * NETSZ, s_outp, obs_dim, act);
}
// output o (0..A-1 action means, A = value) of env e: bias + the per-warp partial sums, in a fixed order
template <int E, int NW, int A>
__device__ __forceinline__ float mlp_out(const float* __restrict__ s_outp, const float* __restrict__ s_blob, const BlobLayout& L, int o, int e) {
    float v = s_blob[L.b3 + o];
#pragma unroll
    for (int w = 0; w < NW; ++w) v += s_outp[(w * (A + 1) + o) * E + e];
    return v;
}

#ifdef SCG_PROF
#define PROF_T(v) long long v = clock64()
#else
#define PROF_T(v) do { } while (0)
#endif

// CTAs per SM the register allocation is capped for: the 16-thread-per-env variants (E=16 x 256 threads, E=8 x 128 threads) halve
// the per-step MLP latency of a CTA and rely on TWO (or more) co-resident CTAs per SM to fill the machine.
__host__ __device__ constexpr int rollout_min_blocks(int HID, int E, int NT) {
    return (NT / E >= 16 && HID <= 64) ? (NT == 256 ? 2 : 3) : 1;
}
template <int SYS, typename R, int HID, int E, int NT>
__global__ void __launch_bounds__(NT, rollout_min_blocks(HID, E, NT)) rollout_kernel(DevTables T, RolloutParams P) {
    constexpr int NX = SysDims<SYS>::NX, A = SysDims<SYS>::NU;
    constexpr int NZ = NX + 2 * A;             // z scratch rows: state, noisy action, clipped action
    // With >= 4 warps the env-step section is split: warp 0 = owners (sampling, action map, physics, done, finalize, reset),
    // warp 1 = constraints, warp 2 = reward + mse, warp 3 = next observation.  Smaller CTAs keep everything on the owner lanes.
    constexpr bool SPLIT = (NT >= 128) && (E <= 32);
    constexpr int NW = NT / 64;                // GEMM warps per network
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ HotCfg<R> s_hot;
    __shared__ double s_z[NZ * E];
    __shared__ float s_eps[A * E];
    __shared__ uint32_t s_tsteps0[E];
    __shared__ int s_cstep[E];                 // ctrl_step BEFORE the increment of the step being processed
    __shared__ int s_cviol[E];
    __shared__ R s_rew[E], s_mse[E];
    // reset candidates: the helper warp pre-draws the next two episodes of every env (slot = episode & 1)
    constexpr int NC = SCG_MAX_INERTIAL + NX;
    __shared__ double s_cand[2 * NC * E];
    __shared__ int s_cand_ep[2 * E];           // episode index held by each candidate slot (-1 = empty)
    __shared__ int s_epi[E];                   // next episode index of each env, published by its owner lane
    __shared__ int s_flag[E], s_hit[E];        // FAST path: done / truncated bits of the step, "helper built the reset observation"

    const int od = T.obs_dim, N = T.N, nc = T.n_con;
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * E;
    const RolloutSmem<HID, E> S(od, A, T.L * NX * (int)(sizeof(R) / 4), NW);
    const BlobLayout L = blob_layout(od, A, HID);
    float* s_blob = smem + S.blob;
    float* s_obs = smem + S.obs;
    float* s_tobs = smem + S.tobs;
    float* s_h1 = smem + S.h1;
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
    // roles: owner lanes sit in the LAST warp (the issue arbiter favours the highest warp id -- the env step is the critical
    // path), helpers are the three warps below it
    const int warp = tid >> 5, lane = tid & 31;
    const int role = (NT / 32 - 1) - warp;          // 0 = owner warp, 1..3 = helper warps
    const int ol = tid - (NT - 32);                 // owner lane index (env inside the CTA) when role == 0
    const bool owner = ol >= 0 && ol < E;
    const int n = n0 + ol;
    const bool live = owner && n < N;
    EnvRegs<SYS, R> e;
    if (live) env_load<SYS, R>(T, n, e);
    if (owner) { s_tsteps0[ol] = live ? e.total_steps : 0u; s_epi[ol] = live ? e.episode : 0; s_cand_ep[ol] = -1; s_cand_ep[E + ol] = -1; }
    const bool use_cand = SPLIT && h->n_off_draws == 0;
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
    const bool fast_done = use_cand && !obs_noise && !h->done_on_violation;
    // obs write-out: element i = tid + j*NT of the [E][od] record maps to (env i / od, dim i % od); advance incrementally
    const int ow_k = tid % od, ow_e = tid / od, ow_dk = NT % od, ow_de = NT / od;
#ifdef SCG_PROF
    const long long cta_t0 = clock64();
    if (tid == NT - 32) scg_tk_last = cta_t0;
    long long pacc[8]; for (int i = 0; i < 8; ++i) pacc[i] = 0;
#define PROF_ON (P.prof != nullptr && tid == NT - 32)
#endif

    int cur = 0;
    const int OBS_TILE = od * E;
    for (int t = 0; t < P.T; ++t) {
        const size_t tn0 = (size_t)t * N + n0;
        float* s_oc = s_obs + cur * OBS_TILE;            // observation of step t (policy input)
        float* s_on = s_obs + (cur ^ 1) * OBS_TILE;      // observation of step t+1 (being built)
        PROF_T(k0);
        // standard-normal draws of this step: table mode -> owner lanes issue the loads now (latency hidden by the MLP);
        // Philox mode -> the FIRST lane group generates them into shared memory, off the owners' critical path
        float z[A];
#pragma unroll
        for (int k = 0; k < A; ++k) z[k] = 0.f;
        if (!P.deterministic) {
            if (!philox_eps) {
                if (live) {
#pragma unroll
                    for (int k = 0; k < A; ++k) z[k] = __ldg(P.eps + (tn0 + ol) * A + k);
                }
            } else if (tid < E) {
                const int ee = tid;
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
                dst[i] = s_oc[tile_idx<E>(k, ee)];
                k += ow_dk; ee += ow_de;
                if (k >= od) { k -= od; ee += 1; }
            }
        }
        PROF_T(k1);
        SCG_TK(13);
        mlp_forward<HID, E, NT, A>(s_blob, L, s_oc, s_h1, s_out, od, P.activation);
        PROF_T(k2);

        // (b) env step.  Owner lanes: sample, action map, physics, termination test; then reward / constraints / next observation
        // run on helper warps (SPLIT); owners finalize, emit the record and auto-reset.  FAST path (no observation noise, no
        // done_on_violation, pre-drawn reset candidates): the owner publishes done / truncated right after the physics, so the
        // observation helper also builds the terminal and the RESET observation -- the owner only reloads its registers on a reset.
        int trunc = 0;
        if (role == 0) {
            SCG_TK(0);
            float act[A], logp = 0.f, val = 0.f;
            R state[NX], rew = (R)0, mse = (R)0;
            StepCtx<SYS, R> cx; NoiseSrc<R> ns;
            int done = 0, goal_reached = 0, oob = 0, cviol = 0, truncated = 0, fl = 0;
            if (live) {
#pragma unroll
                for (int k = 0; k < A; ++k) {
                    const float zk = (philox_eps && !P.deterministic) ? s_eps[k * E + ol] : z[k];
                    const float mu = mlp_out<E, NW, A>(s_out, s_blob, L, k, ol);
                    act[k] = fmaf(sd[k], zk, mu);                         // Normal.sample (ppo_utils.py:228)
                    const float d = act[k] - mu;
                    logp += -(d * d) * inv2var[k] - lsd[k] - log_sqrt_2pi;   // distributions.py:12-21
                }
                val = mlp_out<E, NW, A>(s_out, s_blob, L, A, ol);
                R action[A];
#pragma unroll
                for (int k = 0; k < A; ++k) action[k] = (R)act[k];
                const double* nrow = (!h->philox && T.noise_table && T.D > 0)
                                         ? T.noise_table + ((size_t)((P.noise_cursor + t) % T.Tn) * N + n) * T.D : nullptr;
                env_pre<SYS, R>(T, h, n, e, action, nrow, cx, ns);
                SCG_TK(1);
                env_physics<SYS, R>(h, e, cx, state);
                SCG_TK(6);
                // publish z = [state; noisy action; clipped action] and the step counter for the helpers / the constraint rows
#pragma unroll
                for (int k = 0; k < NX; ++k) s_z[k * E + ol] = (double)state[k];
#pragma unroll
                for (int k = 0; k < A; ++k) { s_z[(NX + k) * E + ol] = (double)cx.noisy[k]; s_z[(NX + A + k) * E + ol] = (double)cx.clipped[k]; }
                s_cstep[ol] = e.ctrl_step;
                if (fast_done) {
                    done = env_done<SYS, R, R>(h, s_xg, state, goal_reached, oob);
                    fl = (e.ctrl_step + 1 >= h->ctrl_steps) ? (done ? 1 : 3) : done;   // bit 0 done, bit 1 truncated
                    s_flag[ol] = fl;
                }
            }
            PROF_T(k3);
            SCG_TK(7);
            const size_t tn = tn0 + ol;
            const int step_before = e.ctrl_step;
            bool publish_epi = false;
            auto store_policy_part = [&]() {                               // record entries that do not depend on the helpers
                if (A == 2) *reinterpret_cast<float2*>(P.out.d_act + tn * A) = make_float2(act[0], act[1 % A]);
                else if (A == 4) *reinterpret_cast<float4*>(P.out.d_act + tn * A) = make_float4(act[0], act[1 % A], act[2 % A], act[3 % A]);
                else P.out.d_act[tn * A] = act[0];
                P.out.d_v[tn] = val;
                P.out.d_logp[tn] = logp;
            };
            auto auto_reset = [&](bool hit, bool write_obs) {              // dummy_vec_env.py:33-38
                const int slot = e.episode & 1;
                env_reset<SYS, R>(T, h, n, e, state, hit ? s_cand + slot * NC * E + ol : nullptr, E);
                if (write_obs) {
                    ns = reset_noise<R>(T, h, n, e.episode);
                    make_obs<SYS, R, R>(T, h, s_xg, state, n, 0, true, ns, TileObsSink<E>{s_on, ol});
                }
            };
            if (SPLIT) {
                named_bar_arrive(1, 128);                                  // helpers may start
                if (live) {
                    if (!fast_done) done = env_done<SYS, R, R>(h, s_xg, state, goal_reached, oob);
                    else {
                        // FAST: everything that does not need the helpers' results happens while they work
                        store_policy_part();
                        P.out.d_mask[tn] = (fl & 1) ? 0.f : 1.f;
                        if (fl != 3) P.out.d_terminal_v[tn] = 0.f;
                        if (fl & 1) {
                            // take the pre-drawn episode if the helper has it (same test as the observation helper), draw inline otherwise;
                            // the new episode index is published after the helpers are done (they still read the old one)
                            const bool hit = *(volatile int*)&s_cand_ep[(e.episode & 1) * E + ol] == e.episode;
                            auto_reset(hit, !hit);
                            publish_epi = true;
                        } else e.ctrl_step += 1;
                    }
                }
                SCG_TK(8);
                named_bar_sync(2, 128);                                    // helpers are done
                SCG_TK(9);
                if (live) { rew = s_rew[ol]; mse = s_mse[ol]; cviol = s_cviol[ol]; }
            } else if (live) {
                env_reward_mse<SYS, R, R>(h, s_xg, state, cx.noisy, cx.clipped, e.ctrl_step, rew, mse);
                done = env_done<SYS, R, R>(h, s_xg, state, goal_reached, oob);
                if (nc > 0) cviol = env_constraints(T, h->rows, nc, h->con_fast, s_z + ol, E, nullptr, P.out.d_cvals ? P.out.d_cvals + (tn0 + ol) * nc : nullptr);
            }
            PROF_T(k4);
            if (live) {
                e.total_steps += 1;
                env_finalize<R>(h, step_before + 1, cviol, rew, done, truncated);   // FAST: reproduces the done / truncated bits of `fl`
                SCG_TK(10);
                e.ep_ret += (float)rew; e.ep_len += 1.f; e.ep_cv += (float)cviol;
                if (done) {
                    ep_count += 1; ep_ret_sum += e.ep_ret; ep_len_sum += e.ep_len; ep_cv_sum += e.ep_cv;
                    e.ep_ret = 0.f; e.ep_len = 0.f; e.ep_cv = 0.f;
                }
                if (fast_done) {
                    trunc = truncated;
                    if (publish_epi) *(volatile int*)&s_epi[ol] = e.episode;
                } else {
                    e.ctrl_step += 1;
                    const bool helper_obs = SPLIT && !obs_noise;           // the helper already wrote the continuing observation
                    if (done) {
                        if (truncated) {                                   // ppo.py:274-283: critic on the terminal observation
                            trunc = 1;
                            if (helper_obs) { for (int k = 0; k < od; ++k) s_tobs[tile_idx<E>(k, ol)] = s_on[tile_idx<E>(k, ol)]; }
                            else make_obs<SYS, R, R>(T, h, s_xg, state, n, step_before, false, ns, TileObsSink<E>{s_tobs, ol});
                        }
                        const bool hit = use_cand && (*(volatile int*)&s_cand_ep[(e.episode & 1) * E + ol] == e.episode);
                        auto_reset(hit, true);
                        __threadfence_block();                             // candidate fully read before the new index is published
                        *(volatile int*)&s_epi[ol] = e.episode;
                    } else if (!helper_obs) {
                        make_obs<SYS, R, R>(T, h, s_xg, state, n, step_before, false, ns, TileObsSink<E>{s_on, ol});
                    }
                    store_policy_part();
                    P.out.d_mask[tn] = done ? 0.f : 1.f;
                    if (!trunc) P.out.d_terminal_v[tn] = 0.f;
                }
                SCG_TK(11);
                P.out.d_rew[tn] = (float)rew;
                if (P.out.d_cviol) P.out.d_cviol[tn] = (float)cviol;
                if (P.out.d_mse) P.out.d_mse[tn] = (float)mse;
            }
            SCG_TK(12);
#ifdef SCG_PROF
            if (PROF_ON) { long long k5 = clock64(); pacc[0] += k1 - k0; pacc[1] += k2 - k1; pacc[2] += k3 - k2; pacc[3] += k4 - k3; pacc[4] += k5 - k4; }
#endif
        } else if (SPLIT && role <= 3) {
            named_bar_sync(1, 128);                                        // z, ctrl_step (and the FAST done flags) published by the owners
            const int ee = lane;
            const bool hl = ee < E && n0 + ee < N;
            if (hl) {
                if (role == 1) {                                           // constraints (benchmark_env.py:472-485)
                    int cv = 0;
                    if (nc > 0) cv = env_constraints(T, h->rows, nc, h->con_fast, s_z + ee, E, nullptr, P.out.d_cvals ? P.out.d_cvals + (tn0 + ee) * nc : nullptr);
                    s_cviol[ee] = cv;
                } else {
                    R st[NX];
#pragma unroll
                    for (int k = 0; k < NX; ++k) st[k] = (R)s_z[k * E + ee];
                    if (role == 2) {                                       // reward + info['mse']
                        R ny[A], cl[A], rw, ms;
#pragma unroll
                        for (int k = 0; k < A; ++k) { ny[k] = (R)s_z[(NX + k) * E + ee]; cl[k] = (R)s_z[(NX + A + k) * E + ee]; }
                        env_reward_mse<SYS, R, R>(h, s_xg, st, ny, cl, s_cstep[ee], rw, ms);
                        s_rew[ee] = rw; s_mse[ee] = ms;
                    } else if (!obs_noise) {                               // next observation
                        NoiseSrc<R> nn; nn.row = nullptr; nn.philox = false; nn.slot = 0; nn.stream = 0; nn.idx = 0; nn.key.k0 = nn.key.k1 = nn.key.genv = 0;
                        const int fl = fast_done ? s_flag[ee] : 0;
                        if (!(fl & 1)) {                                   // continuing episode (or the slow path: the owner fixes up done envs)
                            make_obs<SYS, R, R>(T, h, s_xg, st, n0 + ee, s_cstep[ee], false, nn, TileObsSink<E>{s_on, ee});
                        } else {
                            if (fl & 2) make_obs<SYS, R, R>(T, h, s_xg, st, n0 + ee, s_cstep[ee], false, nn, TileObsSink<E>{s_tobs, ee});   // terminal observation
                            const int ep = *(volatile int*)&s_epi[ee], slot = ep & 1;          // the episode about to start
                            int hit = 0;
                            if (*(volatile int*)&s_cand_ep[slot * E + ee] == ep) {
                                const double* cd = s_cand + slot * NC * E + ee;
                                R initr[NX];
#pragma unroll
                                for (int j = 0; j < NX; ++j) initr[j] = (R)cd[(SCG_MAX_INERTIAL + j) * E];
                                EnvRegs<SYS, R> tmp; R st0[NX];
                                Physics<SYS, R>::set_init(tmp, initr);
                                Physics<SYS, R>::state(tmp, st0);
                                make_obs<SYS, R, R>(T, h, s_xg, st0, n0 + ee, 0, true, nn, TileObsSink<E>{s_on, ee});
                                hit = 1;
                            }
                            s_hit[ee] = hit;
                        }
                    }
                }
            }
            named_bar_arrive(2, 128);
            if (use_cand && role == 3 && hl) {
                // off the critical path: keep the next two episodes of env ee pre-drawn (a pure function of (env, episode))
                const int want = *(volatile int*)&s_epi[ee];
#pragma unroll 1
                for (int d = 0; d < 2; ++d) {
                    const int ep = want + d, slot = ep & 1;
                    if (*(volatile int*)&s_cand_ep[slot * E + ee] != ep) {
                        double inr[SCG_MAX_INERTIAL], ini[NX];
                        reset_draw<SYS, R>(T, h, n0 + ee, ep, inr, ini);
                        double* cd = s_cand + slot * NC * E + ee;
#pragma unroll
                        for (int j = 0; j < SCG_MAX_INERTIAL; ++j) cd[j * E] = inr[j];
#pragma unroll
                        for (int j = 0; j < NX; ++j) cd[(SCG_MAX_INERTIAL + j) * E] = ini[j];
                        __threadfence_block();
                        *(volatile int*)&s_cand_ep[slot * E + ee] = ep;
                    }
                }
            }
        }
        SCG_TK(23);
        PROF_T(k6);
        // (c) critic on the terminal observation of time-limit truncated envs (rare, CTA-uniform branch).
        const int any_trunc = __syncthreads_or(trunc);
#ifdef SCG_PROF
        if (PROF_ON) pacc[5] += clock64() - k6;
#endif
        if (any_trunc) {
            mlp_forward<HID, E, NT, A>(s_blob, L, s_tobs, s_h1, s_out, od, P.activation);
            if (live && trunc) P.out.d_terminal_v[tn0 + ol] = mlp_out<E, NW, A>(s_out, s_blob, L, A, ol);
            __syncthreads();
        }
        cur ^= 1;
        SCG_TK(24);
    }
    float* s_ol = s_obs + cur * OBS_TILE;                  // the observation after the last step
#ifdef SCG_PROF
    if (PROF_ON && blockIdx.x == 0) for (int i = 0; i < 8; ++i) P.prof[i] = pacc[i];
    long long g0 = clock64();
#endif
    // ---- bootstrap value of the last observation, state write-back, GAE
    mlp_forward<HID, E, NT, A>(s_blob, L, s_ol, s_h1, s_out, od, P.activation);
    if (live) {
        env_store<SYS, R>(T, n, e);
        for (int k = 0; k < od; ++k) {
            const float v = s_ol[tile_idx<E>(k, ol)];
            T.st.cur_obs[(size_t)n * od + k] = v;
            if (P.out.d_last_obs) P.out.d_last_obs[(size_t)n * od + k] = v;
        }
        const float lval = mlp_out<E, NW, A>(s_out, s_blob, L, A, ol);
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
    if (PROF_ON && blockIdx.x < 1024) { P.prof[16 + blockIdx.x] = g0 - cta_t0; P.prof[16 + 1024 + blockIdx.x] = ep_count; }
#endif
}

// Philox known-answer hook
static __global__ void philox_selftest_kernel(const uint32_t* ctr, const uint32_t* key, uint32_t* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t o[4];
    philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], key[2 * i], key[2 * i + 1], o);
    out[4 * i] = o[0]; out[4 * i + 1] = o[1]; out[4 * i + 2] = o[2]; out[4 * i + 3] = o[3];
}
