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
//   W1T [OBS][2H]   layer-1 weights, k-major, actor units 0..H-1 then critic units H..2H-1
//   b1  [2H]
//   W2T [2][H][H]   layer-2 weights per net, k-major
//   b2  [2H]
//   W3  [A+1][H]    A actor output rows then the critic row
//   b3  [pad4(A+1)]
//   logstd [pad4(A)]
// every section is a multiple of 4 floats so that the whole blob can be staged with one TMA bulk copy.
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
template <int ACT> __device__ __forceinline__ float hidden_act(float x) {
    if (ACT == SCG_ACT_TANH) return tanh_fast(x);
    if (ACT == SCG_ACT_RELU) return fmaxf(x, 0.0f);
    return x > 0.0f ? x : 0.01f * x;
}
__device__ __forceinline__ float hidden_act_rt(float x, int act) {
    if (act == SCG_ACT_TANH) return tanh_fast(x);
    if (act == SCG_ACT_RELU) return fmaxf(x, 0.0f);
    return x > 0.0f ? x : 0.01f * x;
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
struct SmemObsSink {        // column of the CTA's [obs_dim][EP] float tile
    float* col; int stride;
    template <typename R> __device__ __forceinline__ void operator()(int k, R v) const { col[k * stride] = (float)v; }
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
    env_reset<SYS, R>(T, n, e, state);
    e.ep_ret = 0.f; e.ep_len = 0.f; e.ep_cv = 0.f;
    T.st.initial_reset[n] = 1;
    env_store<SYS, R>(T, n, e);
    NoiseSrc<R> ns = reset_noise<R>(T, h, n, e.episode);
    GlobalObsSink sink{T.st.cur_obs, A.obs, A.obs_f64, (size_t)n * od};
    make_obs<SYS, R>(T, h, state, n, 0, true, ns, sink);
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
    env_step<SYS, R>(T, h, n, e, action, nrow, s_z + threadIdx.x, SCG_LANE_BLOCK, state, rew, mse, info, cv, nullptr, ns);
    e.ep_ret += (float)rew; e.ep_len += 1.f; e.ep_cv += (float)info.cviol;
    const bool reset_now = info.done && c->auto_reset;
    {   // observation of the transition: the returned obs unless the env auto-resets, always the terminal obs
        GlobalObsSink term{o.d_terminal_obs, nullptr, o.d_terminal_obs_f64, (size_t)n * od};
        GlobalObsSink cur{reset_now ? nullptr : T.st.cur_obs, reset_now ? nullptr : o.d_obs, reset_now ? nullptr : o.d_obs_f64, (size_t)n * od};
        auto both = [&](int k, R v) { term(k, v); cur(k, v); };
        make_obs<SYS, R>(T, h, state, n, e.ctrl_step - 1, false, ns, both);
    }
    if (o.d_terminal_state_f64) for (int k = 0; k < NX; ++k) o.d_terminal_state_f64[(size_t)n * NX + k] = (double)state[k];
    if (reset_now) {   // dummy_vec_env.py:33-38
        env_reset<SYS, R>(T, n, e, state);
        e.ep_ret = 0.f; e.ep_len = 0.f; e.ep_cv = 0.f;
        NoiseSrc<R> rs = reset_noise<R>(T, h, n, e.episode);
        GlobalObsSink sink{T.st.cur_obs, o.d_obs, o.d_obs_f64, (size_t)n * od};
        make_obs<SYS, R>(T, h, state, n, 0, true, rs, sink);
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
};

// shared memory carve-up (floats), all offsets multiples of 4
template <int HID, int E> struct RolloutSmem {
    static constexpr int EP = E + 4;   // padded env stride of the observation tiles
    int blob, obs, tobs, h1, h2, out, total_floats;
    __host__ __device__ RolloutSmem(int obs_dim, int act_dim) {
        BlobLayout L = blob_layout(obs_dim, act_dim, HID);
        blob = 0;
        obs = blob + L.total;
        tobs = obs + obs_dim * EP;
        h1 = tobs + obs_dim * EP;
        h2 = h1 + 2 * HID * E;
        out = h2 + 2 * HID * E;
        total_floats = out + pad4i((act_dim + 1) * E);
    }
};

// One fused actor+critic forward for the CTA's E envs.
//   in  : s_in  [obs_dim][EP]  (float observations, env-minor)
//   out : s_out [A+1][E]       (A action means, then the value)
// Layers 1 and 2 are register-tiled FP32 GEMMs: each thread owns a (4 envs x TU units) tile, weights and
// activations stream from shared memory as 128-bit loads (weights broadcast across the env groups of a warp,
// activations broadcast across its unit groups).
template <int HID, int E, int NT, int A>
__device__ __noinline__ void mlp_forward(const float* __restrict__ s_blob, const BlobLayout L, const float* __restrict__ s_in, float* __restrict__ s_h1,
                                         float* __restrict__ s_h2, float* __restrict__ s_out, int obs_dim, int act) {
    constexpr int EP = E + 4;
    constexpr int EG = E / 4;              // env groups
    constexpr int UG = NT / EG;            // unit groups
    constexpr int TU = (2 * HID) / UG;     // units per thread
    static_assert(TU == 2 || TU == 4 || TU == 8, "unsupported tile");
    static_assert(HID % TU == 0 && (2 * HID) % UG == 0, "tile must not straddle nets");
    const int tid = threadIdx.x;
    const int eg = tid % EG, ug = tid / EG;
    const int e0 = eg * 4, u0 = ug * TU;
    const int net = u0 / HID, ul0 = u0 % HID;

    float acc[4][TU];
    // ---- layer 1: [E x obs] . [obs x 2H]
    {
        const float* b1 = s_blob + L.b1 + u0;
#pragma unroll
        for (int j = 0; j < TU; ++j) { float b = b1[j]; acc[0][j] = b; acc[1][j] = b; acc[2][j] = b; acc[3][j] = b; }
        const float* w = s_blob + L.w1t + u0;
        const float* a = s_in + e0;
#pragma unroll 4
        for (int k = 0; k < obs_dim; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(a + k * EP);
            float wv[TU];
            if (TU == 2) { float2 t = *reinterpret_cast<const float2*>(w + k * 2 * HID); wv[0] = t.x; wv[1] = t.y; }
            else {
#pragma unroll
                for (int j = 0; j < TU; j += 4) { float4 t = *reinterpret_cast<const float4*>(w + k * 2 * HID + j); wv[j] = t.x; wv[j + 1] = t.y; wv[j + 2] = t.z; wv[j + 3] = t.w; }
            }
#pragma unroll
            for (int j = 0; j < TU; ++j) {
                acc[0][j] = fmaf(av.x, wv[j], acc[0][j]); acc[1][j] = fmaf(av.y, wv[j], acc[1][j]);
                acc[2][j] = fmaf(av.z, wv[j], acc[2][j]); acc[3][j] = fmaf(av.w, wv[j], acc[3][j]);
            }
        }
#pragma unroll
        for (int j = 0; j < TU; ++j) {
            float4 h;
            h.x = hidden_act_rt(acc[0][j], act); h.y = hidden_act_rt(acc[1][j], act); h.z = hidden_act_rt(acc[2][j], act); h.w = hidden_act_rt(acc[3][j], act);
            *reinterpret_cast<float4*>(s_h1 + (u0 + j) * E + e0) = h;
        }
    }
    __syncthreads();
    // ---- layer 2: per net [E x H] . [H x H]
    {
        const float* b2 = s_blob + L.b2 + u0;
#pragma unroll
        for (int j = 0; j < TU; ++j) { float b = b2[j]; acc[0][j] = b; acc[1][j] = b; acc[2][j] = b; acc[3][j] = b; }
        const float* w = s_blob + L.w2t + net * HID * HID + ul0;
        const float* a = s_h1 + net * HID * E + e0;
#pragma unroll 8
        for (int k = 0; k < HID; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(a + k * E);
            float wv[TU];
            if (TU == 2) { float2 t = *reinterpret_cast<const float2*>(w + k * HID); wv[0] = t.x; wv[1] = t.y; }
            else {
#pragma unroll
                for (int j = 0; j < TU; j += 4) { float4 t = *reinterpret_cast<const float4*>(w + k * HID + j); wv[j] = t.x; wv[j + 1] = t.y; wv[j + 2] = t.z; wv[j + 3] = t.w; }
            }
#pragma unroll
            for (int j = 0; j < TU; ++j) {
                acc[0][j] = fmaf(av.x, wv[j], acc[0][j]); acc[1][j] = fmaf(av.y, wv[j], acc[1][j]);
                acc[2][j] = fmaf(av.z, wv[j], acc[2][j]); acc[3][j] = fmaf(av.w, wv[j], acc[3][j]);
            }
        }
#pragma unroll
        for (int j = 0; j < TU; ++j) {
            float4 h;
            h.x = hidden_act_rt(acc[0][j], act); h.y = hidden_act_rt(acc[1][j], act); h.z = hidden_act_rt(acc[2][j], act); h.w = hidden_act_rt(acc[3][j], act);
            *reinterpret_cast<float4*>(s_h2 + (u0 + j) * E + e0) = h;
        }
    }
    __syncthreads();
    // ---- layer 3: A action means from the actor features, 1 value from the critic features
    {
        const int g = tid / E, e = tid % E;
        if (g < A + 1) {
            const float* w = s_blob + L.w3 + g * HID;
            const float* h = s_h2 + (g < A ? 0 : HID) * E + e;
            float s0 = s_blob[L.b3 + g], s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
            for (int k = 0; k < HID; k += 4) {
                const float4 wv = *reinterpret_cast<const float4*>(w + k);
                s0 = fmaf(wv.x, h[(k + 0) * E], s0); s1 = fmaf(wv.y, h[(k + 1) * E], s1);
                s2 = fmaf(wv.z, h[(k + 2) * E], s2); s3 = fmaf(wv.w, h[(k + 3) * E], s3);
            }
            s_out[g * E + e] = (s0 + s1) + (s2 + s3);
        }
    }
    __syncthreads();
}

template <int SYS, typename R, int HID, int E, int NT>
__global__ void __launch_bounds__(NT) rollout_kernel(DevTables T, RolloutParams P) {
    constexpr int NX = SysDims<SYS>::NX, A = SysDims<SYS>::NU;
    constexpr int EP = E + 4;
    static_assert(NT / E >= A + 1, "layer 3 needs A+1 lane groups");
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ HotCfg<R> s_hot;
    __shared__ double s_z[(NX + A) * E];

    const scg_config* c = T.cfg;
    const int od = T.obs_dim, N = T.N, nc = T.n_con;
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * E;
    const RolloutSmem<HID, E> S(od, A);
    const BlobLayout L = blob_layout(od, A, HID);
    float* s_blob = smem + S.blob;
    float* s_obs = smem + S.obs;
    float* s_tobs = smem + S.tobs;
    float* s_h1 = smem + S.h1;
    float* s_h2 = smem + S.h2;
    float* s_out = smem + S.out;

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
    EnvRegs<SYS, R> e;
    if (live) env_load<SYS, R>(T, n, e);
    for (int i = tid; i < E * od; i += NT) {
        const int ee = i / od, k = i % od;
        s_obs[k * EP + ee] = (n0 + ee < N) ? T.st.cur_obs[(size_t)(n0 + ee) * od + k] : 0.f;
        s_tobs[k * EP + ee] = 0.f;
    }
    float ep_ret_sum = 0.f, ep_len_sum = 0.f, ep_cv_sum = 0.f; int ep_count = 0;
    mbar_wait(&s_bar, 0);
    __syncthreads();
    float sd[A], lsd[A];
#pragma unroll
    for (int k = 0; k < A; ++k) { lsd[k] = s_blob[L.logstd + k]; sd[k] = expf(lsd[k]); }
    const float log_sqrt_2pi = 0.91893853320467267f;
    const bool philox_eps = (P.eps == nullptr);
    RngKey pkey; pkey.k0 = (uint32_t)c->seed; pkey.k1 = (uint32_t)(c->seed >> 32); pkey.genv = (uint32_t)(c->env_index_offset + n);

    for (int t = 0; t < P.T; ++t) {
        const size_t tn0 = (size_t)t * N + n0;
        // standard-normal draws for this step (table mode): issue the loads before the MLP hides their latency
        float z[A];
#pragma unroll
        for (int k = 0; k < A; ++k) z[k] = 0.f;
        if (live && !P.deterministic) {
            if (!philox_eps) {
#pragma unroll
                for (int k = 0; k < A; ++k) z[k] = __ldg(P.eps + (tn0 + tid) * A + k);
            } else {
#pragma unroll
                for (int k = 0; k < A; ++k) z[k] = rng_normal<float>(pkey, SCG_STREAM_POLICY, e.total_steps, (uint32_t)k);
            }
        }
        // (a) PPOBuffer.obs[t] <- observation tile (coalesced), then the policy forward
        {
            float* dst = P.out.d_obs + tn0 * od;
            const int lim = (N - n0 < E ? N - n0 : E) * od;
            for (int i = tid; i < lim; i += NT) dst[i] = s_obs[(i % od) * EP + (i / od)];
        }
        mlp_forward<HID, E, NT, A>(s_blob, L, s_obs, s_h1, s_h2, s_out, od, P.activation);

        // (b) owner lanes: sample, step the env, emit the record
        int trunc = 0;
        if (live) {
            float act[A], logp = 0.f;
#pragma unroll
            for (int k = 0; k < A; ++k) {
                const float mu = s_out[k * E + tid];
                act[k] = fmaf(sd[k], z[k], mu);                       // Normal.sample (ppo_utils.py:228)
                const float d = act[k] - mu;
                logp += -(d * d) / (2.f * sd[k] * sd[k]) - lsd[k] - log_sqrt_2pi;   // distributions.py:12-21
            }
            const float val = s_out[A * E + tid];
            R action[A];
#pragma unroll
            for (int k = 0; k < A; ++k) action[k] = (R)act[k];
            const double* nrow = (!h->philox && T.noise_table && T.D > 0)
                                     ? T.noise_table + ((size_t)((P.noise_cursor + t) % T.Tn) * N + n) * T.D : nullptr;
            R state[NX], rew, mse; StepInfo info; NoiseSrc<R> ns;
            float* cvp = P.out.d_cvals ? P.out.d_cvals + (tn0 + tid) * nc : nullptr;
            env_step<SYS, R>(T, h, n, e, action, nrow, s_z + tid, E, state, rew, mse, info, nullptr, cvp, ns);
            e.ep_ret += (float)rew; e.ep_len += 1.f; e.ep_cv += (float)info.cviol;
            if (info.done) {
                if (info.truncated) {                                  // ppo.py:274-283
                    trunc = 1;
                    make_obs<SYS, R>(T, h, state, n, e.ctrl_step - 1, false, ns, SmemObsSink{s_tobs + tid, EP});
                }
                ep_count += 1; ep_ret_sum += e.ep_ret; ep_len_sum += e.ep_len; ep_cv_sum += e.ep_cv;
                e.ep_ret = 0.f; e.ep_len = 0.f; e.ep_cv = 0.f;
                env_reset<SYS, R>(T, n, e, state);                     // dummy_vec_env.py:33-38
                NoiseSrc<R> rs = reset_noise<R>(T, h, n, e.episode);
                make_obs<SYS, R>(T, h, state, n, 0, true, rs, SmemObsSink{s_obs + tid, EP});
            } else {
                make_obs<SYS, R>(T, h, state, n, e.ctrl_step - 1, false, ns, SmemObsSink{s_obs + tid, EP});
            }
            const size_t tn = tn0 + tid;
            if (A == 2) *reinterpret_cast<float2*>(P.out.d_act + tn * A) = make_float2(act[0], act[1 % A]);
            else if (A == 4) *reinterpret_cast<float4*>(P.out.d_act + tn * A) = make_float4(act[0], act[1 % A], act[2 % A], act[3 % A]);
            else P.out.d_act[tn * A] = act[0];
            P.out.d_rew[tn] = (float)rew;
            P.out.d_mask[tn] = info.done ? 0.f : 1.f;
            P.out.d_v[tn] = val;
            P.out.d_logp[tn] = logp;
            if (P.out.d_cviol) P.out.d_cviol[tn] = (float)info.cviol;
            if (P.out.d_mse) P.out.d_mse[tn] = (float)mse;
            if (!trunc) P.out.d_terminal_v[tn] = 0.f;
        }
        // (c) critic on the terminal observation of time-limit truncated envs (rare, CTA-uniform branch)
        if (__syncthreads_or(trunc)) {
            mlp_forward<HID, E, NT, A>(s_blob, L, s_tobs, s_h1, s_h2, s_out, od, P.activation);
            if (live && trunc) P.out.d_terminal_v[tn0 + tid] = s_out[A * E + tid];
            __syncthreads();
        }
    }
    // ---- bootstrap value of the last observation, state write-back, GAE
    mlp_forward<HID, E, NT, A>(s_blob, L, s_obs, s_h1, s_h2, s_out, od, P.activation);
    if (live) {
        env_store<SYS, R>(T, n, e);
        for (int k = 0; k < od; ++k) {
            const float v = s_obs[k * EP + tid];
            T.st.cur_obs[(size_t)n * od + k] = v;
            if (P.out.d_last_obs) P.out.d_last_obs[(size_t)n * od + k] = v;
        }
        const float lval = s_out[A * E + tid];
        if (P.out.d_last_val) P.out.d_last_val[n] = lval;
        if (P.out.d_ep_count) { P.out.d_ep_count[n] = ep_count; P.out.d_ep_return_sum[n] = ep_ret_sum; P.out.d_ep_length_sum[n] = ep_len_sum; P.out.d_ep_cviol_sum[n] = ep_cv_sum; }
        if (P.out.d_ret && P.out.d_adv) {
            // compute_returns_and_advantages (ppo_utils.py:374-400): float32 ret / td recursions, float64 GAE accumulator
            float ret = lval, vnext = lval; double adv = 0.0;
            for (int t = P.T - 1; t >= 0; --t) {
                const size_t tn = (size_t)t * N + n;
                const float r = P.out.d_rew[tn] + P.gamma * P.out.d_terminal_v[tn];
                P.out.d_rew[tn] = r;                                   // rews += gamma * terminal_v, in place (ppo_utils.py:389)
                const float m = P.out.d_mask[tn], v = P.out.d_v[tn];
                ret = r + P.gamma * m * ret;
                if (!P.use_gae) adv = (double)(ret - v);
                else { const float td = r + P.gamma * m * vnext - v; adv = adv * P.lam_d * P.gamma_d * (double)m + (double)td; }
                P.out.d_ret[tn] = ret; P.out.d_adv[tn] = (float)adv; vnext = v;
            }
        }
    }
}

// Philox known-answer hook
__global__ void philox_selftest_kernel(const uint32_t* ctr, const uint32_t* key, uint32_t* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t o[4];
    philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], key[2 * i], key[2 * i + 1], o);
    out[4 * i] = o[0]; out[4 * i + 1] = o[1]; out[4 * i + 2] = o[2]; out[4 * i + 3] = o[3];
}
