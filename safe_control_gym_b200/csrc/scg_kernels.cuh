// scg_kernels.cuh -- sm_100a kernels of the rollout engine.
//
//   reset_kernel   : lane per env, (re-)initialises masked envs            (Quadrotor.reset / CartPole.reset)
//   step_kernel    : lane per env, one control step with caller actions    (env.step / DummyVecEnv.step_wait)
//   rollout_kernel : one CTA owns E envs for the whole horizon.  Per step: fused actor+critic MLP as
//                    register-tiled FP32 GEMMs out of shared memory (all warps), then the env step on the owner
//                    lanes (Gaussian sampling, action map, physics sub-steps, termination, next observation,
//                    auto-reset) while helper warps evaluate reward / constraints / statistics of the previous
//                    step, prepare the next reset episodes and write the record; terminal-value pass on
//                    truncation; after the horizon the critic bootstrap + reverse GAE scan.  No host round trip
//                    inside the horizon; on several GPUs the record is stored into every peer's gathered buffer.
//                    (PPO.train_step collector, safe_control_gym/controllers/ppo/ppo.py:259-297)
#pragma once
#include <type_traits>
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
// CTA-scope release / acquire on a shared-memory flag (reset candidates: helper warp -> owner lanes).  __threadfence_block() is fence.sc.cta
// (MEMBAR.SC.CTA: it waits for the warp's outstanding GLOBAL record stores as well) -- on the owner lanes' critical path in every step in which
// an episode ends.  The acquire load compiles to a plain LDS (a warp's shared-memory accesses are already ordered), the release to
// MEMBAR.ALL.CTA + STS on the helper, off the critical path.
__device__ __forceinline__ int lds_acquire(const int* p) {
    int v; asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory"); return v;
}
__device__ __forceinline__ void sts_release(int* p, int v) {
    asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
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

struct PlainObsSink {       // column of a [k][stride] float array
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
    env_reset<SYS, R>(T, h, n, e, state);
    e.ep_ret = 0.f; e.ep_len = 0.f; e.ep_cv = 0.f;
    T.st.initial_reset[n] = 1;
    T.st.adv_set[n] = 0;                     // benchmark_env.py:338-339: self.adv_action = None
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
    // adversary action left by scg_set_adversary_control (benchmark_env.py:216-228); the dynamics mode consumes it (quadrotor.py:424-425)
    R advv[SCG_MAX_ADV]; const R* adv = nullptr;
    if (h->adv_mode != SCG_ADV_NONE && T.st.adv_set[n]) {
#pragma unroll
        for (int k = 0; k < SCG_MAX_ADV; ++k) advv[k] = (R)T.st.adv[(size_t)k * T.N + n];
        adv = advv;
        if (h->adv_mode == SCG_ADV_DYNAMICS) T.st.adv_set[n] = 0;
    }
    env_step<SYS, R, double>(T, h, T.x_goal, n, e, action, nrow, s_z + threadIdx.x, SCG_LANE_BLOCK, state, rew, mse, info, cv, nullptr, ns, adv);
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
        T.st.adv_set[n] = 0;
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

// BenchmarkEnv.set_adversary_control (benchmark_env.py:216-228): adv_action = clip(a, -1, 1) * scale + offset
static __global__ void adversary_set_kernel(DevTables T, const float* __restrict__ a, int dim) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= T.N) return;
    // float32 arithmetic like the reference: a float32 action array times / plus python floats stays float32 (two rounded operations)
    const float sc = (float)T.cfg->adversary_scale, of = (float)T.cfg->adversary_offset;
    for (int k = 0; k < SCG_MAX_ADV; ++k) {
        float v = 0.f;
        if (k < dim) { const float x = fminf(fmaxf(a[(size_t)n * dim + k], -1.0f), 1.0f); v = __fadd_rn(__fmul_rn(x, sc), of); }
        T.st.adv[(size_t)k * T.N + n] = v;
    }
    T.st.adv_set[n] = 1;
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
    int n_peers;                 // fused all-gather: peer copies of the PPOBuffer keys (see scg_rollout_args)
    long long peer_off[SCG_MAX_PEERS];
    int safety_C, safety_H;      // SafeExplorer safety layer: constraint models (0 = none), hidden width
    const float* safety;         // [C] x { W1[H][obs], b1[H], W2[A][H], b2[A] }
    float safety_slack[SCG_MAX_SAFETY];
    const double* norm;          // frozen obs / reward normaliser statistics (scg_rollout_args.d_norm), nullable
    const float* obs0;           // policy input of step 0, nullable
    long long mc_off;            // != 0: NVSwitch multicast mapping of the gathered buffer (multimem.st), replaces the unicast peer copies
    // RARL / RAP adversary (scg_rollout_args.d_adv_policy): packed blob with act_dim = adv_dim, its normals, which player is recorded
    const float* adv_policy; const float* eps_adv; int adv_dim, adv_record;
    long long* prof;     // SCG_PROF builds: per-phase cycle counters of CTA 0 (nullable)
};

// shared memory carve-up (floats), all offsets multiples of 4
template <int HID, int E> struct RolloutSmem {
    int blob, obs, tobs, cobs, h1, out, xg, blob2, out2, total_floats;
    // xg_floats: size of the shared X_GOAL copy in float units (L * nx * sizeof(R) / 4); nw: GEMM warps per network;
    // adv_dim > 0: room for the adversary's blob and output partial sums (adv_max = the instantiation's SysDims::NADV)
    __host__ __device__ RolloutSmem(int obs_dim, int act_dim, int xg_floats, int nw, int adv_dim = 0, int adv_max = 0) {
        BlobLayout L = blob_layout(obs_dim, act_dim, HID);
        blob = 0;
        obs = blob + L.total;
        h1 = obs + 2 * obs_dim * E;          // two observation tiles (ping-pong): tile t+1 is built while tile t is still the policy input
        tobs = h1 + 2 * HID * E;             // terminal observations of time-limit truncated envs
        cobs = tobs + obs_dim * E;           // reset observations of the two candidate episodes
        out = cobs + 2 * obs_dim * E;
        xg = out + pad4i(nw * (act_dim + 1) * E);   // per-warp partial sums of the output layer
        blob2 = xg + pad4i(xg_floats);
        out2 = blob2 + (adv_dim > 0 ? blob_layout(obs_dim, adv_dim, HID).total : 0);
        total_floats = out2 + (adv_dim > 0 ? pad4i(nw * (adv_max + 1) * E) : 0);
    }
};

// Forward of ONE network (net 0 = actor with n_out = A outputs, net 1 = critic with 1 output) for the CTA's E envs by a group
// of NTN threads (tl = index inside the group); both groups run in lockstep (__syncthreads).
//   s_in  : pair-swizzled observation tile (tile_idx)          s_h1n : this net's hidden tile
//   s_outp: [NTN/32 warps][A+1 outputs][E] per-warp partial sums of the output layer (see mlp_out)
// Layers 1 and 2 are register-tiled FP32 GEMMs (4 envs x TU units per thread) out of shared memory with packed FFMA2.
// Layer 3 is fused into the epilogue of layer 2: every thread multiplies its 4 x TU activations with the matching slice of the
// output rows, the unit groups of a warp are summed with shuffles and the per-warp partial sums go to shared memory -- the second
// hidden tile is never stored and the whole forward needs two barriers.
// NP1: the input-pair count (obs_dim / 2) whose first layer is fully unrolled in this instantiation (0 = none): one per system -- every unrolled
// copy lengthens the code all eight warps fetch each step, and only one of them is ever executed
// BAR: 0 = both barriers are __syncthreads (the two nets of a CTA run in lockstep); 1 = this net's NTN threads (the CTA's FIRST NTN threads) run on
// their own: the barrier between the layers is a named barrier of that group, and there is no closing barrier (the caller has one)
template <int HID, int E, int NTN, int AMAX, int NP1 = 0, int BAR = 0>
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
    // the same with a compile-time trip count (fully unrolled: the short first layer is all prologue / epilogue otherwise)
    auto gemm_fixed = [&](const float* __restrict__ a, int a_stride, const float* __restrict__ w, int w_stride, auto npair_c) {
        constexpr int NP = decltype(npair_c)::value;
        Ops o0, o1;
        load_ops(o0, a, w);
#pragma unroll
        for (int kp = 0; kp < NP; kp += 2) {
            if (kp + 1 < NP) load_ops(o1, a + (kp + 1) * a_stride, w + (kp + 1) * w_stride);
            fma_ops(o0);
            if (kp + 2 < NP) load_ops(o0, a + (kp + 2) * a_stride, w + (kp + 2) * w_stride);
            if (kp + 1 < NP) fma_ops(o1);
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
    {
        const float* a1 = s_in + eg * 4; const float* w1 = s_blob + L.w1t + (net * HID + ul0) * 2;
        const int np1 = obs_dim >> 1;
        // fully unrolled for the system's usual width: Quadrotor-2D tracking (obs 12: 6 pairs), Quadrotor-3D tracking (obs 24: 12), CartPole (obs 4: 2)
        if (NP1 > 0 && np1 == NP1) gemm_fixed(a1, 2 * EG * 4, w1, 2 * HID * 2, std::integral_constant<int, (NP1 > 0 ? NP1 : 2)>{});
        else gemm(a1, 2 * EG * 4, w1, 2 * HID * 2, np1);
    }
    SCG_TK(15);
    activate();
#pragma unroll
    for (int j = 0; j < TU; j += 2) {              // store pair-swizzled as the input of layer 2
        float* d = s_h1n + ((((ul0 + j) >> 1) * 2) * EG + eg) * 4;
        *reinterpret_cast<float4*>(d) = make_float4(h[0][j], h[0][j + 1], h[1][j], h[1][j + 1]);
        *reinterpret_cast<float4*>(d + EG * 4) = make_float4(h[2][j], h[2][j + 1], h[3][j], h[3][j + 1]);
    }
    SCG_TK(16);
    if constexpr (BAR == 0) __syncthreads();
    else asm volatile("bar.sync 1, %0;" ::"n"(NTN) : "memory");
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
    if constexpr (BAR == 0) __syncthreads();
    SCG_TK(22);
}

// The lower half of the CTA computes the actor, the upper half the critic, in lockstep.
//   s_outp : [NT/64 warps][A+1][E] partial sums; read the outputs with mlp_out()
template <int HID, int E, int NT, int A, int NP1 = 0>
__device__ __noinline__ void mlp_forward(const float* __restrict__ s_blob, const BlobLayout L, const float* __restrict__ s_in, float* __restrict__ s_h1,
                                         float* __restrict__ s_outp, int obs_dim, int act) {
    constexpr int NTN = NT / 2;
    constexpr int NETSZ = (HID / 2) * (2 * (E / 4) * 4);       // floats of one net's hidden tile
    const int net = threadIdx.x / NTN, tl = threadIdx.x % NTN;
    net_forward<HID, E, NTN, A, NP1>(tl, net, net ? 1 : A, net ? A : 0, s_blob, L, s_in, s_h1 + net * NETSZ, s_outp, obs_dim, act);
}
// The forward of the horizon loop, inlined: the call costs ~600 cycles per step (far jump, callee-saved registers through the stack, return) on
// the path every warp waits for.  Measured (gpurun_out/r2_ab19.log): -1.2 % on C3, -3.8 % on C2; the register-starved Quadrotor-3D collector is
// 0.7 % FASTER with the call (the inlined body competes with the integrator state for registers), so it keeps it.
template <int HID, int E, int NT, int A, int NP1 = 0>
__device__ __forceinline__ void mlp_forward_inl(const float* __restrict__ s_blob, const BlobLayout L, const float* __restrict__ s_in, float* __restrict__ s_h1,
                                                float* __restrict__ s_outp, int obs_dim, int act) {
    constexpr int NTN = NT / 2;
    constexpr int NETSZ = (HID / 2) * (2 * (E / 4) * 4);
    const int net = threadIdx.x / NTN, tl = threadIdx.x % NTN;
    net_forward<HID, E, NTN, A, NP1>(tl, net, net ? 1 : A, net ? A : 0, s_blob, L, s_in, s_h1 + net * NETSZ, s_outp, obs_dim, act);
}
// ONE net (0 = actor, 1 = critic) on the CTA's first NT/2 threads, on their own (net_forward BAR = 1): the decoupled schedule of rollout_kernel
template <int HID, int E, int NT, int A, int NP1 = 0>
__device__ __noinline__ void half_forward(int net, const float* __restrict__ s_blob, const BlobLayout L, const float* __restrict__ s_in, float* __restrict__ s_h1,
                                          float* __restrict__ s_outp, int obs_dim, int act) {
    constexpr int NTN = NT / 2;
    constexpr int NETSZ = (HID / 2) * (2 * (E / 4) * 4);
    net_forward<HID, E, NTN, A, NP1, 1>(threadIdx.x, net, net ? 1 : A, net ? A : 0, s_blob, L, s_in, s_h1 + net * NETSZ, s_outp, obs_dim, act);
}
// the same for a network pair whose action width is only known at run time (the adversary: a_dim <= AMAX)
template <int HID, int E, int NT, int AMAX>
__device__ __noinline__ void mlp_forward_n(const float* __restrict__ s_blob, const BlobLayout L, const float* __restrict__ s_in, float* __restrict__ s_h1,
                                           float* __restrict__ s_outp, int obs_dim, int act, int a_dim) {
    constexpr int NTN = NT / 2;
    constexpr int NETSZ = (HID / 2) * (2 * (E / 4) * 4);
    const int net = threadIdx.x / NTN, tl = threadIdx.x % NTN;
    net_forward<HID, E, NTN, AMAX>(tl, net, net ? 1 : a_dim, net ? a_dim : 0, s_blob, L, s_in, s_h1 + net * NETSZ, s_outp, obs_dim, act);
}
// output o (0..A-1 action means, A = value) of env e: bias + the per-warp partial sums, in a fixed order
template <int E, int NW, int A>
__device__ __forceinline__ float mlp_out(const float* __restrict__ s_outp, const float* __restrict__ s_blob, const BlobLayout& L, int o, int e) {
    float v = s_blob[L.b3 + o];
#pragma unroll
    for (int w = 0; w < NW; ++w) v += s_outp[(w * (A + 1) + o) * E + e];
    return v;
}

// frozen normaliser statistics in shared memory (MeanStdNormalizer / RewardStdNormalizer, normalization.py:88-159)
struct NormState { double mean[SCG_MAX_OBS], inv[SCG_MAX_OBS], clip_obs, rew_scale, clip_rew; int obs_on, rew_on; };
__device__ __forceinline__ void norm_load(NormState* s, const double* g, int od, int tid, int nthreads) {
    for (int i = tid; i < od; i += nthreads) { s->mean[i] = g[i]; s->inv[i] = g[od + i]; }
    if (tid == 0) { s->clip_obs = g[2 * od]; s->obs_on = g[2 * od + 1] != 0.0; s->rew_scale = g[2 * od + 2]; s->clip_rew = g[2 * od + 3]; s->rew_on = g[2 * od + 4] != 0.0; }
}
__device__ __forceinline__ float norm_obs_value(const NormState* s, int k, float x) {     // np.clip((x - mean) / sqrt(var + eps), -clip, clip)
    const double y = ((double)x - s->mean[k]) * s->inv[k];
    return (float)fmin(fmax(y, -s->clip_obs), s->clip_obs);
}
__device__ __forceinline__ float norm_rew_value(const NormState* s, float r) {            // np.clip(x / sqrt(var + eps), -clip, clip)
    const double y = (double)r * s->rew_scale;
    return (float)fmin(fmax(y, -s->clip_rew), s->clip_rew);
}

// ---------------------------------------------------------------------------------------------
// SafeExplorer safety layer (safe_explorer_utils.py:15-176).  All three pieces are out of line: they only run when a safety layer is given.
// ---------------------------------------------------------------------------------------------
// g_i(obs) of every constraint model for the CTA's E envs: thread (e = tid % E, group = tid / E) handles models group, group + NT/E, ...
// s_g[(i * A + a) * E + e].  obs comes from a pair-swizzled tile (tile_idx<E>).
template <int E, int A> static __device__ __noinline__ void safety_models(const float* __restrict__ blob, int C, int H, int od, const float* tile,
                                                                          float* s_g, int tid, int nthreads) {
    const int e = tid % E, per = H * od + H + A * H + A;
    for (int i = tid / E; i < C; i += nthreads / E) {
        const float* w1 = blob + (size_t)i * per; const float* b1 = w1 + H * od; const float* w2 = b1 + H; const float* b2 = w2 + A * H;
        float g[A];
#pragma unroll
        for (int a = 0; a < A; ++a) g[a] = __ldg(b2 + a);
        for (int j = 0; j < H; ++j) {
            float hj = __ldg(b1 + j);
            for (int k = 0; k < od; ++k)
                hj = fmaf(__ldg(w1 + j * od + k), tile[((((k >> 1) * 2 + ((e >> 1) & 1)) * (E / 4) + (e >> 2)) << 2) + ((e & 1) << 1) + (k & 1)], hj);
            hj = fmaxf(hj, 0.0f);                                  // MLP default activation: relu (neural_networks.py:25)
#pragma unroll
            for (int a = 0; a < A; ++a) g[a] = fmaf(__ldg(w2 + a * H + j), hj, g[a]);
        }
#pragma unroll
        for (int a = 0; a < A; ++a) s_g[(i * A + a) * E + e] = g[a];
    }
}
// Owner lane: c_i = np.round(a_i z - b_i, 8) of the CURRENT state for the first C (single-variable) constraint rows (float32 like
// torch.FloatTensor(c)), recorded as SafePPOBuffer 'c', then mu <- mu - lambda* g_i* (get_safe_action, safe_explorer_utils.py:120-176) in
// float32 like the reference's torch code.  `st` / `mu` are caller-local arrays that only exist inside the safety branch.
template <int NX, int E, int A> static __device__ __noinline__ void safety_apply(const double* st, const ConRow* rows, int C, const float* s_g, const float* slack,
                                                                                 int e, float* mu, float* c_rec) {
    float best = -1.0f; int bi = 0;
    for (int i = 0; i < C; ++i) {
        const ConRow row = rows[i];
        double acc = __dmul_rn(row.val, st[row.col < NX ? row.col : 0]);
        if (row.flags & 2) acc = fabs(acc);
        const float ci = (float)(rint(__dsub_rn(acc, row.b) * 1e8) / 1e8);
        if (c_rec) c_rec[i] = ci;
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int a = 0; a < A; ++a) { const float g = s_g[(i * A + a) * E + e]; num = fmaf(g, mu[a], num); den = fmaf(g, g, den); }
        num = (num + ci) + slack[i];
        const float m = fmaxf(num / (den + 1e-8f), 0.0f);
        if (m > best) { best = m; bi = i; }                       // topk(..., 1): the first maximum
    }
#pragma unroll
    for (int a = 0; a < A; ++a) mu[a] = mu[a] - best * s_g[(bi * A + a) * E + e];
}

// Out of line on purpose: the collectors' owner path is register-starved (Quadrotor-3D: 255 registers); inlined, the disabled normaliser
// code alone cost +8 % on C4.  A call is only made when normalisation / raw recording is requested (kernel-parameter test).
static __device__ __noinline__ float norm_rew_apply(const NormState* s, float r) { return s->rew_on ? norm_rew_value(s, r) : r; }
// observation of env `ol` in a pair-swizzled tile: raw copy to `raw` (nullable), normalised in place
template <int E> static __device__ __noinline__ void norm_obs_pass(float* tile, int ol, int od, float* raw, const NormState* s, bool have_norm) {
    const bool on = have_norm && s->obs_on;
    for (int k = 0; k < od; ++k) {
        const int i = ((((k >> 1) * 2 + ((ol >> 1) & 1)) * (E / 4) + (ol >> 2)) << 2) + ((ol & 1) << 1) + (k & 1);     // tile_idx<E>(k, ol)
        const float x = tile[i];
        if (raw) raw[k] = x;
        if (on) tile[i] = norm_obs_value(s, k, x);
    }
}

#ifdef SCG_PROF
#define PROF_T(v) long long v = clock64()
#else
#define PROF_T(v) do { } while (0)
#endif

// Record store that is repeated into the peers' gathered buffers (NVLink peer memory; plain posted writes, coalesced like the
// local ones): the all-gather of the rollout record happens inside the collector, there is no separate collective to wait for.
// With an NVSwitch multicast mapping of the gathered buffer (P.mc_off), ONE `multimem.st` lands in every rank's copy instead of
// n_peers unicast stores: the replication happens in the switch, not in this SM's LSU.
// (no "memory" clobber on purpose: a clobber inside one arm of pstore's branch is a compiler barrier for EVERY record store -- it undid the
//  "all loads, then all stores" batching of the record writers and of the GAE scan and cost the single-GPU collector 360 cycles per step on the
//  owner lanes, measured with the SCG_PROF ticks.  The multicast store has no consumer inside the kernel, so its order is free.)
#ifdef SCG_MC_CLOBBER
#define SCG_MC_CL : "memory"
#else
#define SCG_MC_CL
#endif
__device__ __forceinline__ void multimem_st(float* p, float v) { asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(p), "f"(v) SCG_MC_CL); }
__device__ __forceinline__ void multimem_st(float2* p, float2 v) {
    asm volatile("multimem.st.weak.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) SCG_MC_CL);
}
__device__ __forceinline__ void multimem_st(float4* p, float4 v) {
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) SCG_MC_CL);
}
template <typename V> __device__ __forceinline__ void pstore(const RolloutParams& P, V* ptr, V v) {
    *ptr = v;
    if (P.mc_off) multimem_st(reinterpret_cast<V*>(reinterpret_cast<char*>(ptr) + P.mc_off), v);
    else {
        // not unrolled on purpose: pstore is inlined at every record store, and the unrolled peer loops were 30 KB of the horizon loop's code
        // (instruction fetch is what the single-GPU collector is short of); the unicast path is the fallback of the multicast one anyway
#pragma unroll 1
        for (int g = 0; g < P.n_peers; ++g) *reinterpret_cast<V*>(reinterpret_cast<char*>(ptr) + P.peer_off[g]) = v;
    }
}

// One row of the episode table (global) -> shared memory with a TMA bulk copy; the calling lane waits for it.  The mbarrier word is
// re-armed for every copy and its phase parity lives next to it (bar[1]), so the collector keeps no state in registers.
static __device__ __noinline__ const double* stage_table_row(double* dst, const double* src, int bytes, uint64_t* bar) {
    const uint32_t parity = (uint32_t)(bar[1] & 1u);
    mbar_expect_tx(bar, (uint32_t)bytes);
    tma_bulk_g2s(dst, src, (uint32_t)bytes, bar);
    mbar_wait(bar, parity);
    bar[1] = parity ^ 1u;
    return dst;
}

// CTAs per SM the register allocation is capped for: the 16-thread-per-env variants (E=16 x 256 threads, E=8 x 128 threads) halve
// the per-step MLP latency of a CTA and rely on TWO (or more) co-resident CTAs per SM to fill the machine.
__host__ __device__ constexpr int rollout_min_blocks(int HID, int E, int NT) {
    return (NT / E >= 16 && HID <= 64) ? (NT == 256 ? 2 : 3) : 1;
}
// FAST schedule, episode end before the helper warp has the next episode's start ready (only the first steps of a launch): draw it on the
// owner lane.  Out of line and by value -- this block is ~10 KB of code that would otherwise sit in the middle of the owner warp's hot path.
template <int SYS, typename R, int E> static __device__ SCG_COLD EnvRegs<SYS, R> reset_on_owner_lane(const DevTables& T, const HotCfg<R>* h, const R* s_xg, int n,
                                                                                                 EnvRegs<SYS, R> e, float* s_on, int ol) {
    R state[SysDims<SYS>::NX];
    env_reset<SYS, R>(T, h, n, e, state);
    NoiseSrc<R> nn; nn.row = nullptr; nn.philox = false; nn.slot = 0; nn.stream = 0; nn.idx = 0; nn.key.k0 = nn.key.k1 = nn.key.genv = 0;
    make_obs<SYS, R, R, false>(T, h, s_xg, state, n, 0, true, nn, TileObsSink<E>{s_on, ol});
    return e;
}

// SAFE: the SafeExplorer variant (safety layer between actor and env); a separate instantiation so that the default collector carries none of its code
// XT:   the extended variant -- RARL / RAP adversary (second actor-critic on the same observation, its action through set_adversary_control
//       into the env step; rarl.py:349-418), the env's adversary input and the non-default quadrotor physics (pyb_gnd / pyb_drag / pyb_dw);
//       again its own instantiation: none of it costs the default collector anything
// FM:   schedule selection.  SCG_FM_RUNTIME: the FAST / SERIAL schedule (see the env step below) is chosen inside the kernel from the configuration;
//       SCG_FM_FAST / SCG_FM_SERIAL: the HOST has made that choice (every input of it is known at launch, rollout_fast_ok) and the instantiation
//       only contains the chosen schedule -- the horizon loop of the default collector loses the whole SERIAL branch (reward / constraints /
//       finalize / reset on the owner lanes), which shortens the straight-line code the single owner warp has to fetch every step.
#define SCG_FM_RUNTIME 0
#define SCG_FM_FAST 1
#define SCG_FM_SERIAL 2
template <int SYS, typename R, int HID, int E, int NT, bool SAFE = false, bool XT = false, int FM = SCG_FM_RUNTIME>
__global__ void __launch_bounds__(NT, rollout_min_blocks(HID, E, NT)) rollout_kernel(DevTables T, RolloutParams P) {
    constexpr int NX = SysDims<SYS>::NX, A = SysDims<SYS>::NU, A2M = SysDims<SYS>::NADV;
    constexpr int NP1 = SYS == SCG_SYS_CARTPOLE ? 2 : SYS == SCG_SYS_QUAD2D ? 6 : SYS == SCG_SYS_QUAD3D ? 12 : 0;   // first-layer width unrolled for this system
    constexpr int NZ = NX + 2 * A;             // z scratch rows: state, noisy action, clipped action
    // With >= 4 warps the env-step section is split: warp 0 = owners (sampling, action map, physics, done, finalize, reset),
    // warp 1 = constraints, warp 2 = reward + mse, warp 3 = next observation.  Smaller CTAs keep everything on the owner lanes.
    constexpr bool SPLIT = (NT >= 128) && (E <= 32);
    constexpr int NW = NT / 64;                // GEMM warps per network
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ HotCfg<R> s_hot;
    __shared__ double s_z[2 * NZ * E];         // [state; noisy action; clipped action] of the last two steps (FAST: double-buffered)
    __shared__ float s_eps[2 * A * E];
    __shared__ uint32_t s_tsteps0[E];
    __shared__ int s_cstep[2 * E];             // ctrl_step BEFORE the increment, of the last two steps
    __shared__ int s_flag[2 * E];              // bit 0 done, bit 1 truncated, of the last two steps
    // reset candidates: a helper warp keeps the next two episodes of every env ready (slot = episode & 1):
    // inertial properties, their reciprocals, the integrator state of the initial condition (+ its observation in s_cobs)
    constexpr int PS = SysDims<SYS>::PS;
    constexpr int NC = SCG_MAX_INERTIAL + SCG_NINV + PS;
    __shared__ double s_cand[2 * NC * E];
    __shared__ int s_cand_ep[2 * E];           // episode index held by each candidate slot (-1 = empty)
    __shared__ int s_cand_raw[2 * E];          // episode index whose random draws sit in the slot, not yet turned into an episode start
    __shared__ int s_epi[E];                   // next episode index of each env, published by its owner lane
    __shared__ float s_rec[2 * (A + 2) * E];   // FAST: act, v, logp of the last two steps, stored to the record by a helper warp
    __shared__ float s_stat[6 * E];            // FAST: episode statistics kept by the record helpers, handed back at the end
    __shared__ int s_epc[E];
    __shared__ NormState s_norm;
    __shared__ float s_g[SAFE ? SCG_MAX_SAFETY * A * E : 1];       // safety layer: g_i(obs) of the current step
    __shared__ float s_slack[SAFE ? SCG_MAX_SAFETY : 1];
    // table mode: per-env rows of the [K][N][P] episode table, staged by TMA bulk copies (one mbarrier per env and candidate slot)
    constexpr int PROW = SCG_MAX_INERTIAL + NX;
    __shared__ __align__(16) double s_row[SPLIT ? 2 * E * PROW : 2];
    __shared__ __align__(8) uint64_t s_rowbar[SPLIT ? 2 * (2 * E) : 2];     // (mbarrier, parity) pairs

    const int od = T.obs_dim, N = T.N, nc = T.n_con;
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * E;
    const bool have_adv = XT && P.adv_policy != nullptr;
    const int A2 = XT ? P.adv_dim : 0;
    const RolloutSmem<HID, E> S(od, A, T.L * NX * (int)(sizeof(R) / 4), NW, have_adv ? A2 : 0, A2M);
    const BlobLayout L = blob_layout(od, A, HID);
    const BlobLayout L2 = blob_layout(od, have_adv ? A2 : 1, HID);
    float* s_blob2 = smem + S.blob2;                  // XT: the adversary's packed actor-critic and its output partial sums
    float* s_out2 = smem + S.out2;
    float* s_blob = smem + S.blob;
    float* s_obs = smem + S.obs;
    float* s_tobs = smem + S.tobs;
    float* s_cobs = smem + S.cobs;                    // reset observations of the two candidate slots, [slot][k][E]
    float* s_h1 = smem + S.h1;
    float* s_out = smem + S.out;
    R* s_xg = reinterpret_cast<R*>(smem + S.xg);      // X_GOAL [L][nx] in the integrator's precision

    // ---- stage the packed policy into shared memory with one TMA bulk copy
    if (tid == 0) { mbar_init(&s_bar, 1); if (SPLIT) for (int i = 0; i < 2 * E; ++i) { mbar_init(&s_rowbar[2 * i], 1); s_rowbar[2 * i + 1] = 0; } fence_mbar_init(); }
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
    if (owner) { s_tsteps0[ol] = live ? e.total_steps : 0u; s_epi[ol] = live ? e.episode : 0; s_cand_ep[ol] = -1; s_cand_ep[E + ol] = -1; s_cand_raw[ol] = -1; s_cand_raw[E + ol] = -1; }
    const bool use_cand = SPLIT && h->n_off_draws == 0;
    for (int i = tid; i < T.L * NX; i += NT) s_xg[i] = (R)T.x_goal[i];
    const float* obs_src = P.obs0 ? P.obs0 : T.st.cur_obs;
    for (int i = tid; i < E * od; i += NT) {
        const int ee = i / od, k = i % od;
        s_obs[tile_idx<E>(k, ee)] = (n0 + ee < N) ? obs_src[(size_t)(n0 + ee) * od + k] : 0.f;
    }
    if (P.norm) norm_load(&s_norm, P.norm, od, tid, NT);
    if constexpr (SAFE) { if (tid < SCG_MAX_SAFETY) s_slack[tid] = P.safety_slack[tid]; }
    if constexpr (XT) { if (have_adv) for (int i = tid; i < L2.total; i += NT) s_blob2[i] = P.adv_policy[i]; }
    // (the normaliser switches are read from s_norm / the kernel parameters where they are needed: a flag kept in a register across the
    //  horizon loop costs the register-starved 3-D instantiations spills -- measured +10 % on C4)
    float ep_ret_sum = 0.f, ep_len_sum = 0.f, ep_cv_sum = 0.f; int ep_count = 0;
    mbar_wait(&s_bar, 0);
    __syncthreads();
    float sd[A], lsd[A], inv2var[A];
#pragma unroll
    for (int k = 0; k < A; ++k) { lsd[k] = s_blob[L.logstd + k]; sd[k] = expf(lsd[k]); inv2var[k] = 1.0f / (2.f * sd[k] * sd[k]); }
    const float log_sqrt_2pi = 0.91893853320467267f;
    const bool philox_eps = (P.eps == nullptr);
    const bool obs_noise = h->has_dist[SCG_MODE_OBSERVATION] != 0;
    // (with an adversary the owner lanes keep the whole step: which player's values go into the record is decided there)
    const bool fast = FM == SCG_FM_FAST ? true : FM == SCG_FM_SERIAL ? false : (use_cand && !obs_noise && !h->done_on_violation && !h->use_penalty && !have_adv);
#ifndef SCG_RELAY_ALWAYS
#define SCG_RELAY_ALWAYS 1      // measured (gpurun_out/r2_cold_ab.log): -0.8 % on C3, -3.2 % on the Quadrotor-3D collector, -0.9 % on C2
#endif
    const bool relay = fast && (P.n_peers > 0 || P.mc_off != 0 || SCG_RELAY_ALWAYS);   // the owner lanes leave act / v / logp to a helper warp when every store has peer copies
    // DECOUPLED CRITIC (CA).  Nothing inside a step waits for V(s_t): it only goes into the record.  So the step is cut differently: in the ACTOR
    // phase warps 0-3 evaluate the actor alone (the critical path: obs -> action) while the helper warps 4-6 work off step t-1 (reward / mse,
    // constraint rows, record stores, the draws of coming episodes) and the owner warp waits; in the ENV phase the owner warp integrates while warps
    // 0-3 evaluate the CRITIC on the same observation tile and warps 4-6 do what the idle warps did before (obs record, next normals, second half of
    // the episode preparation).  The value reaches the record through shared memory one phase later.  Same tiles and summation order: same bits.
    // MEASURED (gpurun_out/r2_ab33.log, profiles/ab_r02_critic_async.log; all 126 GPU parity tests green with it): SLOWER -- C3 1.173 -> 1.191 ms, CartPole
    // 0.941 -> 1.043 ms, Quadrotor-3D 4.64 -> 5.18 ms.  Per-warp timing shows why: one net on four warps (one per SM sub-partition) takes 5.0 k cycles
    // alone and 6.1 k next to the owner warp, against 6.7 k for BOTH nets in lockstep on eight warps -- the forward is bound by the length of each
    // warp's own instruction stream (LDS latency behind a one-deep operand ping-pong), not by FFMA2 throughput, so halving the work per phase does
    // not halve the phase.  Off by default; kept as a compile-time variant.
#ifndef SCG_CRITIC_ASYNC
#define SCG_CRITIC_ASYNC 0
#endif
    constexpr bool CA = SCG_CRITIC_ASYNC && SCG_RELAY_ALWAYS && FM == SCG_FM_FAST && !SAFE && !XT && SPLIT && NT == 256 && E == 32;
    __shared__ __align__(16) float s_vp[CA ? NW * (A + 1) * E : 4];     // the critic's per-warp partial sums (layout of s_out; read with mlp_out)
    const int OBS_TILE = od * E;
    // ---- helper-warp work of the FAST schedule (lane ee = env ee of the CTA)
    const int ee = lane;
    const bool hl = SPLIT && role >= 1 && role <= 3 && ee < E && n0 + ee < N;
    float h_a = 0.f, h_b = 0.f, h_sa = 0.f, h_sb = 0.f; int h_cnt = 0;   // role 2: ep return / length (+ sums), role 1: ep violations (+ sum)
    if (hl && fast) {
        if (role == 2) { h_a = T.st.ep_acc[n0 + ee]; h_b = T.st.ep_acc[N + n0 + ee]; }
        if (role == 1) h_a = T.st.ep_acc[2 * N + n0 + ee];
    }
    auto store_policy_record = [&](size_t tn, const float* act, float val, float logp, int done, int truncated) {
        if (A == 2) pstore(P, reinterpret_cast<float2*>(P.out.d_act + tn * A), make_float2(act[0], act[1 % A]));
        else if (A == 4) pstore(P, reinterpret_cast<float4*>(P.out.d_act + tn * A), make_float4(act[0], act[1 % A], act[2 % A], act[3 % A]));
        else pstore(P, P.out.d_act + tn * A, act[0]);
        pstore(P, P.out.d_v + tn, val);
        pstore(P, P.out.d_logp + tn, logp);
        pstore(P, P.out.d_mask + tn, done ? 0.f : 1.f);
        if (!truncated) pstore(P, P.out.d_terminal_v + tn, 0.f);       // truncated: the terminal-value pass stores critic(terminal obs)
    };
    // record entries of step tp that are not on the owners' critical path (reward, mse, constraint rows, episode statistics,
    // and the stores of act / v / logp / mask / terminal_v the owner lane left in shared memory)
    auto process_record = [&](int tp) {
        const int bsel = tp & 1;
        const double* zb = s_z + bsel * NZ * E + ee;
        const size_t tnp = (size_t)tp * N + n0 + ee;
        const int fl = s_flag[bsel * E + ee], dn = fl & 1;
        if (role == 1) {                                                   // constraints (benchmark_env.py:472-485) + policy part of the record
            if (relay) {
                const float* rb = s_rec + bsel * (A + 2) * E + ee;
                float act[A];
#pragma unroll
                for (int k = 0; k < A; ++k) act[k] = rb[k * E];
                const float vrec = CA ? mlp_out<E, NW, A>(s_vp, s_blob, L, A, ee) : rb[A * E];
                store_policy_record(tnp, act, vrec, rb[(A + 1) * E], dn, fl >> 1);
            }
            int cv = 0;
            if (nc > 0) cv = env_constraints(T, h->rows, nc, h->con_fast, zb, E, nullptr, P.out.d_cvals ? P.out.d_cvals + tnp * nc : nullptr);
            if (P.out.d_cviol) P.out.d_cviol[tnp] = (float)cv;
            h_a += (float)cv;
            if (dn) { h_sa += h_a; h_a = 0.f; }
        } else {                                                           // reward + info['mse']
            R st[NX], ny[A], cl[A], rw, ms;
#pragma unroll
            for (int k = 0; k < NX; ++k) st[k] = (R)zb[k * E];
#pragma unroll
            for (int k = 0; k < A; ++k) { ny[k] = (R)zb[(NX + k) * E]; cl[k] = (R)zb[(NX + A + k) * E]; }
            env_reward_mse<SYS, R, R>(h, s_xg, st, ny, cl, s_cstep[bsel * E + ee], rw, ms);
            if (P.out.d_rew_raw) P.out.d_rew_raw[tnp] = (float)rw;
            pstore(P, P.out.d_rew + tnp, P.norm ? norm_rew_apply(&s_norm, (float)rw) : (float)rw);
            if (P.out.d_mse) P.out.d_mse[tnp] = (float)ms;
            h_a += (float)rw; h_b += 1.f;
            if (dn) { h_cnt += 1; h_sa += h_a; h_sb += h_b; h_a = 0.f; h_b = 0.f; }
        }
    };
    // keep the next two episodes of env ee ready (a pure function of (env, episode)); runs while the owners sample and integrate
    // table mode: the env's row of episode `ep` of the [K][N][P] table is staged with a TMA bulk copy (cp.async.bulk + mbarrier) by the
    // out-of-line stage_table_row(): one mbarrier per (candidate slot, env), its parity kept in shared memory
    // Preparing an episode start costs ~1100 instructions on one warp (~5-6 k cycles: three Philox blocks, the additive draws, reciprocals,
    // sin / cos of the initial pitch, the reset observation) -- MORE than the owner lanes' whole env step, and some env of the CTA needs one in
    // 62 % of the steps (C3): the end-of-step barrier was waiting for this helper, not for the owners.  So the job is split in two halves
    // that run on two warps concurrently: `draw_candidates` (helper role 3: the random draws of the episode into the slot, then a release on
    // s_cand_raw) and `finish_candidates` (one of the otherwise idle warps: everything derived from the draws, then the release on s_cand_ep the
    // owner lanes look at).  A candidate now takes two steps instead of one; it is not needed before the env's NEXT episode ends.
    // Philox mode: the draws of ONE env's episode are made by the whole warp, lane j draw slot j (each slot is an independent counter-based
    // draw), instead of by the env's lane looping over the slots -- ~10 x shorter per episode, and only a few envs need one in a step.
    auto draw_candidates_coop = [&]() {              // called by ALL 32 lanes of the helper warp
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
                __syncwarp();                                    // every lane's row is written (and ordered) before the flag
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
                double* cd = s_cand + slot * NC * E + ee;
                const double* staged = nullptr;
#ifndef SCG_NO_TMA_ROWS
                if (SPLIT && !h->philox && T.ep_table != nullptr && T.P == PROW && (T.P & 1) == 0)
                    staged = stage_table_row(s_row + (slot * E + ee) * PROW, T.ep_table + ((size_t)(ep % T.K) * T.N + n0 + ee) * T.P, PROW * 8, &s_rowbar[2 * (slot * E + ee)]);
#endif
                reset_draw<SYS, R>(T, h, n0 + ee, ep, cd, E, SCG_MAX_INERTIAL + SCG_NINV, staged);   // inertial at rows 0..3, initial state after the reciprocals
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
                make_obs<SYS, R, R>(T, h, s_xg, st0, n0 + ee, 0, true, nn, PlainObsSink{s_cobs + slot * OBS_TILE + ee, E});
                sts_release(&s_cand_ep[slot * E + ee], ep);
            }
        }
    };
#ifdef SCG_PROF
    const long long cta_t0 = clock64();
    if (tid == NT - 32) scg_tk_last = cta_t0;
    long long pacc[8]; for (int i = 0; i < 8; ++i) pacc[i] = 0;
#define PROF_ON (P.prof != nullptr && tid == NT - 32)
#endif

    // With idle warps next to the owner / helper warps (8-warp CTAs) the two jobs that only feed the record or the NEXT step --
    // PPOBuffer.obs[t] <- observation tile, and the Philox normals of step t+1 -- run on those warps during the env phase.
    constexpr int NIDLE = SPLIT ? NT - 128 : 0;
    constexpr bool DEFER = NIDLE >= 32 && NIDLE >= E;
    const bool gen_eps = !P.deterministic && philox_eps;
    auto make_eps = [&](int tt, int el) {                // threads el < E (el = tid: the first warp(s)): standard normals of step tt into s_eps[tt & 1]
        if (el < E && n0 + el < N) {
            const uint32_t idx = s_tsteps0[el] + (uint32_t)tt;           // == that env's total_steps at step tt
            float* dst = s_eps + (tt & 1) * A * E + el;
#pragma unroll
            for (int k = 0; k < A; k += 2) {                              // one Philox block serves two action dims
                uint32_t x[4]; philox4x32_10((uint32_t)(k >> 1), idx, h->genv0 + (uint32_t)(n0 + el), SCG_STREAM_POLICY, h->key0, h->key1, x);
                dst[k * E] = normal_from_words<float>(x[0], x[1]);
                if (k + 1 < A) dst[(k + 1) * E] = normal_from_words<float>(x[2], x[3]);
            }
        }
    };
    auto write_obs_record = [&](int tt, const float* tile, int t0, int nthr) {   // threads t0 = 0..nthr-1 copy the [E][od] tile to obs[tt]
        float* dst = P.out.d_obs + ((size_t)tt * N + n0) * od;
        const int lim = (N - n0 < E ? N - n0 : E) * od;
        int k = t0 % od, en = t0 / od;
        const int dk = nthr % od, de = nthr / od;
        for (int i = t0; i < lim; i += nthr) {
            pstore(P, dst + i, tile[tile_idx<E>(k, en)]);
            k += dk; en += de;
            if (k >= od) { k -= od; en += 1; }
        }
    };
    if (DEFER && gen_eps) make_eps(0, tid);
    __syncthreads();

    int cur = 0;
    for (int t = 0; t < P.T; ++t) {
        const size_t tn0 = (size_t)t * N + n0;
        float* s_oc = s_obs + cur * OBS_TILE;            // observation of step t (policy input)
        float* s_on = s_obs + (cur ^ 1) * OBS_TILE;      // observation of step t+1 (being built)
        PROF_T(k0);
        // standard-normal draws of this step.  Table mode: the owner lanes issue the loads now (latency hidden by the MLP).
        // Philox mode: s_eps[t & 1] -- generated by an idle warp during the env phase of step t-1 (DEFER), by the first lane
        // group right here otherwise.
        float z[A];
#pragma unroll
        for (int k = 0; k < A; ++k) z[k] = 0.f;
        if (!P.deterministic && !philox_eps && live) {
#pragma unroll
            for (int k = 0; k < A; ++k) z[k] = __ldg(P.eps + (tn0 + ol) * A + k);
        }
        if (!DEFER) {
            if (gen_eps) make_eps(t, tid);
            write_obs_record(t, s_oc, tid, NT);      // (a) PPOBuffer.obs[t] <- observation tile (coalesced)
        }
        PROF_T(k1);
        SCG_TK(13);
        if constexpr (CA) {
            // ACTOR phase: warps 0-3 the actor, warps 4-6 the helpers' work on step t-1, then everybody meets
            if (warp < 4) {
                if constexpr (SYS != SCG_SYS_QUAD3D) net_forward<HID, E, NT / 2, A, NP1, 1>(tid, 0, A, 0, s_blob, L, s_oc, s_h1, s_out, od, P.activation);
                else half_forward<HID, E, NT, A, NP1>(0, s_blob, L, s_oc, s_h1, s_out, od, P.activation);
            } else if (role == 3) {
                if (h->philox) draw_candidates_coop();              // (warp-uniform: the cooperative form needs all 32 lanes)
                else if (hl) draw_candidates();
            } else if (role != 0 && hl && t > 0) process_record(t - 1);
            __syncthreads();
        } else if constexpr (SYS != SCG_SYS_QUAD3D && FM != SCG_FM_RUNTIME) mlp_forward_inl<HID, E, NT, A, NP1>(s_blob, L, s_oc, s_h1, s_out, od, P.activation);
        else mlp_forward<HID, E, NT, A, NP1>(s_blob, L, s_oc, s_h1, s_out, od, P.activation);
        if constexpr (SAFE) { safety_models<E, A>(P.safety, P.safety_C, P.safety_H, od, s_oc, s_g, tid, NT); __syncthreads(); }
        if constexpr (XT) { if (have_adv) mlp_forward_n<HID, E, NT, A2M>(s_blob2, L2, s_oc, s_h1, s_out2, od, P.activation, A2); }
        PROF_T(k2);

        // (b) env step.
        // FAST schedule (helper warps available, pre-drawn reset candidates, no observation noise, no done_on_violation / constraint
        // penalty): the owner lanes run ONLY the chain the next policy forward depends on -- sample, action map, physics, termination
        // test, next observation, register reload on a reset -- and publish [state; actions], the step counter and the done bits in
        // a double-buffered scratch.  Reward, info['mse'], the constraint rows and the episode statistics of step t are evaluated by
        // two helper warps one step later, while the owners integrate step t+1: nobody waits for anybody inside the env phase.
        // Otherwise (SERIAL): the owner lanes do everything themselves.
        int trunc = 0;
        if (role == 0) {
            SCG_TK(0);
            float act[A], logp = 0.f, val = 0.f;
            if (live) {
                float mu[A];
#pragma unroll
                for (int k = 0; k < A; ++k) mu[k] = mlp_out<E, NW, A>(s_out, s_blob, L, k, ol);
                if constexpr (SAFE) {                                        // SafeExplorer: project the mean action (safe_ppo_utils.py:88-104)
                    R stc[NX]; double std_[NX]; float m2[A];
                    Physics<SYS, R>::state(e, stc);
#pragma unroll
                    for (int k = 0; k < NX; ++k) std_[k] = (double)stc[k];
#pragma unroll
                    for (int k = 0; k < A; ++k) m2[k] = mu[k];
                    safety_apply<NX, E, A>(std_, h->rows, P.safety_C, s_g, s_slack, ol, m2, P.out.d_c ? P.out.d_c + (tn0 + ol) * P.safety_C : nullptr);
#pragma unroll
                    for (int k = 0; k < A; ++k) mu[k] = m2[k];
                }
#pragma unroll
                for (int k = 0; k < A; ++k) {
                    const float zk = gen_eps ? s_eps[(t & 1) * A * E + k * E + ol] : z[k];
                    act[k] = fmaf(sd[k], zk, mu[k]);                      // Normal.sample (ppo_utils.py:228)
                    const float d = act[k] - mu[k];
                    logp += -(d * d) * inv2var[k] - lsd[k] - log_sqrt_2pi;   // distributions.py:12-21
                }
                if constexpr (!CA) val = mlp_out<E, NW, A>(s_out, s_blob, L, A, ol);      // (CA: the critic is still running; the record helper reads it)
                R action[A], state[NX];
#pragma unroll
                for (int k = 0; k < A; ++k) action[k] = (R)act[k];
                const double* nrow = (!h->philox && T.noise_table && T.D > 0)
                                         ? T.noise_table + ((size_t)((P.noise_cursor + t) % T.Tn) * N + n) * T.D : nullptr;
                StepCtx<SYS, R> cx; NoiseSrc<R> ns;
                // XT: the adversary's action for the same observation (rarl.py:362-367), sampled like the protagonist's, then
                // set_adversary_control: clip to [-1, 1], scale, offset in float32 (benchmark_env.py:223-224)
                float act2[A2M], logp2 = 0.f, val2 = 0.f; R advv[A2M]; const R* advp = nullptr;
                if constexpr (XT) {
                    if (have_adv) {
                        const float sc = (float)T.cfg->adversary_scale, of = (float)T.cfg->adversary_offset;
#pragma unroll
                        for (int k = 0; k < A2M; ++k) {
                            act2[k] = 0.f; advv[k] = (R)0;
                            if (k < A2) {
                                const float mu2 = mlp_out<E, NW, A2M>(s_out2, s_blob2, L2, k, ol);
                                const float lsd2 = s_blob2[L2.logstd + k], sd2 = expf(lsd2);
                                float z2 = 0.f;
                                if (!P.deterministic) {
                                    if (P.eps_adv) z2 = __ldg(P.eps_adv + (tn0 + ol) * A2 + k);
                                    else { RngKey kk; kk.k0 = h->key0; kk.k1 = h->key1; kk.genv = h->genv0 + (uint32_t)n; z2 = rng_normal<float>(kk, SCG_STREAM_ADVERSARY, e.total_steps, (uint32_t)k); }
                                }
                                act2[k] = fmaf(sd2, z2, mu2);
                                const float d2 = act2[k] - mu2;
                                logp2 += -(d2 * d2) * (1.0f / (2.f * sd2 * sd2)) - lsd2 - log_sqrt_2pi;
                                advv[k] = (R)__fadd_rn(__fmul_rn(fminf(fmaxf(act2[k], -1.0f), 1.0f), sc), of);
                            }
                        }
                        val2 = mlp_out<E, NW, A2M>(s_out2, s_blob2, L2, A2, ol);
                        advp = advv;
                    }
                }
                env_pre<SYS, R, XT>(T, h, n, e, action, nrow, cx, ns, advp);
                SCG_TK(1);
                env_physics<SYS, R, XT>(h, e, cx, state);
                SCG_TK(6);
                // z = [state; noisy action; clipped action] and the step counter of this step (constraints.py:166-171)
                double* zb = s_z + (fast ? (t & 1) * NZ * E : 0) + ol;
#pragma unroll
                for (int k = 0; k < NX; ++k) zb[k * E] = (double)state[k];
#pragma unroll
                for (int k = 0; k < A; ++k) { zb[(NX + k) * E] = (double)cx.noisy[k]; zb[(NX + A + k) * E] = (double)cx.clipped[k]; }
                const int step_before = e.ctrl_step;
                const size_t tn = tn0 + ol;
                int goal_reached = 0, oob = 0, done = env_done<SYS, R, R>(h, s_xg, state, goal_reached, oob), truncated = 0;
                SCG_TK(7);
                if (fast) {
                    const int fl = (step_before + 1 >= h->ctrl_steps) ? (done ? 1 : 3) : done;   // bit 0 done, bit 1 truncated (benchmark_env.py:497-502)
                    s_cstep[(t & 1) * E + ol] = step_before;
                    s_flag[(t & 1) * E + ol] = fl;
                    done = fl & 1; truncated = fl >> 1;
                    e.total_steps += 1;
                    // observation of the transition: the next policy input unless the episode ended; the terminal observation if truncated
                    if (fl != 1) make_obs<SYS, R, R, false>(T, h, s_xg, state, n, step_before, false, ns, TileObsSink<E>{truncated ? s_tobs : s_on, ol});   // FAST: no observation noise
                    SCG_TK(8);
                    if (done) {                                            // dummy_vec_env.py:33-38
                        const int slot = e.episode & 1;
                        if (lds_acquire(&s_cand_ep[slot * E + ol]) == e.episode) {     // acquire: no candidate load may move above the flag check
                            // the helper has the whole episode start ready: inertial properties, reciprocals, integrator state, observation
                            const double* cd = s_cand + slot * NC * E + ol;
#pragma unroll
                            for (int j = 0; j < SCG_MAX_INERTIAL; ++j) e.inertial[j] = (R)cd[j * E];
#pragma unroll
                            for (int j = 0; j < SCG_NINV; ++j) e.inv[j] = (R)cd[(SCG_MAX_INERTIAL + j) * E];
#pragma unroll
                            for (int j = 0; j < PS; ++j) e.ps[j] = (R)cd[(SCG_MAX_INERTIAL + SCG_NINV + j) * E];
                            const float* co = s_cobs + slot * OBS_TILE + ol;
                            for (int k = 0; k < od; k += 4) {               // loads first, then stores (both shared memory: no serial chain)
                                float cv[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) cv[j] = (k + j < od) ? co[(k + j) * E] : 0.f;
#pragma unroll
                                for (int j = 0; j < 4; ++j) if (k + j < od) s_on[tile_idx<E>(k + j, ol)] = cv[j];
                            }
                            e.ctrl_step = 0; e.episode += 1;
                        } else {                                           // not drawn yet (first steps of a launch): draw on this lane
                            e = reset_on_owner_lane<SYS, R, E>(T, h, s_xg, n, e, s_on, ol);
                        }
                        *(volatile int*)&s_epi[ol] = e.episode;
                    } else e.ctrl_step += 1;
                    SCG_TK(9);
                } else {
                    R rew, mse; int cviol = 0;
                    env_reward_mse<SYS, R, R>(h, s_xg, state, cx.noisy, cx.clipped, step_before, rew, mse);
                    if (nc > 0) cviol = env_constraints(T, h->rows, nc, h->con_fast, zb, E, nullptr, P.out.d_cvals ? P.out.d_cvals + tn * nc : nullptr);
                    e.ctrl_step += 1; e.total_steps += 1;
                    env_finalize<R>(h, e.ctrl_step, cviol, rew, done, truncated);
                    e.ep_ret += (float)rew; e.ep_len += 1.f; e.ep_cv += (float)cviol;
                    if (done) {
                        if (truncated) make_obs<SYS, R, R>(T, h, s_xg, state, n, step_before, false, ns, TileObsSink<E>{s_tobs, ol});
                        ep_count += 1; ep_ret_sum += e.ep_ret; ep_len_sum += e.ep_len; ep_cv_sum += e.ep_cv;
                        e.ep_ret = 0.f; e.ep_len = 0.f; e.ep_cv = 0.f;
                        env_reset<SYS, R>(T, h, n, e, state);
                        ns = reset_noise<R>(T, h, n, e.episode);
                        make_obs<SYS, R, R>(T, h, s_xg, state, n, 0, true, ns, TileObsSink<E>{s_on, ol});
                    } else {
                        make_obs<SYS, R, R>(T, h, s_xg, state, n, step_before, false, ns, TileObsSink<E>{s_on, ol});
                    }
                    if (P.out.d_rew_raw) P.out.d_rew_raw[tn] = (float)rew;
                    float rew_rec = P.norm ? norm_rew_apply(&s_norm, (float)rew) : (float)rew;
                    if constexpr (XT) { if (have_adv && P.adv_record) rew_rec = -rew_rec; }          // rarl.py:394: the adversary's reward
                    pstore(P, P.out.d_rew + tn, rew_rec);
                    if (P.out.d_cviol) P.out.d_cviol[tn] = (float)cviol;
                    if (P.out.d_mse) P.out.d_mse[tn] = (float)mse;
                }
                if (P.norm || P.out.d_obs_raw)           // the observation env.step returned: raw copy for the statistics update, normalised in place
                    norm_obs_pass<E>(s_on, ol, od, P.out.d_obs_raw ? P.out.d_obs_raw + tn * od : nullptr, &s_norm, P.norm != nullptr);
                trunc = truncated;
                if (relay) {
                    // multi-GPU: the record entries go through shared memory, a helper warp stores them and their peer copies one step later
                    float* rb = s_rec + (t & 1) * (A + 2) * E + ol;
#pragma unroll
                    for (int k = 0; k < A; ++k) rb[k * E] = act[k];
                    rb[A * E] = val; rb[(A + 1) * E] = logp;
                } else {
                    bool stored = false;
                    if constexpr (XT) {
                        if (have_adv && P.adv_record) {                    // rarl.py:391-399: act_adv, v_adv, logp_adv
#pragma unroll
                            for (int k = 0; k < A2M; ++k) if (k < A2) pstore(P, P.out.d_act + tn * A2 + k, act2[k]);
                            pstore(P, P.out.d_v + tn, val2); pstore(P, P.out.d_logp + tn, logp2);
                            pstore(P, P.out.d_mask + tn, done ? 0.f : 1.f);
                            if (!truncated) pstore(P, P.out.d_terminal_v + tn, 0.f);
                            stored = true;
                        }
                    }
                    if (!stored) store_policy_record(tn, act, val, logp, done, truncated);
                }
            }
            SCG_TK(12);
#ifdef SCG_PROF
            if (PROF_ON) { long long k5 = clock64(); pacc[0] += k1 - k0; pacc[1] += k2 - k1; pacc[2] += k5 - k2; }
#endif
        } else if constexpr (CA) {
            if (warp < 4) {                                               // ENV phase: the critic on the observation the actor just used
                half_forward<HID, E, NT, A, NP1>(1, s_blob, L, s_oc, s_h1, s_vp, od, P.activation);
            } else {                                                      // warps 4-6: record the policy input, draw the next normals, finish episode starts
                write_obs_record(t, s_oc, tid - NT / 2, NT / 2 - 32);
                if (role == 3 && hl) finish_candidates();
                if (role == 2 && gen_eps && t + 1 < P.T) make_eps(t + 1, lane);
            }
        } else if (SPLIT && role <= 3) {
            if (fast && role == 3 && E == 32 && h->philox) {   // (warp-uniform condition: the cooperative form needs all 32 lanes)
                draw_candidates_coop();
                if (!DEFER && hl) finish_candidates();
            } else if (fast && hl) {
                if (role == 3) { draw_candidates(); if (!DEFER) finish_candidates(); }     // (no idle warp in the small CTA shapes: both halves here)
                else if (t > 0) process_record(t - 1);
            }
        } else if (DEFER) {                                               // idle warps: record the policy input, draw the next normals
            write_obs_record(t, s_oc, tid, NIDLE);
            if (gen_eps && t + 1 < P.T) make_eps(t + 1, tid);
            if (fast && warp == 1 && ee < E && n0 + ee < N) finish_candidates();   // second half of the episode-start preparation (see above)
        }
        SCG_TK(23);
        PROF_T(k6);
#ifdef SCG_PROF
        // time every warp spends in its env-phase job (lane 0 of each warp of CTA 0): who does the end-of-step barrier wait for?
        if (P.prof != nullptr && blockIdx.x == 0 && lane == 0) atomicAdd((unsigned long long*)&P.prof[8 + warp], (unsigned long long)(k6 - k2));
#endif
        // (c) critic on the terminal observation of time-limit truncated envs (rare, CTA-uniform branch)
        const int any_trunc = __syncthreads_or(trunc);
#ifdef SCG_PROF
        if (PROF_ON) pacc[5] += clock64() - k6;
#endif
        if (any_trunc) {
            bool by_adv = false;
            if constexpr (XT) by_adv = have_adv && P.adv_record;               // rarl.py:385-386: the recorded player's critic
            if (by_adv) {
                mlp_forward_n<HID, E, NT, A2M>(s_blob2, L2, s_tobs, s_h1, s_out2, od, P.activation, A2);
                if (live && trunc) pstore(P, P.out.d_terminal_v + tn0 + ol, mlp_out<E, NW, A2M>(s_out2, s_blob2, L2, A2, ol));
            } else {
                mlp_forward<HID, E, NT, A, NP1>(s_blob, L, s_tobs, s_h1, s_out, od, P.activation);
                if (live && trunc) pstore(P, P.out.d_terminal_v + tn0 + ol, mlp_out<E, NW, A>(s_out, s_blob, L, A, ol));
            }
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
    // ---- FAST: the record helpers still owe the last step; they hand the episode statistics back to the owner lanes
    if (fast && hl && role != 3) {
        process_record(P.T - 1);
        if (role == 2) { s_stat[ee] = h_a; s_stat[E + ee] = h_b; s_stat[3 * E + ee] = h_sa; s_stat[4 * E + ee] = h_sb; s_epc[ee] = h_cnt; }
        else { s_stat[2 * E + ee] = h_a; s_stat[5 * E + ee] = h_sa; }
    }
    // ---- bootstrap value of the last observation (its barriers also order the helpers' record stores before the GAE scan),
    //      state write-back, GAE
    bool boot_adv = false;
    if constexpr (XT) boot_adv = have_adv && P.adv_record;                     // rarl.py:407-408
    if (boot_adv) mlp_forward_n<HID, E, NT, A2M>(s_blob2, L2, s_ol, s_h1, s_out2, od, P.activation, A2);
    else mlp_forward<HID, E, NT, A, NP1>(s_blob, L, s_ol, s_h1, s_out, od, P.activation);
    if (live) {
        if (fast) {
            e.ep_ret = s_stat[ol]; e.ep_len = s_stat[E + ol]; e.ep_cv = s_stat[2 * E + ol];
            ep_ret_sum = s_stat[3 * E + ol]; ep_len_sum = s_stat[4 * E + ol]; ep_cv_sum = s_stat[5 * E + ol]; ep_count = s_epc[ol];
        }
        env_store<SYS, R>(T, n, e);
        for (int k = 0; k < od; ++k) {
            const float v = s_ol[tile_idx<E>(k, ol)];
            T.st.cur_obs[(size_t)n * od + k] = v;
            if (P.out.d_last_obs) P.out.d_last_obs[(size_t)n * od + k] = v;
        }
        const float lval = boot_adv ? mlp_out<E, NW, A2M>(s_out2, s_blob2, L2, A2, ol) : mlp_out<E, NW, A>(s_out, s_blob, L, A, ol);
        if (P.out.d_last_val) P.out.d_last_val[n] = lval;
        if (P.out.d_ep_count) { P.out.d_ep_count[n] = ep_count; P.out.d_ep_return_sum[n] = ep_ret_sum; P.out.d_ep_length_sum[n] = ep_len_sum; P.out.d_ep_cviol_sum[n] = ep_cv_sum; }
        if (P.out.d_ret && P.out.d_adv) {
            // compute_returns_and_advantages (ppo_utils.py:374-400): float32 ret / td recursions, float64 GAE accumulator
            float ret = lval, vnext = lval; double adv = 0.0;
            constexpr int CH = 8;            // loads of CH steps are issued together: the scan itself is a serial chain
            float rr[CH], tv[CH], mm[CH], vv[CH];
            auto load_chunk = [&](int t1, float* r_, float* t_, float* m_, float* v_) {
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    if (j < t1) {
                        const size_t tn = (size_t)(t1 - 1 - j) * N + n;
                        r_[j] = P.out.d_rew[tn]; t_[j] = P.out.d_terminal_v[tn]; m_[j] = P.out.d_mask[tn]; v_[j] = P.out.d_v[tn];
                    }
                }
            };
            load_chunk(P.T, rr, tv, mm, vv);
            for (int t1 = P.T; t1 > 0; t1 -= CH) {
                const int cnt = t1 < CH ? t1 : CH;
                // the NEXT chunk's loads are issued before this chunk's serial chain (the stores below touch other steps): the L2 round
                // trip of a chunk used to sit between two chains
                float rn[CH], tn_[CH], mn[CH], vn[CH];
#ifndef SCG_GAE_PREFETCH
#define SCG_GAE_PREFETCH 1
#endif
                if (SCG_GAE_PREFETCH && t1 > CH) load_chunk(t1 - CH, rn, tn_, mn, vn);
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    if (j < cnt) {
                        const size_t tn = (size_t)(t1 - 1 - j) * N + n;
                        const float r = rr[j] + P.gamma * tv[j];
                        if (P.mc_off) pstore(P, P.out.d_rew + tn, r); else P.out.d_rew[tn] = r;   // rews += gamma * terminal_v, in place (ppo_utils.py:389)
                        const float m = mm[j], v = vv[j];
                        ret = r + P.gamma * m * ret;
                        if (!P.use_gae) adv = (double)(ret - v);
                        else { const float td = r + P.gamma * m * vnext - v; adv = adv * P.lam_d * P.gamma_d * (double)m + (double)td; }
                        if (P.mc_off) { pstore(P, P.out.d_ret + tn, ret); pstore(P, P.out.d_adv + tn, (float)adv); }   // one multicast store each: no tail copy
                        else { P.out.d_ret[tn] = ret; P.out.d_adv[tn] = (float)adv; }
                        vnext = v;
                    }
                }
                if (SCG_GAE_PREFETCH) {
#pragma unroll
                    for (int j = 0; j < CH; ++j) { rr[j] = rn[j]; tv[j] = tn_[j]; mm[j] = mn[j]; vv[j] = vn[j]; }
                } else if (t1 > CH) load_chunk(t1 - CH, rr, tv, mm, vv);
            }
        }
    }
    // fused all-gather: the scan above is one serial chain per env, so its three columns (rew, ret, adv) reach the peers afterwards,
    // copied by the whole CTA (rows of E consecutive floats per step)
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
