// scg_device.cuh -- device-side environment logic of the B200 rollout engine.
//
// One lane owns one environment for the physics / env bookkeeping; the integrator state lives
// in that lane's registers for the whole horizon.  Everything is templated on
//   SYS : SCG_SYS_CARTPOLE / QUAD1D / QUAD2D / QUAD3D
//   R   : double (SCG_PREC_F64, the reference's arithmetic: numpy / PyBullet double) or float.
//
// Reference semantics implemented here (paths relative to the reference root):
//   Quadrotor.step            safe_control_gym/envs/gym_pybullet_drones/quadrotor.py:394-445
//   _preprocess_control       quadrotor.py:722-747, quadrotor_utils.py:16-60
//   BaseAviary physics        safe_control_gym/envs/gym_pybullet_drones/base_aviary.py:232-286,364-384
//   CartPole.step             safe_control_gym/envs/gym_control/cartpole.py:238-264,479-583
//   reward/done/info          quadrotor.py:819-923, cartpole.py:611-696
//   BenchmarkEnv.after_step   safe_control_gym/envs/benchmark_env.py:447-502, extend_obs :422-445
//   constraints               safe_control_gym/envs/constraints.py:97-130,274,445-447
//   disturbances              safe_control_gym/envs/disturbances.py:54-259
//   reset / randomisation     quadrotor.py:328-392, cartpole.py:266-352, benchmark_env.py:237-268
// The 2D quadrotor is integrated in its plane (x, z, theta): PyBullet's full 3D update reduces to it
// exactly because rotation about a fixed axis commutes; the CPU oracle keeps the
// full quaternion form and the parity tests compare the two.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "scg.h"

#define SCG_PI 3.14159265358979323846

// -DSCG_PROF developer builds: fine-grained cycle accounting of the owner warp's critical path (lane 0 of the last warp of CTA 0).
// SCG_TK(i) adds the cycles since the previous tick to bucket i; each tick costs ~40 cycles itself.
#ifdef SCG_PROF
static __device__ long long scg_ticks[64];
static __shared__ long long scg_tk_last;
#define SCG_TK(i) do { if (threadIdx.x == blockDim.x - 32 && blockIdx.x == 0) { const long long c_ = clock64(); scg_ticks[i] += c_ - scg_tk_last; scg_tk_last = c_; } } while (0)
#else
#define SCG_TK(i) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------
// math wrappers
// ---------------------------------------------------------------------------------------------
template <typename R> struct Mth;
template <> struct Mth<double> {
    static __device__ __forceinline__ double sin_(double x) { return sin(x); }
    static __device__ __forceinline__ double cos_(double x) { return cos(x); }
    static __device__ __forceinline__ void sincos_(double x, double* s, double* c) { sincos(x, s, c); }
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
    static __device__ __forceinline__ double exp_(double x) { return exp(x); }
    static __device__ __forceinline__ double log_(double x) { return log(x); }
    static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
    static __device__ __forceinline__ double asin_(double x) { return asin(x); }
    static __device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
    static __device__ __forceinline__ double fmod_(double x, double y) { return fmod(x, y); }
    static __device__ __forceinline__ double pow_(double x, double y) { return pow(x, y); }
};
template <> struct Mth<float> {
    static __device__ __forceinline__ float sin_(float x) { return sinf(x); }
    static __device__ __forceinline__ float cos_(float x) { return cosf(x); }
    static __device__ __forceinline__ void sincos_(float x, float* s, float* c) { sincosf(x, s, c); }
    static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float exp_(float x) { return expf(x); }
    static __device__ __forceinline__ float log_(float x) { return logf(x); }
    static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
    static __device__ __forceinline__ float asin_(float x) { return asinf(x); }
    static __device__ __forceinline__ float atan2_(float y, float x) { return atan2f(y, x); }
    static __device__ __forceinline__ float fmod_(float x, float y) { return fmodf(x, y); }
    static __device__ __forceinline__ float pow_(float x, float y) { return powf(x, y); }
};

// normalization.py:8-10 (numpy floored modulo).  The modulo is the identity on [0, 2pi): the fmod path is cold and out of line.
template <typename R> __device__ __noinline__ R normalize_angle_slow(R y) {
    const R two_pi = (R)(2.0 * SCG_PI);
    y = Mth<R>::fmod_(y, two_pi);
    if (y < (R)0) y += two_pi;
    return y;
}
template <typename R> __device__ __forceinline__ R normalize_angle(R x) {
    const R two_pi = (R)(2.0 * SCG_PI);
    R y = x + (R)SCG_PI;
    if (!(y >= (R)0 && y < two_pi)) y = normalize_angle_slow<R>(y);
    return y - (R)SCG_PI;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10, stream layout shared with the CPU oracle: ctr = (slot, idx, global env, stream),
// key = (seed lo, seed hi).  Two draw slots per block (2 words each); uniform / randint use the first word, normal both.
// ---------------------------------------------------------------------------------------------
#define SCG_STREAM_EPISODE 1u
#define SCG_STREAM_STEP 2u
#define SCG_STREAM_POLICY 3u
#define SCG_STREAM_RESET 4u
#define SCG_STREAM_ADVERSARY 5u   // sampling noise of the RARL / RAP adversary policy

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double u01_d(uint32_t x) { return ((double)x + 0.5) * (1.0 / 4294967296.0); }

struct RngKey { uint32_t k0, k1, genv; };

// draw slot j of (stream, idx) = words 2(j&1), 2(j&1)+1 of block j>>1
__device__ __forceinline__ void rng_words(const RngKey& k, uint32_t stream, uint32_t idx, uint32_t slot, uint32_t& w0, uint32_t& w1) {
    uint32_t x[4]; philox4x32_10(slot >> 1, idx, k.genv, stream, k.k0, k.k1, x);
    w0 = (slot & 1u) ? x[2] : x[0]; w1 = (slot & 1u) ? x[3] : x[1];
}
template <typename R> __device__ __forceinline__ R normal_from_words(uint32_t w0, uint32_t w1) {
    if (sizeof(R) == 8) {
        return (R)(sqrt(-2.0 * log(u01_d(w0))) * cos(2.0 * SCG_PI * u01_d(w1)));
    } else {
        float u0 = (float)u01_d(w0), u1 = (float)u01_d(w1);
        // u0 may round to 1.0f (log -> 0) but never to 0: (x+0.5)/2^32 >= 1.16e-10
        return (R)(sqrtf(-2.0f * logf(u0)) * cospif(2.0f * u1));
    }
}
template <typename R> __device__ __forceinline__ R rng_uniform(const RngKey& k, uint32_t stream, uint32_t idx, uint32_t slot) {
    uint32_t w0, w1; rng_words(k, stream, idx, slot, w0, w1);
    return (R)u01_d(w0);
}
template <typename R> __device__ __forceinline__ R rng_normal(const RngKey& k, uint32_t stream, uint32_t idx, uint32_t slot) {
    uint32_t w0, w1; rng_words(k, stream, idx, slot, w0, w1);
    return normal_from_words<R>(w0, w1);
}
__device__ __forceinline__ int rng_randint(const RngKey& k, uint32_t stream, uint32_t idx, uint32_t slot, uint32_t n) {
    uint32_t w0, w1; rng_words(k, stream, idx, slot, w0, w1);
    return (int)__umulhi(w0, n);
}

// ---------------------------------------------------------------------------------------------
// device-side view of one handle
// ---------------------------------------------------------------------------------------------
template <int SYS> struct SysDims;
// NADV: widest adversary action (= dim of the disturbance mode it drives: action -> NU, dynamics -> 2 / 1 / 2 / 3; benchmark_env.py:289-295)
#ifndef SCG_CP_CARRY
#define SCG_CP_CARRY 0      // measured (gpurun_out/r2_ab25.log): carrying sin / cos theta in the cart-pole state is 3.8 % SLOWER on C2 than one exact sincos per control step
#endif
template <> struct SysDims<SCG_SYS_CARTPOLE> { static constexpr int NX = 4, NU = 1, PS = SCG_CP_CARRY ? 6 : 4, NADV = 2; };   // x, xdot, theta, thetadot, sin theta, cos theta
template <> struct SysDims<SCG_SYS_QUAD1D> { static constexpr int NX = 2, NU = 1, PS = 4, NADV = 1; };   // x, vx, z, vz (x only carries the init quirk)
template <> struct SysDims<SCG_SYS_QUAD2D> { static constexpr int NX = 6, NU = 2, PS = 8, NADV = 2; };   // x, vx, z, vz, theta, omega_y, sin theta, cos theta
template <> struct SysDims<SCG_SYS_QUAD3D> { static constexpr int NX = 12, NU = 4, PS = 13, NADV = 4; }; // pos3, quat4, vel3, omega_world3

#define SCG_PS_MAX 13
#define SCG_NINV 4

// Persistent per-env state, SoA in one device allocation ("state blob", checkpointable as one memcpy).
struct DevState {
    double* phys;        // [SCG_PS_MAX][N]
    double* inertial;    // [SCG_MAX_INERTIAL][N]
    float* cur_obs;      // [N][obs_dim] observation to feed the policy next
    int32_t* ctrl_step;  // [N]
    int32_t* episode;    // [N] resets done so far
    uint32_t* total_steps; // [N]
    int32_t* initial_reset; // [N]
    int32_t* dist_off;   // [3*SCG_MAX_DIST][N]
    float* ep_acc;       // [3][N] running episode return / length / violations
    float* adv;          // [SCG_MAX_ADV][N] adversary action after clip / scale / offset (benchmark_env.py:223-224)
    int32_t* adv_set;    // [N] 1 = adv holds an action (self.adv_action is not None)
};

struct DevTables {
    const scg_config* cfg;     // device copy
    const double* x_goal;      // [L][nx]
    int32_t L;
    int32_t n_con;
    const int32_t* con_kind;   // [n_con]
    const int32_t* con_strict; // [n_con]
    const int32_t* con_ptr;    // [n_con+1]
    const int32_t* con_col;    // [nnz]
    const double* con_val;     // [nnz]
    const double* con_b;       // [n_con]
    const double* ep_table;    // [K][N][P] or null
    int32_t K, P;
    const double* noise_table; // [Tn][N][D] or null
    int32_t Tn, D;
    int32_t n_off_slots;
    int32_t obs_dim;
    int32_t N;
    DevState st;
};

// Everything the owner lane keeps in registers for one env.
template <int SYS, typename R> struct EnvRegs {
    R ps[SysDims<SYS>::PS];
    R inertial[SCG_MAX_INERTIAL];
    R inv[SCG_NINV];     // reciprocals of the inertial properties the integrator divides by (2D: 1/M, 1/Iyy; 3D: 1/M, 1/Ixx, 1/Iyy, 1/Izz)
    int ctrl_step;
    int episode;
    uint32_t total_steps;
    float ep_ret, ep_len, ep_cv;
};

template <int SYS, typename R> __device__ __forceinline__ void env_derived(EnvRegs<SYS, R>& e) {
    e.inv[0] = (R)1 / e.inertial[0];
    e.inv[1] = (SYS == SCG_SYS_QUAD2D) ? (R)1 / e.inertial[2] : (SYS == SCG_SYS_QUAD3D) ? (R)1 / e.inertial[1] : (R)0;
    e.inv[2] = (SYS == SCG_SYS_QUAD3D) ? (R)1 / e.inertial[2] : (R)0;
    e.inv[3] = (SYS == SCG_SYS_QUAD3D) ? (R)1 / e.inertial[3] : (R)0;
}
template <int SYS, typename R>
__device__ __forceinline__ void env_load(const DevTables& T, int n, EnvRegs<SYS, R>& e) {
    const int N = T.N;
#pragma unroll
    for (int j = 0; j < SysDims<SYS>::PS; ++j) e.ps[j] = (R)T.st.phys[(size_t)j * N + n];
#pragma unroll
    for (int j = 0; j < SCG_MAX_INERTIAL; ++j) e.inertial[j] = (R)T.st.inertial[(size_t)j * N + n];
    env_derived<SYS, R>(e);
    e.ctrl_step = T.st.ctrl_step[n]; e.episode = T.st.episode[n]; e.total_steps = T.st.total_steps[n];
    e.ep_ret = T.st.ep_acc[n]; e.ep_len = T.st.ep_acc[N + n]; e.ep_cv = T.st.ep_acc[2 * N + n];
}
template <int SYS, typename R>
__device__ __forceinline__ void env_store(const DevTables& T, int n, const EnvRegs<SYS, R>& e) {
    const int N = T.N;
#pragma unroll
    for (int j = 0; j < SysDims<SYS>::PS; ++j) T.st.phys[(size_t)j * N + n] = (double)e.ps[j];
#pragma unroll
    for (int j = 0; j < SCG_MAX_INERTIAL; ++j) T.st.inertial[(size_t)j * N + n] = (double)e.inertial[j];
    T.st.ctrl_step[n] = e.ctrl_step; T.st.episode[n] = e.episode; T.st.total_steps[n] = e.total_steps;
    T.st.ep_acc[n] = e.ep_ret; T.st.ep_acc[N + n] = e.ep_len; T.st.ep_acc[2 * N + n] = e.ep_cv;
}

// ---------------------------------------------------------------------------------------------
// noise source for one control step of one env (table row or Philox stream)
// ---------------------------------------------------------------------------------------------
template <typename R> struct NoiseSrc {
    const double* row;   // table row (may be null -> zeros) in table mode
    bool philox;
    RngKey key;
    uint32_t stream, idx;
    int slot;
    __device__ __forceinline__ R u01() {
        R v = philox ? rng_uniform<R>(key, stream, idx, (uint32_t)slot) : (row ? (R)row[slot] : (R)0);
        ++slot; return v;
    }
    __device__ __forceinline__ R normal() {
        R v = philox ? rng_normal<R>(key, stream, idx, (uint32_t)slot) : (row ? (R)row[slot] : (R)0);
        ++slot; return v;
    }
};

// DisturbanceList.apply for one mode (disturbances.py:54-62)
template <typename R>
__device__ __noinline__ void apply_disturbances(const scg_config* c, const int32_t* dist_off, int N_stride, int n,
                                                int mode, int ctrl_step, int pyb_step, R* target, int tdim, NoiseSrc<R>& ns) {
    const int nd = c->n_dist[mode];
    for (int j = 0; j < nd; ++j) {
        const scg_disturbance* d = &c->dist[mode][j];
        const int kind = d->kind, dim = d->dim;
        int off = d->step_offset;
        if ((kind == SCG_DIST_IMPULSE || kind == SCG_DIST_STEP) && off < 0) off = dist_off[(size_t)(mode * SCG_MAX_DIST + j) * N_stride + n];
        R common = (R)0;
        if (kind == SCG_DIST_IMPULSE) {          // disturbances.py:108-123
            if (ctrl_step >= off) {
                int peak = (int)((double)off + d->duration / 2);
                int po = abs(ctrl_step - peak);
                common = ((double)po < d->duration / 2) ? (R)(d->magnitude * pow(d->decay_rate, (double)po)) : (R)0;
            }
        } else if (kind == SCG_DIST_STEP) {      // disturbances.py:153-163
            common = (ctrl_step >= off) ? (R)d->magnitude : (R)0;
        }
        for (int k = 0; k < dim; ++k) {
            R nz = common;
            if (kind == SCG_DIST_UNIFORM) nz = (R)d->low[k] + (R)(d->high[k] - d->low[k]) * ns.u01();       // :187-192
            else if (kind == SCG_DIST_WHITE) nz = (R)d->std[k] * ns.normal();                               // :215-223
            else if (kind == SCG_DIST_PERIODIC) {                                                           // :249-259
                R phase = (R)(-SCG_PI) + (R)(2.0 * SCG_PI) * ns.u01();
                R t = (R)((double)pyb_step * (1.0 / (double)c->pyb_freq));
                nz = (R)d->scale * Mth<R>::sin_((R)(2.0 * SCG_PI * d->frequency) * t + phase);
            }
            if (d->has_mask && kind != SCG_DIST_PERIODIC) nz *= (R)d->mask[k];
            if (k < tdim) target[k] += nz;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// hot per-step constants, staged once per CTA into shared memory (the scg_config in global memory is only
// touched again on the rare paths: disturbances and resets)
// ---------------------------------------------------------------------------------------------
struct ConRow { double val; double b; int32_t col; int32_t flags; };   // flags: bit0 strict, bit1 abs

template <typename R> struct HotCfg {
    int steps, ctrl_steps, task, cost, normalized_action, rew_exponential, done_on_oob, done_on_violation, use_penalty;
    int obs_goal_horizon, obs_wrap, philox, has_dist[3], n_con, con_fast, L;
    R dt, hover, act_scale, goal_tol, penalty, x_thr, th_thr;
    R gravity, kf, inv_nm_kf, pwm_c, pwm_s, inv_pwm_s, pwm_min, pwm_max, arm, km_over_kf, pole_w2;
    R thr_lo, thr_hi, inv_nm;   // thrust per input at the PWM limits (with a margin), 1 / motors per input
    // non-default physics (base_aviary.py:386-460) and the adversary input (benchmark_env.py:216-228): only read by the EXT instantiations
    int physics, adv_mode;
    R gnd_coeff, prop_radius, h_clip, drag_xy, drag_z, inv_kf;
    R act_low[SCG_MAX_NU], act_high[SCG_MAX_NU], u_goal[SCG_MAX_NU], rew_act_w[SCG_MAX_NU], Rw[SCG_MAX_NU];
    R rew_state_w[SCG_MAX_NX], mse_w[SCG_MAX_NX], Qw[SCG_MAX_NX];
    double state_low[SCG_MAX_NX], state_high[SCG_MAX_NX];   // float32-rounded bounds, compared in double like numpy does
    // reset / randomisation (rare path, but a warp executes it whenever ANY of its lanes terminates)
    uint32_t key0, key1, genv0;
    int randomized_init, randomized_inertial, n_off_draws;
    int init_kind[SCG_MAX_NX], inertial_kind[SCG_MAX_INERTIAL];
    double init_nominal[SCG_MAX_NX], init_a[SCG_MAX_NX], init_b[SCG_MAX_NX];
    double inertial_nominal[SCG_MAX_INERTIAL], inertial_a[SCG_MAX_INERTIAL], inertial_b[SCG_MAX_INERTIAL];
    ConRow rows[SCG_MAX_CON_ROWS];
};

// executed by ONE thread of the CTA, followed by a __syncthreads() in the caller
template <int SYS, typename R>
__device__ __noinline__ void hot_load(const DevTables& T, HotCfg<R>* h) {
    constexpr int NU = SysDims<SYS>::NU;
    const scg_config* c = T.cfg;
    h->steps = c->pyb_freq / c->ctrl_freq; h->ctrl_steps = c->ctrl_steps; h->task = c->task; h->cost = c->cost;
    h->normalized_action = c->normalized_action; h->rew_exponential = c->rew_exponential; h->done_on_oob = c->done_on_out_of_bound;
    h->done_on_violation = c->done_on_violation; h->use_penalty = c->use_constraint_penalty;
    h->obs_goal_horizon = (c->cost == SCG_COST_RL_REWARD) ? c->obs_goal_horizon : 0; h->obs_wrap = c->obs_wrap_angle;
    h->philox = c->rng_mode == SCG_RNG_PHILOX;
    for (int m = 0; m < 3; ++m) h->has_dist[m] = c->n_dist[m] > 0;
    h->L = T.L;
    h->dt = (R)(1.0 / (double)c->pyb_freq); h->hover = (R)c->hover_thrust; h->act_scale = (R)c->norm_act_scale;
    h->goal_tol = (R)c->goal_tolerance; h->penalty = (R)c->constraint_penalty; h->x_thr = (R)c->x_threshold; h->th_thr = (R)c->theta_threshold;
    h->gravity = (R)c->gravity; h->kf = (R)c->kf; h->inv_nm_kf = (R)(1.0 / ((double)(4 / NU) * c->kf));
    h->pwm_c = (R)c->pwm2rpm_const; h->pwm_s = (R)c->pwm2rpm_scale; h->inv_pwm_s = (R)(1.0 / c->pwm2rpm_scale);
    h->pwm_min = (R)c->pwm_min; h->pwm_max = (R)c->pwm_max; h->arm = (R)c->arm; h->km_over_kf = (R)(c->km / c->kf);
    h->pole_w2 = (R)(c->pole_width * c->pole_width);
    h->physics = c->physics; h->adv_mode = c->adversary_mode;
    h->gnd_coeff = (R)c->gnd_eff_coeff; h->prop_radius = (R)c->prop_radius; h->h_clip = (R)c->gnd_eff_h_clip;
    h->drag_xy = (R)c->drag_coeff_xy; h->drag_z = (R)c->drag_coeff_z; h->inv_kf = (R)(1.0 / c->kf);
    {
        const double nmkf = (double)(4 / NU) * c->kf;
        const double rlo = c->pwm2rpm_scale * c->pwm_min + c->pwm2rpm_const, rhi = c->pwm2rpm_scale * c->pwm_max + c->pwm2rpm_const;
        h->thr_lo = (R)(rlo * rlo * nmkf * (1.0 + 1e-6)); h->thr_hi = (R)(rhi * rhi * nmkf * (1.0 - 1e-6));
        h->inv_nm = (R)(1.0 / (double)(4 / NU));
    }
    for (int k = 0; k < SCG_MAX_NU; ++k) {
        h->act_low[k] = (R)c->act_low[k]; h->act_high[k] = (R)c->act_high[k]; h->u_goal[k] = (R)c->u_goal[k];
        h->rew_act_w[k] = (R)c->rew_act_weight[k]; h->Rw[k] = (R)c->R[k];
    }
    for (int k = 0; k < SCG_MAX_NX; ++k) {
        h->rew_state_w[k] = (R)c->rew_state_weight[k]; h->mse_w[k] = (R)c->mse_weight[k]; h->Qw[k] = (R)c->Q[k];
        h->state_low[k] = c->state_low[k]; h->state_high[k] = c->state_high[k];
    }
    h->key0 = (uint32_t)c->seed; h->key1 = (uint32_t)(c->seed >> 32); h->genv0 = (uint32_t)c->env_index_offset;
    h->randomized_init = c->randomized_init; h->randomized_inertial = c->randomized_inertial; h->n_off_draws = T.n_off_slots;
    for (int k = 0; k < SCG_MAX_NX; ++k) { h->init_kind[k] = c->init_rand_kind[k]; h->init_nominal[k] = c->init_nominal[k]; h->init_a[k] = c->init_rand_a[k]; h->init_b[k] = c->init_rand_b[k]; }
    for (int k = 0; k < SCG_MAX_INERTIAL; ++k) { h->inertial_kind[k] = c->inertial_rand_kind[k]; h->inertial_nominal[k] = c->inertial_nominal[k]; h->inertial_a[k] = c->inertial_rand_a[k]; h->inertial_b[k] = c->inertial_rand_b[k]; }
    h->n_con = T.n_con; h->con_fast = 1;
    for (int r = 0; r < T.n_con; ++r) {
        const int p0 = T.con_ptr[r], p1 = T.con_ptr[r + 1];
        if (p1 - p0 > 1 || T.con_kind[r] == SCG_CON_QUAD) h->con_fast = 0;
        ConRow row; row.col = p1 > p0 ? T.con_col[p0] : 0; row.val = p1 > p0 ? T.con_val[p0] : 0.0; row.b = T.con_b[r];
        row.flags = (T.con_strict[r] ? 1 : 0) | (T.con_kind[r] == SCG_CON_ABS ? 2 : 0);
        h->rows[r] = row;
    }
}

// small-angle sine / cosine (|x| <= 0.05): Taylor series, error below 1 ulp of the result scale in both precisions
template <typename R> __device__ __forceinline__ void sincos_small(R x, R* s, R* c) {
    const R x2 = x * x;
    if (sizeof(R) == 8) {
        R ps = (R)(-1.0 / 5040.0) + x2 * (R)(1.0 / 362880.0);
        ps = (R)(1.0 / 120.0) + x2 * ps; ps = (R)(-1.0 / 6.0) + x2 * ps;
        *s = x + x * x2 * ps;
        R pc = (R)(-1.0 / 720.0) + x2 * ((R)(1.0 / 40320.0) - x2 * (R)(1.0 / 3628800.0));
        pc = (R)(1.0 / 24.0) + x2 * pc; pc = (R)(-0.5) + x2 * pc;
        *c = (R)1 + x2 * pc;
    } else {
        R ps = (R)(-1.0 / 6.0) + x2 * (R)(1.0 / 120.0);
        *s = x + x * x2 * ps;
        R pc = (R)(-0.5) + x2 * ((R)(1.0 / 24.0) - x2 * (R)(1.0 / 720.0));
        *c = (R)1 + x2 * pc;
    }
}
#define SCG_SMALL_ANGLE 0.05

// ---------------------------------------------------------------------------------------------
// physics
// ---------------------------------------------------------------------------------------------
// quadrotor.py:722-747 + quadrotor_utils.py:16-60: clipped thrust command(s) -> per-motor thrust kf*rpm^2
// Cold paths of the collector's owner warp are kept OUT OF LINE (SCG_COLD_OUTLINE): that warp is bound by the fetch of its straight-line code
// (24 % of its stall samples), and ptxas lays a rarely taken block out in source order, in the middle of the hot path.  Values travel by value
// (small structs), never by reference: an address-taken register array would move to local memory for its whole lifetime.
#ifndef SCG_COLD_OUTLINE
#define SCG_COLD_OUTLINE 1
#endif
#if SCG_COLD_OUTLINE
#define SCG_COLD __noinline__
#else
#define SCG_COLD __forceinline__
#endif
template <typename R, int NU> struct MotorF { R v[NU]; };
// full cmd2pwm -> clip -> pwm2rpm -> kf rpm^2 chain (quadrotor_utils.py:16-60): only needed when a command sits on the PWM limits
template <typename R, int NU> static __device__ SCG_COLD MotorF<R, NU> quad_motor_forces_clipped(const HotCfg<R>* h, MotorF<R, NU> thr) {
    MotorF<R, NU> o;
#pragma unroll
    for (int k = 0; k < NU; ++k) {
        R t = thr.v[k] < (R)0 ? (R)0 : thr.v[k];
        R pwm = (Mth<R>::sqrt_(t * h->inv_nm_kf) - h->pwm_c) * h->inv_pwm_s;
        pwm = pwm < h->pwm_min ? h->pwm_min : pwm;
        pwm = pwm > h->pwm_max ? h->pwm_max : pwm;
        R rpm = h->pwm_s * pwm + h->pwm_c;
        o.v[k] = rpm * rpm * h->kf;
    }
    return o;
}
template <int SYS, typename R>
__device__ __forceinline__ void quad_motor_forces(const HotCfg<R>* h, const R* thrust, R f[4]) {
    constexpr int NU = SysDims<SYS>::NU;
    R mf[NU];
    // cmd2pwm -> clip -> pwm2rpm -> kf rpm^2 is the identity thrust/n_motor (up to a few ulp) while the PWM stays strictly inside its
    // bounds -- always, unless the command sits on the action bounds, which are those PWM limits in thrust units.  The square root is
    // only taken in that boundary case (warp-uniformly rare): thr_lo / thr_hi are the thrusts of the PWM limits with a 1e-6 margin.
    bool inside = true;
#pragma unroll
    for (int k = 0; k < NU; ++k) inside = inside && thrust[k] > h->thr_lo && thrust[k] < h->thr_hi;
    if (inside) {
#pragma unroll
        for (int k = 0; k < NU; ++k) mf[k] = thrust[k] * h->inv_nm;
    } else {
        MotorF<R, NU> t;
#pragma unroll
        for (int k = 0; k < NU; ++k) t.v[k] = thrust[k];
        t = quad_motor_forces_clipped<R, NU>(h, t);
#pragma unroll
        for (int k = 0; k < NU; ++k) mf[k] = t.v[k];
    }
    if (NU == 1) { f[0] = f[1] = f[2] = f[3] = mf[0]; }
    else if (NU == 2) { f[0] = mf[0]; f[1] = mf[1 % NU]; f[2] = mf[1 % NU]; f[3] = mf[0]; }   // quadrotor_utils.py:39
    else { f[0] = mf[0]; f[1] = mf[1 % NU]; f[2] = mf[2 % NU]; f[3] = mf[3 % NU]; }
}

// Bullet btQuaternion::getEulerZYX (pybullet.getEulerFromQuaternion, base_aviary.py:329)
template <typename R> __device__ __forceinline__ void rpy_from_quat(const R q[4], R rpy[3]) {
    R sqx = q[0] * q[0], sqy = q[1] * q[1], sqz = q[2] * q[2], squ = q[3] * q[3];
    R sarg = (R)-2 * (q[0] * q[2] - q[3] * q[1]);
    if (sarg <= (R)-0.99999) { rpy[1] = (R)(-0.5 * SCG_PI); rpy[0] = 0; rpy[2] = (R)2 * Mth<R>::atan2_(q[0], -q[1]); }
    else if (sarg >= (R)0.99999) { rpy[1] = (R)(0.5 * SCG_PI); rpy[0] = 0; rpy[2] = (R)2 * Mth<R>::atan2_(-q[0], q[1]); }
    else {
        rpy[1] = Mth<R>::asin_(sarg);
        rpy[0] = Mth<R>::atan2_((R)2 * (q[1] * q[2] + q[3] * q[0]), squ - sqx - sqy + sqz);
        rpy[2] = Mth<R>::atan2_((R)2 * (q[0] * q[1] + q[3] * q[2]), squ + sqx - sqy - sqz);
    }
}
// body -> world rotation of a UNIT quaternion (the integrator renormalises every sub-step, like Bullet)
template <typename R> __device__ __forceinline__ void rot_from_unit_quat(const R q[4], R M[9]) {
    R xs = q[0] + q[0], ys = q[1] + q[1], zs = q[2] + q[2];
    R wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
    R xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs;
    R yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
    M[0] = (R)1 - (yy + zz); M[1] = xy - wz; M[2] = xz + wy;
    M[3] = xy + wz; M[4] = (R)1 - (xx + zz); M[5] = yz - wx;
    M[6] = xz - wy; M[7] = yz + wx; M[8] = (R)1 - (xx + yy);
}

// advance one control step = `steps` Bullet sub-steps (base_aviary.py:243-282 / cartpole.py:552-583):
// semi-implicit Euler per sub-step -- velocities with the accelerations at the current pose, then poses with the NEW velocities.
template <int SYS, typename R> struct Physics;

template <typename R> struct SinCos2 { R sd, cd, se, ce; };
template <typename R> static __device__ SCG_COLD SinCos2<R> sincos_pair_exact(R d0, R eps) {     // large per-sub-step angle increments: exact evaluation
    SinCos2<R> o;
    Mth<R>::sincos_(d0, &o.sd, &o.cd);
    Mth<R>::sincos_(eps, &o.se, &o.ce);
    return o;
}

template <typename R> struct Physics<SCG_SYS_CARTPOLE, R> {
    // ps = x, xdot, theta, thetadot, sin theta, cos theta ; inertial = pole_length, cart_mass, pole_mass
    // (sin theta, cos theta) are carried in the state and rotated by every sub-step's angle increment (see Physics<QUAD2D>): no sincos(theta)
    // per control step on the owner lanes' critical path
    static __device__ __forceinline__ void advance(const HotCfg<R>* h, EnvRegs<SCG_SYS_CARTPOLE, R>& e, const R* force, const R* tap, bool has_tap) {
        const R l = e.inertial[0], M = e.inertial[1], m = e.inertial[2], g = h->gravity, dt = h->dt;
        const R Ic = m * (R)(1.0 / 12.0) * (((R)2 * l) * ((R)2 * l) + h->pole_w2);   // box inertia PyBullet recomputes (cartpole.py:301-304,318-322)
        const R a11 = M + m, a22 = Ic + m * l * l, ml = m * l, mgl = m * g * l;
        R x = e.ps[0], xd = e.ps[1], th = e.ps[2], thd = e.ps[3];
#if SCG_CP_CARRY
        R st = e.ps[4], ct = e.ps[5];
#else
        R st, ct; Mth<R>::sincos_(th, &st, &ct);
#endif
        const int steps = h->steps;
        for (int s = 0; s < steps; ++s) {
            R Qx = force[0], Qth = (R)0;
            if (has_tap) { Qx += tap[0]; Qth += tap[0] * l * ct - tap[1] * l * st; }
            const R a12 = ml * ct;
            const R b1 = Qx + ml * st * thd * thd, b2 = Qth + mgl * st;
            const R idet = (R)1 / (a11 * a22 - a12 * a12);
            const R xdd = (a22 * b1 - a12 * b2) * idet, thdd = (a11 * b2 - a12 * b1) * idet;
            xd += xdd * dt; thd += thdd * dt;
            x += xd * dt;
            const R dth = thd * dt; th += dth;
            // sin/cos of the new angle: rotate by the (small) increment, exact re-evaluation otherwise
            if (!SCG_CP_CARRY && s + 1 >= steps) break;
            if (Mth<R>::abs_(dth) <= (R)SCG_SMALL_ANGLE) {
                R sd, cd; sincos_small<R>(dth, &sd, &cd);
                const R s2 = st * cd + ct * sd; ct = ct * cd - st * sd; st = s2;
            } else { const SinCos2<R> sc = sincos_pair_exact<R>(th, (R)0); st = sc.sd; ct = sc.cd; }
        }
        e.ps[0] = x; e.ps[1] = xd; e.ps[2] = th; e.ps[3] = thd;
#if SCG_CP_CARRY
        const R nrm = (R)1.5 - (R)0.5 * (st * st + ct * ct);           // back onto the unit circle
        e.ps[4] = st * nrm; e.ps[5] = ct * nrm;
#endif
    }
    static __device__ __forceinline__ void state(const EnvRegs<SCG_SYS_CARTPOLE, R>& e, R* s) {
        s[0] = e.ps[0]; s[1] = e.ps[1]; s[2] = e.ps[2]; s[3] = e.ps[3];
    }
    static __device__ __forceinline__ void set_init(EnvRegs<SCG_SYS_CARTPOLE, R>& e, const R* init) {
        e.ps[0] = init[0]; e.ps[1] = init[1]; e.ps[2] = init[2]; e.ps[3] = init[3];
#if SCG_CP_CARRY
        R s0, c0; Mth<R>::sincos_(init[2], &s0, &c0);
        e.ps[4] = s0; e.ps[5] = c0;
#endif
    }
};

// Non-default quadrotor physics (Physics.PYB_GND / PYB_DRAG / PYB_GND_DRAG_DW, base_aviary.py:243-268).  Per sub-step, with the kinematic
// information refreshed at its start (base_aviary.py:246-251):
//   ground effect (base_aviary.py:386-418): every propeller gains KF rpm^2 GND_EFF_COEFF (PROP_RADIUS / (4 h_i))^2 along its LINK z axis at
//     the prop link -- the thrust's own application point and direction, so f_i grows by that factor (the yaw torque does not) -- with h_i
//     the world height of prop link i clipped from below, while |roll| and |pitch| < pi/2;
//   drag (base_aviary.py:420-440): -DRAG_COEFF sum(2 pi rpm / 60) * v_world is rotated by the base rotation and then applied as a
//     LINK_FRAME force at the CoM link, i.e. rotated once more: F_world = R R (k v);
//   downwash (base_aviary.py:442-460): compares the drone with itself (delta_z = 0): never a force, PYB_DW == PYB.
template <typename R> struct PhysExt {
    bool gnd, drag; R kxy, kz;
    __device__ __forceinline__ PhysExt() : gnd(false), drag(false), kxy((R)0), kz((R)0) {}
    __device__ __forceinline__ PhysExt(const HotCfg<R>* h, const R f[4]) {
        gnd = h->physics == SCG_PHYS_PYB_GND || h->physics == SCG_PHYS_PYB_GND_DRAG_DW;
        drag = h->physics == SCG_PHYS_PYB_DRAG || h->physics == SCG_PHYS_PYB_GND_DRAG_DW;
        kxy = kz = (R)0;
        if (drag) {
            R srpm = (R)0;
#pragma unroll
            for (int i = 0; i < 4; ++i) srpm += Mth<R>::sqrt_(f[i] * h->inv_kf);       // rpm_i: f_i = KF rpm_i^2
            srpm *= (R)(2.0 * SCG_PI / 60.0);
            kxy = -h->drag_xy * srpm; kz = -h->drag_z * srpm;
        }
    }
    __device__ __forceinline__ bool any() const { return gnd || drag; }
    // thrust gain of a propeller at world height hh
    __device__ __forceinline__ R gain(const HotCfg<R>* h, R hh) const {
        hh = hh < h->h_clip ? h->h_clip : hh;
        const R q = h->prop_radius / ((R)4 * hh);
        return (R)1 + h->gnd_coeff * q * q;
    }
};

template <typename R> struct Physics<SCG_SYS_QUAD1D, R> {
    // vertical motion only: orientation stays identity, thrust along world z
    template <bool EXT = false>
    static __device__ __forceinline__ void advance(const HotCfg<R>* h, EnvRegs<SCG_SYS_QUAD1D, R>& e, const R f[4], const R* fd, bool has_dist) {
        const R m = e.inertial[0], dt = h->dt;
        if constexpr (EXT) {
            const PhysExt<R> px(h, f);
            if (px.any()) {                      // identity orientation: all four propellers at height z, roll = pitch = 0, R R (k v) = k v
                const R thrust0 = f[0] + f[1] + f[2] + f[3];
                R z = e.ps[2], vz = e.ps[3], x = e.ps[0], vx = e.ps[1];
                for (int s = 0; s < h->steps; ++s) {
                    R Fz = (px.gnd ? thrust0 * px.gain(h, z) : thrust0) - m * h->gravity, Fx = (R)0;
                    if (px.drag) { Fx = px.kxy * vx; Fz += px.kz * vz; }
                    if (has_dist) Fz += fd[0];
                    vx += Fx / m * dt; vz += Fz / m * dt; x += vx * dt; z += vz * dt;
                }
                e.ps[0] = x; e.ps[1] = vx; e.ps[2] = z; e.ps[3] = vz;
                return;
            }
        }
        R thrust = f[0] + f[1] + f[2] + f[3];
        R Fz = thrust - m * h->gravity; if (has_dist) Fz += fd[0];
        R az = Fz / m;
        R z = e.ps[2], vz = e.ps[3], x = e.ps[0], vx = e.ps[1];
        for (int s = 0; s < h->steps; ++s) { vz += az * dt; z += vz * dt; x += vx * dt; }
        e.ps[2] = z; e.ps[3] = vz; e.ps[0] = x;
    }
    static __device__ __forceinline__ void state(const EnvRegs<SCG_SYS_QUAD1D, R>& e, R* s) { s[0] = e.ps[2]; s[1] = e.ps[3]; }
    static __device__ __forceinline__ void set_init(EnvRegs<SCG_SYS_QUAD1D, R>& e, const R* init) {
        // quirk: 1D labels are init_x / init_x_dot and quadrotor.py:373-374 routes them to x; z starts at 0
        e.ps[0] = init[0]; e.ps[1] = init[1]; e.ps[2] = (R)0; e.ps[3] = (R)0;
    }
};

// pitch as pybullet.getEulerFromQuaternion reports it for q = (0, sin(th/2), 0, cos(th/2)): asin(sin th), clamped at +-pi/2 when
// |sin th| >= 0.99999.  Identical to th while |th| < ~1.5663, so only large angles come here (cold, out of line).
template <typename R> __device__ __noinline__ R pitch_of_theta(R th) {
    const R sarg = Mth<R>::sin_(th);
    return (sarg <= (R)-0.99999) ? (R)(-0.5 * SCG_PI) : (sarg >= (R)0.99999) ? (R)(0.5 * SCG_PI) : Mth<R>::asin_(sarg);
}

template <typename R> struct Physics<SCG_SYS_QUAD2D, R> {
    // ps = x, vx, z, vz, theta, omega_y, sin theta, cos theta ; inertial = M, Ixx, Iyy, Izz
    // Within a control step thrust and torque are constant, so omega grows linearly and the per-sub-step angle increment
    // delta_k = omega_k dt grows by the constant eps = alpha dt^2: sin/cos(theta_k) follow from two plane rotations per
    // sub-step.  (sin theta, cos theta) are part of the integrator state: they are rotated along with theta and carried from one control
    // step to the next (one Newton step per control step keeps them on the unit circle), so no sine / cosine is evaluated inside an
    // episode at all -- the exact sincos(theta) per control step was ~250 cycles of the owner lanes' critical path.  The phase error of the
    // rotation recurrence grows by < 1.2e-16 per sub-step: < 6e-13 over a 250-step episode, three orders below the parity tolerance (1e-9).
    // non-default physics: thrust and torque change with the pose every sub-step, so the angle is re-evaluated instead of rotated
    // (out of line and BY VALUE: a reference to the register-resident env would pin it in local memory for the whole kernel)
    static __device__ __noinline__ EnvRegs<SCG_SYS_QUAD2D, R> advance_ext(const HotCfg<R>* h, EnvRegs<SCG_SYS_QUAD2D, R> e, const R f0, const R f1, const R fdx, const R fdz,
                                                                          bool has_dist, const PhysExt<R> px) {
        const R dt = h->dt, inv_m = e.inv[0], a = h->arm;
        const R f[2] = {f0, f1}, fd[2] = {fdx, fdz};
        R x = e.ps[0], vx = e.ps[1], z = e.ps[2], vz = e.ps[3], th = e.ps[4], om = e.ps[5];
        for (int s = 0; s < h->steps; ++s) {
            R st, ct; Mth<R>::sincos_(th, &st, &ct);
            R f03 = f[0], f12 = f[1];                                  // f = [p0, p1, p1, p0] (quadrotor_utils.py:39)
            // q = (0, sin(th/2), 0, cos(th/2)): getEulerFromQuaternion gives roll = atan2(0, cos th), pitch = asin(sin th) away from the poles
            if (px.gnd && ct > (R)0 && Mth<R>::abs_(st) < (R)0.99999) { f03 *= px.gain(h, z - st * a); f12 *= px.gain(h, z + st * a); }
            const R tm = (R)2 * (f03 + f12) * inv_m, alpha = (R)2 * a * (f12 - f03) * e.inv[1];
            R ax = st * tm, az = ct * tm - h->gravity;
            if (px.drag) {                                             // R R (k v) with R = R_y(th)
                const R kvx = px.kxy * vx, kvz = px.kz * vz;
                const R d1x = ct * kvx + st * kvz, d1z = -st * kvx + ct * kvz;
                ax += (ct * d1x + st * d1z) * inv_m; az += (-st * d1x + ct * d1z) * inv_m;
            }
            if (has_dist) { ax += fd[0] * inv_m; az += fd[1] * inv_m; }
            vx += ax * dt; vz += az * dt; om += alpha * dt;
            x += vx * dt; z += vz * dt; th += om * dt;
        }
        e.ps[0] = x; e.ps[1] = vx; e.ps[2] = z; e.ps[3] = vz; e.ps[4] = th; e.ps[5] = om;
        R sf, cf; Mth<R>::sincos_(th, &sf, &cf);
        e.ps[6] = sf; e.ps[7] = cf;
        return e;
    }
    template <bool EXT = false>
    static __device__ __forceinline__ void advance(const HotCfg<R>* h, EnvRegs<SCG_SYS_QUAD2D, R>& e, const R f[4], const R* fd, bool has_dist) {
        if constexpr (EXT) {
            const PhysExt<R> px(h, f);
            if (px.any()) { e = advance_ext(h, e, f[0], f[1], has_dist ? fd[0] : (R)0, has_dist ? fd[1] : (R)0, has_dist, px); return; }
        }
        const R dt = h->dt;
        const R inv_m = e.inv[0];                                      // 1/M, 1/Iyy: per-episode constants (env_load / env_reset)
        const R thrust = f[0] + f[1] + f[2] + f[3];
        const R tau_y = h->arm * (-f[0] + f[1] + f[2] - f[3]);         // r x F with the prop offsets of cf2x.urdf:42-78
        const R alpha = tau_y * e.inv[1];
        const R tm = thrust * inv_m;
        R fx = (R)0, fz = -h->gravity;                                 // gravity enters as the force m g
        if (has_dist) { fx += fd[0] * inv_m; fz += fd[1] * inv_m; }
        R x = e.ps[0], vx = e.ps[1], z = e.ps[2], vz = e.ps[3], th = e.ps[4], om = e.ps[5];
        SCG_TK(3);
        R st = e.ps[6], ct = e.ps[7];
        const R adt = alpha * dt, eps = adt * dt;
        R sd, cd, se, ce;
        // delta_0 = omega dt (becomes delta_1 after the first eps rotation) and eps = alpha dt^2 are small: series, no range reduction
        const R d0 = om * dt;
        if (Mth<R>::abs_(d0) <= (R)SCG_SMALL_ANGLE && Mth<R>::abs_(eps) <= (R)SCG_SMALL_ANGLE) {
            sincos_small<R>(d0, &sd, &cd);
            sincos_small<R>(eps, &se, &ce);
        } else {
            const SinCos2<R> sc = sincos_pair_exact<R>(d0, eps);
            sd = sc.sd; cd = sc.cd; se = sc.se; ce = sc.ce;
        }
        SCG_TK(4);
        const R tmdt = tm * dt, fxdt = fx * dt, fzdt = fz * dt;
        const int steps = h->steps;
#ifndef SCG_Q2D_SUMS
#define SCG_Q2D_SUMS 1
#endif
#if SCG_Q2D_SUMS
        // The owner lanes' FP64 work is THROUGHPUT bound: one warp issues a double-precision operation every ~4 cycles (measured with the
        // phase ticks: the 16-operation sub-step takes 57 cycles, a blocked variant with shorter dependent chains but 17 operations per
        // sub-step was 20 % SLOWER).  So the sub-step only keeps what it cannot avoid -- the two plane rotations (8 operations) and four
        // additions -- and velocities, positions, angle and rate follow in closed form from the running sums of symplectic Euler:
        //   S1_j = sum_{k<j} s_k,  P = sum_{j<=n} S1_j:   v_n = v_0 + dt (tm S1_n + n f),   x_n = x_0 + dt (n v_0 + dt (tm P + f n (n + 1) / 2)).
        {
            R S1s = (R)0, S1c = (R)0, Ps = (R)0, Pc = (R)0;
#pragma unroll 4
            for (int s = 0; s < steps; ++s) {
                S1s += st; S1c += ct; Ps += S1s; Pc += S1c;
                const R sd2 = sd * ce + cd * se; cd = cd * ce - sd * se; sd = sd2;        // delta += eps
                const R st2 = st * cd + ct * sd; ct = ct * cd - st * sd; st = st2;        // theta += delta
            }
            const R n = (R)steps, t2 = (R)0.5 * n * (n + (R)1);
            x += dt * (n * vx + dt * (tm * Ps + fx * t2)); z += dt * (n * vz + dt * (tm * Pc + fz * t2));
            vx += dt * (tm * S1s + n * fx); vz += dt * (tm * S1c + n * fz);
            th += dt * (n * om + adt * t2); om += n * adt;
        }
#else
#pragma unroll 4
        for (int s = 0; s < steps; ++s) {
            vx += st * tmdt + fxdt; vz += ct * tmdt + fzdt; om += adt;
            x += vx * dt; z += vz * dt; th += om * dt;
            const R sd2 = sd * ce + cd * se; cd = cd * ce - sd * se; sd = sd2;        // delta += eps
            const R st2 = st * cd + ct * sd; ct = ct * cd - st * sd; st = st2;        // theta += delta
        }
#endif
        SCG_TK(5);
        const R nrm = (R)1.5 - (R)0.5 * (st * st + ct * ct);           // back onto the unit circle (|1 - norm| is O(1e-15) per control step)
        e.ps[0] = x; e.ps[1] = vx; e.ps[2] = z; e.ps[3] = vz; e.ps[4] = th; e.ps[5] = om; e.ps[6] = st * nrm; e.ps[7] = ct * nrm;
    }
    static __device__ __forceinline__ void state(const EnvRegs<SCG_SYS_QUAD2D, R>& e, R* s) {
        s[0] = e.ps[0]; s[1] = e.ps[1]; s[2] = e.ps[2]; s[3] = e.ps[3];
        R th = e.ps[4];
        if (Mth<R>::abs_(th) > (R)1.5) th = pitch_of_theta<R>(th);       // beyond the 85 deg bound: the episode is ending anyway
        s[4] = th; s[5] = e.ps[5];
    }
    static __device__ __forceinline__ void set_init(EnvRegs<SCG_SYS_QUAD2D, R>& e, const R* init) {
#pragma unroll
        for (int j = 0; j < 6; ++j) e.ps[j] = init[j];                 // quadrotor.py:373-377
        R s0, c0; Mth<R>::sincos_(init[4], &s0, &c0);
        e.ps[6] = s0; e.ps[7] = c0;
    }
};

template <typename R> struct Physics<SCG_SYS_QUAD3D, R> {
    // ps = pos[0..2], quat[3..6] (x,y,z,w), vel[7..9], omega_world[10..12]
    template <bool EXT = false>
    static __device__ void advance(const HotCfg<R>* h, EnvRegs<SCG_SYS_QUAD3D, R>& e, const R f[4], const R* fd, bool has_dist) {
        const R dt = h->dt;
        const R inv_m = e.inv[0];                                      // per-episode reciprocals (env_load / reset candidates)
        const R J0 = e.inertial[1], J1 = e.inertial[2], J2 = e.inertial[3];
        const R iJ0 = e.inv[1], iJ1 = e.inv[2], iJ2 = e.inv[3];
        const R a = h->arm;
        R tm = (f[0] + f[1] + f[2] + f[3]) * inv_m;
        R tbx = a * (f[0] + f[1] - f[2] - f[3]);
        R tby = a * (-f[0] + f[1] + f[2] - f[3]);
        const PhysExt<R> pe = EXT ? PhysExt<R>(h, f) : PhysExt<R>();
        const R tbz = h->km_over_kf * (-f[0] + f[1] - f[2] + f[3]);   // base_aviary.py:371-372: rpm^2*KM = f*KM/KF
        R px = e.ps[0], py = e.ps[1], pz = e.ps[2];
        R q[4] = {e.ps[3], e.ps[4], e.ps[5], e.ps[6]};
        R vx = e.ps[7], vy = e.ps[8], vz = e.ps[9];
        R wx = e.ps[10], wy = e.ps[11], wz = e.ps[12];
        R ax0 = (R)0, ay0 = (R)0, az0 = -h->gravity;
        if (has_dist) { ax0 += fd[0] * inv_m; ay0 += fd[1] * inv_m; az0 += fd[2] * inv_m; }
        const R half_dt = (R)0.5 * dt;
        for (int s = 0; s < h->steps; ++s) {
            R M[9]; rot_from_unit_quat<R>(q, M);
            R dax = (R)0, day = (R)0, daz = (R)0;                      // drag acceleration (EXT)
            if constexpr (EXT) {
                if (pe.gnd) {
                    // |roll| < pi/2 <=> the x argument of its atan2 is positive; |pitch| < pi/2 <=> away from the asin clamp (rpy_from_quat)
                    const R sarg = (R)-2 * (q[0] * q[2] - q[3] * q[1]), rollx = q[3] * q[3] - q[0] * q[0] - q[1] * q[1] + q[2] * q[2];
                    R fe[4] = {f[0], f[1], f[2], f[3]};
                    if (rollx > (R)0 && Mth<R>::abs_(sarg) < (R)0.99999) {
                        // prop link offsets (cf2x.urdf:42,54,66,78): (a,a) (-a,a) (-a,-a) (a,-a); world height = z + R[2][0] x + R[2][1] y
                        fe[0] *= pe.gain(h, pz + a * (M[6] + M[7])); fe[1] *= pe.gain(h, pz + a * (-M[6] + M[7]));
                        fe[2] *= pe.gain(h, pz + a * (-M[6] - M[7])); fe[3] *= pe.gain(h, pz + a * (M[6] - M[7]));
                    }
                    tm = (fe[0] + fe[1] + fe[2] + fe[3]) * inv_m;
                    tbx = a * (fe[0] + fe[1] - fe[2] - fe[3]); tby = a * (-fe[0] + fe[1] + fe[2] - fe[3]);
                }
                if (pe.drag) {
                    const R kx = pe.kxy * vx, ky = pe.kxy * vy, kz = pe.kz * vz;
                    const R d0 = M[0] * kx + M[1] * ky + M[2] * kz, d1 = M[3] * kx + M[4] * ky + M[5] * kz, d2 = M[6] * kx + M[7] * ky + M[8] * kz;
                    dax = (M[0] * d0 + M[1] * d1 + M[2] * d2) * inv_m; day = (M[3] * d0 + M[4] * d1 + M[5] * d2) * inv_m;
                    daz = (M[6] * d0 + M[7] * d1 + M[8] * d2) * inv_m;
                }
            }
            // body rates, Euler's equation with the gyroscopic term (btMultiBody m_useGyroTerm)
            const R bx = M[0] * wx + M[3] * wy + M[6] * wz;
            const R by = M[1] * wx + M[4] * wy + M[7] * wz;
            const R bz = M[2] * wx + M[5] * wy + M[8] * wz;
            const R jx = J0 * bx, jy = J1 * by, jz = J2 * bz;
            const R dbx = (tbx - (by * jz - bz * jy)) * iJ0;
            const R dby = (tby - (bz * jx - bx * jz)) * iJ1;
            const R dbz = (tbz - (bx * jy - by * jx)) * iJ2;
            if constexpr (EXT) { vx += (M[2] * tm + ax0 + dax) * dt; vy += (M[5] * tm + ay0 + day) * dt; vz += (M[8] * tm + az0 + daz) * dt; }
            else { vx += (M[2] * tm + ax0) * dt; vy += (M[5] * tm + ay0) * dt; vz += (M[8] * tm + az0) * dt; }
            wx += (M[0] * dbx + M[1] * dby + M[2] * dbz) * dt;
            wy += (M[3] * dbx + M[4] * dby + M[5] * dbz) * dt;
            wz += (M[6] * dbx + M[7] * dby + M[8] * dbz) * dt;
            px += vx * dt; py += vy * dt; pz += vz * dt;
            // exponential-map orientation update with the NEW world angular velocity (btMultiBody::stepPositionsMultiDof)
            const R w2 = wx * wx + wy * wy + wz * wz;
            R k, cw;
            if (w2 * half_dt * half_dt <= (R)(SCG_SMALL_ANGLE * SCG_SMALL_ANGLE)) {
                // sin(|w| dt/2)/|w| and cos(|w| dt/2) as even power series in |w| dt/2: no square root, no trig
                // (Bullet's own branch below 1e-3 rad/s, btTransformUtil, is this series truncated; the difference is < 1e-17)
                const R y2 = w2 * half_dt * half_dt;
                R ps, pc;
                if (sizeof(R) == 8) {
                    ps = (R)(-1.0 / 5040.0) + y2 * (R)(1.0 / 362880.0); ps = (R)(1.0 / 120.0) + y2 * ps; ps = (R)(-1.0 / 6.0) + y2 * ps;
                    pc = (R)(-1.0 / 720.0) + y2 * ((R)(1.0 / 40320.0) - y2 * (R)(1.0 / 3628800.0)); pc = (R)(1.0 / 24.0) + y2 * pc; pc = (R)(-0.5) + y2 * pc;
                } else {
                    ps = (R)(-1.0 / 6.0) + y2 * (R)(1.0 / 120.0);
                    pc = (R)(-0.5) + y2 * ((R)(1.0 / 24.0) - y2 * (R)(1.0 / 720.0));
                }
                k = half_dt * ((R)1 + y2 * ps); cw = (R)1 + y2 * pc;
            } else {
                R ang = Mth<R>::sqrt_(w2);
                if (ang * dt > (R)(0.25 * SCG_PI)) ang = (R)(0.25 * SCG_PI) / dt;   // ANGULAR_MOTION_THRESHOLD
                k = Mth<R>::sin_(half_dt * ang) / ang; cw = Mth<R>::cos_(half_dt * ang);
            }
            const R ax = wx * k, ay = wy * k, az = wz * k;
            const R n0 = cw * q[0] + ax * q[3] + ay * q[2] - az * q[1];
            const R n1 = cw * q[1] + ay * q[3] + az * q[0] - ax * q[2];
            const R n2 = cw * q[2] + az * q[3] + ax * q[1] - ay * q[0];
            const R n3 = cw * q[3] - ax * q[0] - ay * q[1] - az * q[2];
            // renormalise like Bullet.  q was unit and the increment is a unit quaternion up to the series truncation, so
            // |n|^2 = 1 + d with |d| tiny: 1/sqrt(1 + d) = 1 - d/2 + 3 d^2/8 - 5 d^3/16 (error < 1e-19 for |d| < 1e-5), no sqrt / divide
            const R d = (n0 * n0 + n1 * n1 + n2 * n2 + n3 * n3) - (R)1;
            R inn;
            if (Mth<R>::abs_(d) < (R)1e-5) inn = (R)1 + d * ((R)-0.5 + d * ((R)0.375 - d * (R)0.3125));
            else inn = (R)1 / Mth<R>::sqrt_((R)1 + d);
            q[0] = n0 * inn; q[1] = n1 * inn; q[2] = n2 * inn; q[3] = n3 * inn;
        }
        e.ps[0] = px; e.ps[1] = py; e.ps[2] = pz; e.ps[3] = q[0]; e.ps[4] = q[1]; e.ps[5] = q[2]; e.ps[6] = q[3];
        e.ps[7] = vx; e.ps[8] = vy; e.ps[9] = vz; e.ps[10] = wx; e.ps[11] = wy; e.ps[12] = wz;
    }
    static __device__ void state(const EnvRegs<SCG_SYS_QUAD3D, R>& e, R* s) {
        R q[4] = {e.ps[3], e.ps[4], e.ps[5], e.ps[6]};
        R rpy[3]; rpy_from_quat<R>(q, rpy);
        R M[9]; rot_from_unit_quat<R>(q, M);
        s[0] = e.ps[0]; s[1] = e.ps[7]; s[2] = e.ps[1]; s[3] = e.ps[8]; s[4] = e.ps[2]; s[5] = e.ps[9];
        s[6] = rpy[0]; s[7] = rpy[1]; s[8] = rpy[2];
        const R wx = e.ps[10], wy = e.ps[11], wz = e.ps[12];
        s[9] = M[0] * wx + M[3] * wy + M[6] * wz;      // R^T omega_world (quadrotor.py:795-797)
        s[10] = M[1] * wx + M[4] * wy + M[7] * wz;
        s[11] = M[2] * wx + M[5] * wy + M[8] * wz;
    }
    static __device__ void set_init(EnvRegs<SCG_SYS_QUAD3D, R>& e, const R* init) {
        e.ps[0] = init[0]; e.ps[7] = init[1]; e.ps[1] = init[2]; e.ps[8] = init[3]; e.ps[2] = init[4]; e.ps[9] = init[5];
        // pybullet.getQuaternionFromEuler (quadrotor.py:381)
        R hr = init[6] * (R)0.5, hp = init[7] * (R)0.5, hy = init[8] * (R)0.5;
        R sr, cr, sp, cp, sy, cy;
        Mth<R>::sincos_(hr, &sr, &cr); Mth<R>::sincos_(hp, &sp, &cp); Mth<R>::sincos_(hy, &sy, &cy);
        R q0 = sr * cp * cy - cr * sp * sy, q1 = cr * sp * cy + sr * cp * sy;
        R q2 = cr * cp * sy - sr * sp * cy, q3 = cr * cp * cy + sr * sp * sy;
        R nn = Mth<R>::sqrt_(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
        e.ps[3] = q0 / nn; e.ps[4] = q1 / nn; e.ps[5] = q2 / nn; e.ps[6] = q3 / nn;
        e.ps[10] = init[9]; e.ps[11] = init[10]; e.ps[12] = init[11];  // quadrotor.py:379: p,q,r passed as WORLD angular velocity
    }
};

// ---------------------------------------------------------------------------------------------
// observation / reset / step
// ---------------------------------------------------------------------------------------------
// X_GOAL row (clamped like quadrotor.py:837); GT = storage type of the table (double in global memory, R in the CTA's shared copy)
template <typename GT> __device__ __forceinline__ const GT* goal_row(const GT* xg, int L, int task, int nx, int idx) {
    if (task == SCG_TASK_STABILIZATION) return xg;
    idx = idx > L - 1 ? L - 1 : idx;
    return xg + idx * nx;
}

// _get_observation + extend_obs (quadrotor.py:777-817, cartpole.py:585-609, benchmark_env.py:422-445).
// The observation is handed element by element to `sink(k, value)` (global buffers or the CTA's shared tile).
// NOISE = false: the caller knows that no observation disturbance is configured (the collector's FAST schedule), the branch is compiled out.
template <int SYS, typename R, typename GT, bool NOISE = true, typename Sink>
__device__ __forceinline__ void make_obs(const DevTables& T, const HotCfg<R>* h, const GT* xg, const R* state, int n, int ctrl_step, bool at_reset,
                                         NoiseSrc<R>& ns, Sink sink) {
    constexpr int NX = SysDims<SYS>::NX;
    if (NOISE && h->has_dist[SCG_MODE_OBSERVATION]) {
        R o[NX];
#pragma unroll
        for (int k = 0; k < NX; ++k) o[k] = state[k];
        apply_disturbances<R>(T.cfg, T.st.dist_off, T.N, n, SCG_MODE_OBSERVATION, ctrl_step, ctrl_step * h->steps, o, NX, ns);
        if (SYS == SCG_SYS_CARTPOLE && h->obs_wrap) o[2] = normalize_angle<R>(o[2]);
#pragma unroll
        for (int k = 0; k < NX; ++k) sink(k, o[k]);
    } else {
#pragma unroll
        for (int k = 0; k < NX; ++k) {
            R v = state[k];
            if (SYS == SCG_SYS_CARTPOLE && k == 2 && h->obs_wrap) v = normalize_angle<R>(v);
            sink(k, v);
        }
    }
    // goal part: the row is read into registers BEFORE the first store -- sink and table may both be shared memory, and the
    // compiler must otherwise keep every load behind the previous store (a serial LDS -> convert -> STS chain per element)
    const int horizon = h->obs_goal_horizon;
    if (horizon > 0) {
        const bool tracking = h->task == SCG_TASK_TRAJ_TRACKING;
        const int next_step = at_reset ? 1 : ctrl_step + 2, Lr = h->L;
        for (int hh = 0; hh < (tracking ? horizon : 1); ++hh) {
            const GT* g = tracking ? goal_row<GT>(xg, Lr, SCG_TASK_TRAJ_TRACKING, NX, next_step + hh) : xg;
            GT gv[NX];
#pragma unroll
            for (int k = 0; k < NX; ++k) gv[k] = g[k];
#pragma unroll
            for (int k = 0; k < NX; ++k) sink(NX * (1 + hh) + k, (R)gv[k]);
        }
    }
}

// One additive randomisation sample (benchmark_env.py:260-267) from two Philox words.  Cold code, deliberately out of line:
// inlined per slot it used to be the largest part of the rollout kernel (float64 log / sqrt / cos per copy).
static __device__ __noinline__ double rand_draw(int kind, double a, double b, uint32_t w0, uint32_t w1) {
    if (kind == SCG_RAND_UNIFORM) return a + (b - a) * u01_d(w0);
    if (kind == SCG_RAND_NORMAL) return a + b * normal_from_words<double>(w0, w1);
    return 0.0;
}
// Random draws of episode `k` of env `n` (a pure function of (n, k): host table row or Philox stream EPISODE), i.e. the inertial
// properties and the initial state after the additive randomisation of benchmark_env.py:237-268 (quadrotor.py:346-372, cartpole.py:287-326).
// Disturbance step offsets are NOT drawn here (see env_reset).  Draw slot j = words 2(j&1), 2(j&1)+1 of Philox block j>>1.
// out[j * stride], j < SCG_MAX_INERTIAL: inertial properties; out[(init_row + i) * stride], i < NX: the initial state.
// One compact out-of-line copy.
// `staged_row`: the table row of (n, k) already staged in shared memory (TMA bulk copy issued by the caller), or nullptr.
template <int SYS, typename R>
__device__ __noinline__ void reset_draw(const DevTables& T, const HotCfg<R>* h, int n, int k, double* out, int stride, int init_row,
                                        const double* staged_row = nullptr) {
    constexpr int NX = SysDims<SYS>::NX;
    const bool philox = h->philox != 0;
    const uint32_t genv = h->genv0 + (uint32_t)n;
    const double* row = staged_row ? staged_row : (!philox && T.ep_table) ? T.ep_table + ((size_t)(k % T.K) * T.N + n) * T.P : nullptr;
    const int slot0 = h->n_off_draws;
    const int n_inertial = (SYS == SCG_SYS_CARTPOLE) ? 3 : 4;
    uint32_t x0 = 0, x1 = 0, x2 = 0, x3 = 0; int have = -1;   // Philox block currently held
#pragma unroll 1
    for (int j = 0; j < SCG_MAX_INERTIAL + NX; ++j) {
        const bool inertial = j < SCG_MAX_INERTIAL;
        const int i = inertial ? j : j - SCG_MAX_INERTIAL;
        double v = inertial ? h->inertial_nominal[i] : h->init_nominal[i];
        const bool on = inertial ? (h->randomized_inertial && i < n_inertial) : (h->randomized_init != 0);
        if (on) {
            const int slot = slot0 + j;
            if (row) v += row[slot];
            else if (philox) {
                const int kind = inertial ? h->inertial_kind[i] : h->init_kind[i];
                if (kind != SCG_RAND_NONE) {
                    if (have != (slot >> 1)) {
                        have = slot >> 1;
                        uint32_t x[4]; philox4x32_10((uint32_t)have, (uint32_t)k, genv, SCG_STREAM_EPISODE, h->key0, h->key1, x);
                        x0 = x[0]; x1 = x[1]; x2 = x[2]; x3 = x[3];
                    }
                    const uint32_t w0 = (slot & 1) ? x2 : x0, w1 = (slot & 1) ? x3 : x1;
                    const double a = inertial ? h->inertial_a[i] : h->init_a[i], b = inertial ? h->inertial_b[i] : h->init_b[i];
                    if (kind == SCG_RAND_UNIFORM) v += a + (b - a) * u01_d(w0);
                    else v += rand_draw(kind, a, b, w0, w1);
                }
            }
        }
        out[(inertial ? i : init_row + i) * stride] = v;
    }
}

// ONE draw slot j of reset_draw (Philox mode only): the warp-cooperative form -- 32 lanes prepare one env's episode, lane j its slot j -- of the
// collector's reset-candidate helper.  Same stream layout, same arithmetic, same output rows as reset_draw.
template <int SYS, typename R>
__device__ __forceinline__ void reset_draw_slot(const HotCfg<R>* h, int n, int k, int j, double* out, int stride, int init_row) {
    constexpr int NX = SysDims<SYS>::NX;
    if (j >= SCG_MAX_INERTIAL + NX) return;
    const bool inertial = j < SCG_MAX_INERTIAL;
    const int i = inertial ? j : j - SCG_MAX_INERTIAL;
    const int n_inertial = (SYS == SCG_SYS_CARTPOLE) ? 3 : 4;
    double v = inertial ? h->inertial_nominal[i] : h->init_nominal[i];
    const bool on = inertial ? (h->randomized_inertial && i < n_inertial) : (h->randomized_init != 0);
    if (on) {
        const int kind = inertial ? h->inertial_kind[i] : h->init_kind[i];
        if (kind != SCG_RAND_NONE) {
            const int slot = h->n_off_draws + j;
            uint32_t x[4]; philox4x32_10((uint32_t)(slot >> 1), (uint32_t)k, h->genv0 + (uint32_t)n, SCG_STREAM_EPISODE, h->key0, h->key1, x);
            const uint32_t w0 = (slot & 1) ? x[2] : x[0], w1 = (slot & 1) ? x[3] : x[1];
            const double a = inertial ? h->inertial_a[i] : h->init_a[i], b = inertial ? h->inertial_b[i] : h->init_b[i];
            if (kind == SCG_RAND_UNIFORM) v += a + (b - a) * u01_d(w0);
            else v += rand_draw(kind, a, b, w0, w1);
        }
    }
    out[(inertial ? i : init_row + i) * stride] = v;
}

// reset() of one env (quadrotor.py:328-392, cartpole.py:266-352); the caller builds the observation with make_obs(at_reset).
template <int SYS, typename R>
__device__ __forceinline__ void env_reset(const DevTables& T, const HotCfg<R>* h, int n, EnvRegs<SYS, R>& e, R* state) {
    constexpr int NX = SysDims<SYS>::NX;
    const int k = e.episode;
    e.ctrl_step = 0;
    if (h->n_off_draws > 0) {
        // disturbance reset draws (benchmark_env.py:336-337, disturbances.py:102-106,148-151)
        const scg_config* c = T.cfg;
        const bool philox = h->philox != 0;
        RngKey key; key.k0 = h->key0; key.k1 = h->key1; key.genv = h->genv0 + (uint32_t)n;
        const double* row = (!philox && T.ep_table) ? T.ep_table + ((size_t)(k % T.K) * T.N + n) * T.P : nullptr;
        int slot = 0;
        for (int m = 0; m < 3; ++m)
            for (int j = 0; j < c->n_dist[m]; ++j) {
                const scg_disturbance* d = &c->dist[m][j];
                if ((d->kind == SCG_DIST_IMPULSE || d->kind == SCG_DIST_STEP) && d->step_offset < 0) {
                    int v = row ? (int)row[slot] : (philox ? rng_randint(key, SCG_STREAM_EPISODE, (uint32_t)k, (uint32_t)slot, (uint32_t)h->ctrl_steps) : 0);
                    T.st.dist_off[(size_t)(m * SCG_MAX_DIST + j) * T.N + n] = v; ++slot;
                }
            }
    }
    double inr[SCG_MAX_INERTIAL], init[NX];
    {
        double tmp[SCG_MAX_INERTIAL + NX];
        reset_draw<SYS, R>(T, h, n, k, tmp, 1, SCG_MAX_INERTIAL);
#pragma unroll
        for (int j = 0; j < SCG_MAX_INERTIAL; ++j) inr[j] = tmp[j];
#pragma unroll
        for (int j = 0; j < NX; ++j) init[j] = tmp[SCG_MAX_INERTIAL + j];
    }
#pragma unroll
    for (int j = 0; j < SCG_MAX_INERTIAL; ++j) e.inertial[j] = (R)inr[j];
    env_derived<SYS, R>(e);
    R initr[NX];
#pragma unroll
    for (int j = 0; j < NX; ++j) initr[j] = (R)init[j];
    e.episode = k + 1;
    Physics<SYS, R>::set_init(e, initr);
    Physics<SYS, R>::state(e, state);
}

// noise source of the reset-time observation (Philox stream RESET keyed by the episode; zero in table mode)
template <typename R> __device__ __forceinline__ NoiseSrc<R> reset_noise(const DevTables& T, const HotCfg<R>* h, int n, int episode_done) {
    NoiseSrc<R> ns; ns.row = nullptr; ns.philox = h->philox != 0;
    ns.key.k0 = h->key0; ns.key.k1 = h->key1; ns.key.genv = h->genv0 + (uint32_t)n;
    ns.stream = SCG_STREAM_RESET; ns.idx = (uint32_t)(episode_done - 1); ns.slot = 0;
    return ns;
}

struct StepInfo {
    int done, truncated, cviol, out_of_bounds, goal_reached, current_step;
};


// ---------------------------------------------------------------------------------------------
// One control step, split into phases so that the fused collector can spread them over warps:
//   env_pre (action map, action / dynamics disturbances, clip)  ->  env_physics  ->  { env_reward_mse | env_done | env_constraints }
//   ->  env_finalize (penalty, done_on_violation, time limit).   `env_step` chains them on one lane (gym-style kernels).
// ---------------------------------------------------------------------------------------------
template <int SYS, typename R> struct StepCtx {
    R noisy[SysDims<SYS>::NU];     // current_noisy_physical_action
    R clipped[SysDims<SYS>::NU];   // current_clipped_action
    R fd[3];                       // dynamics disturbance force
    bool has_dyn;
};

// quadrotor.py:722-747 / cartpole.py:479-502 + the dynamics disturbance draw (quadrotor.py:413-435, cartpole.py:540-551)
// EXT: the instantiations that also take the adversary's action (`adv`: clipped / scaled / offset, or nullptr = self.adv_action is None)
template <int SYS, typename R, bool EXT = false>
__device__ __forceinline__ void env_pre(const DevTables& T, const HotCfg<R>* h, int n, const EnvRegs<SYS, R>& e, const R* action, const double* noise_row,
                                        StepCtx<SYS, R>& cx, NoiseSrc<R>& ns, const R* adv = nullptr) {
    constexpr int NU = SysDims<SYS>::NU;
    const int pyb_step = e.ctrl_step * h->steps;
    ns.row = noise_row; ns.philox = h->philox != 0;
    ns.key.k0 = h->key0; ns.key.k1 = h->key1; ns.key.genv = h->genv0 + (uint32_t)n;
    ns.stream = SCG_STREAM_STEP; ns.idx = e.total_steps; ns.slot = 0;
#pragma unroll
    for (int k = 0; k < NU; ++k) {
        if (SYS == SCG_SYS_CARTPOLE) cx.noisy[k] = h->normalized_action ? h->hover * action[k] : action[k];                   // cartpole.py:527-528
        else cx.noisy[k] = h->normalized_action ? ((R)1 + h->act_scale * action[k]) * h->hover : action[k];                    // quadrotor.py:772-773
    }
    if (h->has_dist[SCG_MODE_ACTION])
        apply_disturbances<R>(T.cfg, T.st.dist_off, T.N, n, SCG_MODE_ACTION, e.ctrl_step, pyb_step, cx.noisy, NU, ns);
    if constexpr (EXT) {
        if (h->adv_mode == SCG_ADV_ACTION && adv) {                    // quadrotor.py:737-738, cartpole.py:494-495
#pragma unroll
            for (int k = 0; k < NU; ++k) cx.noisy[k] += adv[k];
        }
    }
#pragma unroll
    for (int k = 0; k < NU; ++k) {
        R t = cx.noisy[k];
        t = t < h->act_low[k] ? h->act_low[k] : t;
        t = t > h->act_high[k] ? h->act_high[k] : t;
        cx.clipped[k] = t;
    }
    cx.has_dyn = h->has_dist[SCG_MODE_DYNAMICS] != 0;
    cx.fd[0] = cx.fd[1] = cx.fd[2] = (R)0;
    if (cx.has_dyn) {
        const int ddim = (SYS == SCG_SYS_CARTPOLE) ? 2 : (SYS == SCG_SYS_QUAD1D) ? 1 : (SYS == SCG_SYS_QUAD2D) ? 2 : 3;
        apply_disturbances<R>(T.cfg, T.st.dist_off, T.N, n, SCG_MODE_DYNAMICS, e.ctrl_step, pyb_step, cx.fd, ddim, ns);
    }
    if constexpr (EXT) {
        if (h->adv_mode == SCG_ADV_DYNAMICS) {                         // quadrotor.py:416-425, cartpole.py:543-551 (same layout as the passive force)
            constexpr int ddim = (SYS == SCG_SYS_CARTPOLE) ? 2 : (SYS == SCG_SYS_QUAD1D) ? 1 : (SYS == SCG_SYS_QUAD2D) ? 2 : 3;
            cx.has_dyn = true;
            if (adv) {
#pragma unroll
                for (int k = 0; k < ddim; ++k) cx.fd[k] += adv[k];
            }
        }
    }
}

template <int SYS, typename R, bool EXT = false>
__device__ __forceinline__ void env_physics(const HotCfg<R>* h, EnvRegs<SYS, R>& e, const StepCtx<SYS, R>& cx, R* state) {
    if constexpr (SYS == SCG_SYS_CARTPOLE) {
        Physics<SYS, R>::advance(h, e, cx.clipped, cx.fd, cx.has_dyn);
    } else {
        R f[4]; quad_motor_forces<SYS, R>(h, cx.clipped, f);
        SCG_TK(2);
        Physics<SYS, R>::template advance<EXT>(h, e, f, cx.fd, cx.has_dyn);
    }
    Physics<SYS, R>::state(e, state);
}

// reward (quadrotor.py:819-862, cartpole.py:611-652) and info['mse'] (quadrotor.py:907-922, cartpole.py:685-695);
// ctrl_step = the counter BEFORE this step's increment
template <int SYS, typename R, typename GT>
__device__ __forceinline__ void env_reward_mse(const HotCfg<R>* h, const GT* xg, const R* state, const R* noisy, const R* clipped, int ctrl_step,
                                               R& rew_out, R& mse_out) {
    constexpr int NX = SysDims<SYS>::NX, NU = SysDims<SYS>::NU;
    R rew, mse = (R)0;
    const GT* g = goal_row<GT>(xg, h->L, h->task, NX, ctrl_step + 1);
    R gk[NX];
#pragma unroll
    for (int k = 0; k < NX; ++k) gk[k] = (R)g[k];
    if (h->cost == SCG_COST_RL_REWARD) {
        R dist = (R)0;
#pragma unroll
        for (int k = 0; k < NX; ++k) {
            R sk = state[k];
            if (SYS == SCG_SYS_CARTPOLE && k == 2) sk = normalize_angle<R>(sk);                                  // cartpole.py:620
            const R er = sk - gk[k];
            dist += h->rew_state_w[k] * er * er;
        }
#pragma unroll
        for (int k = 0; k < NU; ++k) {
            const R ae = (SYS == SCG_SYS_CARTPOLE) ? noisy[k] : noisy[k] - h->u_goal[k];                          // cartpole.py:625 / quadrotor.py:829
            dist += h->rew_act_w[k] * ae * ae;
        }
        rew = -dist;
        if (h->rew_exponential) rew = Mth<R>::exp_(rew);
    } else {
        const int idx = (SYS == SCG_SYS_CARTPOLE) ? ctrl_step : ctrl_step + 1;                                    // cartpole.py:648 / quadrotor.py:858
        const GT* gq = goal_row<GT>(xg, h->L, h->task, NX, idx);
        R l = (R)0;
#pragma unroll
        for (int k = 0; k < NX; ++k) { const R er = state[k] - (R)gq[k]; l += (R)0.5 * h->Qw[k] * er * er; }
#pragma unroll
        for (int k = 0; k < NU; ++k) { const R ae = clipped[k] - h->u_goal[k]; l += (R)0.5 * h->Rw[k] * ae * ae; }
        rew = -l;
    }
#pragma unroll
    for (int k = 0; k < NX; ++k) {
        R sk = state[k];
        if (h->task == SCG_TASK_TRAJ_TRACKING) {
            const bool is_angle = (SYS == SCG_SYS_CARTPOLE && k == 2) || (SYS == SCG_SYS_QUAD2D && k == 4) || (SYS == SCG_SYS_QUAD3D && k >= 6 && k < 9);
            if (is_angle) sk = normalize_angle<R>(sk);
        }
        const R er = (sk - gk[k]) * h->mse_w[k];
        mse += er * er;
    }
    rew_out = rew; mse_out = mse;
}

// np.linalg.norm(state - goal) < tol (quadrotor.py:868, cartpole.py:659) without a square root on the common path: the squares decide unless
// they are within rounding of each other, and only then the root is taken (out of line)
template <typename R> static __device__ SCG_COLD bool goal_norm_below_exact(R n2, R tol) { return Mth<R>::sqrt_(n2) < tol; }
template <typename R> __device__ __forceinline__ bool goal_norm_below(R n2, R tol) {
    const R t2 = tol * tol, eps = (R)(sizeof(R) == 8 ? 1e-12 : 1e-5);
    if (tol <= (R)0 || n2 > t2 * ((R)1 + eps)) return false;
    if (n2 < t2 * ((R)1 - eps)) return true;
    return goal_norm_below_exact<R>(n2, tol);
}
// done before the constraint / time-limit logic (quadrotor.py:864-894, cartpole.py:654-672)
template <int SYS, typename R, typename GT>
__device__ __forceinline__ int env_done(const HotCfg<R>* h, const GT* xg, const R* state, int& goal_reached, int& oob) {
    constexpr int NX = SysDims<SYS>::NX;
    int done = 0; goal_reached = 0; oob = 0;
    if (h->task == SCG_TASK_STABILIZATION) {
        R n2 = (R)0;
#pragma unroll
        for (int k = 0; k < NX; ++k) { const R er = state[k] - (R)xg[k]; n2 += er * er; }
        goal_reached = goal_norm_below<R>(n2, h->goal_tol);
        done = goal_reached;
    }
    if (!done && h->done_on_oob) {
        if (SYS == SCG_SYS_CARTPOLE) {
            const R x = state[0], th = state[2];
            oob = (x < -h->x_thr) || (x > h->x_thr) || (th < -h->th_thr) || (th > h->th_thr);
        } else {
#pragma unroll
            for (int k = 0; k < NX; ++k) {
                const bool masked = (SYS == SCG_SYS_QUAD3D) ? ((k % 2 == 0 && k < 6) || (k >= 6 && k < 9)) : (k % 2 == 0);  // quadrotor.py:878-883
                if (masked) oob |= ((double)state[k] < h->state_low[k]) || ((double)state[k] > h->state_high[k]);
            }
        }
        done = oob;
    }
    return done;
}

// constraint_violation (benchmark_env.py:468-485) for z = [state ; current_noisy_physical_action] already staged in `zbuf`
// (constraints.py:166-171); `zbuf` is this env's column of a shared scratch [NX+NU][zstride] of doubles.
static __device__ __noinline__ int env_constraints(const DevTables& T, const ConRow* __restrict__ rows, int n_con, int con_fast,
                                               const double* __restrict__ zbuf, int zstride, double* cvals, float* cvals_f32) {
    int cviol = 0;
    if (con_fast) {
        // violated <=> np.round(g, 8) > 0 (>= 0 if strict) <=> g*1e8 > 0.5 (>= -0.5): rint() rounds half to even
        // (constraints.py:109,126-129); no stores inside this loop so the rows pipeline
#pragma unroll 4
        for (int r = 0; r < n_con; ++r) {
            const ConRow row = rows[r];
            double acc = __dmul_rn(row.val, zbuf[row.col * zstride]);
            if (row.flags & 2) acc = fabs(acc);
            const double t8 = __dsub_rn(acc, row.b) * 1e8;
            cviol |= (row.flags & 1) ? (t8 >= -0.5) : (t8 > 0.5);
        }
        if (cvals || cvals_f32) {
            for (int r = 0; r < n_con; ++r) {
                const ConRow row = rows[r];
                double acc = __dmul_rn(row.val, zbuf[row.col * zstride]);
                if (row.flags & 2) acc = fabs(acc);
                const double g8 = rint(__dsub_rn(acc, row.b) * 1e8) / 1e8;   // np.round(g, 8)
                if (cvals) cvals[r] = g8;
                if (cvals_f32) cvals_f32[r] = (float)g8;
            }
        }
    } else {
        for (int r = 0; r < n_con; ++r) {
            double acc = 0;
            if (T.con_kind[r] == SCG_CON_QUAD) {       // x' F' P F x (constraints.py:222): every non-zero carries a pair of z indices
                for (int p = T.con_ptr[r]; p < T.con_ptr[r + 1]; ++p) { const int cc = T.con_col[p]; acc += T.con_val[p] * zbuf[(cc & 0xff) * zstride] * zbuf[(cc >> 8) * zstride]; }
            } else {
                for (int p = T.con_ptr[r]; p < T.con_ptr[r + 1]; ++p) acc += T.con_val[p] * zbuf[T.con_col[p] * zstride];
            }
            if (T.con_kind[r] == SCG_CON_ABS) acc = fabs(acc);
            const double gi = rint((acc - T.con_b[r]) * 1e8);
            cviol |= T.con_strict[r] ? (gi >= 0.0) : (gi > 0.0);
            if (cvals) cvals[r] = gi / 1e8;
            if (cvals_f32) cvals_f32[r] = (float)(gi / 1e8);
        }
    }
    return cviol;
}

// tail of after_step (benchmark_env.py:476-502); ctrl_step_after = the counter AFTER the increment
template <typename R>
__device__ __forceinline__ void env_finalize(const HotCfg<R>* h, int ctrl_step_after, int cviol, R& rew, int& done, int& truncated) {
    if (h->n_con > 0) {
        if (cviol && h->done_on_violation) { done = 1; if (h->cost == SCG_COST_RL_REWARD && h->use_penalty) rew = (R)0; }
        if (h->cost == SCG_COST_RL_REWARD && h->use_penalty && cviol) {
            if (h->rew_exponential) rew = Mth<R>::exp_(Mth<R>::log_(rew) - h->penalty);
            else rew -= h->penalty;
        }
    }
    truncated = 0;
    if (ctrl_step_after >= h->ctrl_steps) { truncated = !done; done = 1; }
}

// The whole step on one lane, WITHOUT auto-reset and WITHOUT building the observation (call make_obs next; `ns` is positioned
// for the observation noise).  action[NU] is the raw controller action; on return `state` holds env.state.
template <int SYS, typename R, typename GT>
__device__ __forceinline__ void env_step(const DevTables& T, const HotCfg<R>* h, const GT* xg, int n, EnvRegs<SYS, R>& e, const R* action, const double* noise_row,
                                         double* __restrict__ zbuf, int zstride, R* state, R& rew_out, R& mse_out, StepInfo& info,
                                         double* cvals, float* cvals_f32, NoiseSrc<R>& ns, const R* adv = nullptr) {
    constexpr int NX = SysDims<SYS>::NX, NU = SysDims<SYS>::NU;
    StepCtx<SYS, R> cx;
    env_pre<SYS, R, true>(T, h, n, e, action, noise_row, cx, ns, adv);      // gym-style API: always the full feature set
    env_physics<SYS, R, true>(h, e, cx, state);
    R rew, mse;
    env_reward_mse<SYS, R, GT>(h, xg, state, cx.noisy, cx.clipped, e.ctrl_step, rew, mse);
    int goal_reached, oob;
    int done = env_done<SYS, R, GT>(h, xg, state, goal_reached, oob);
    e.ctrl_step += 1; e.total_steps += 1;
    int cviol = 0;
    if (h->n_con > 0) {
#pragma unroll
        for (int k = 0; k < NX; ++k) zbuf[k * zstride] = (double)state[k];
#pragma unroll
        for (int k = 0; k < NU; ++k) zbuf[(NX + k) * zstride] = (double)cx.noisy[k];
        cviol = env_constraints(T, h->rows, h->n_con, h->con_fast, zbuf, zstride, cvals, cvals_f32);
    }
    int truncated;
    env_finalize<R>(h, e.ctrl_step, cviol, rew, done, truncated);
    rew_out = rew; mse_out = mse;
    info.done = done; info.truncated = truncated; info.cviol = cviol; info.out_of_bounds = oob; info.goal_reached = goal_reached; info.current_step = e.ctrl_step;
}
