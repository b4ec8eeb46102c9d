// scg_api.cu -- C-ABI of the rollout engine (see include/scg.h for the contract and the reference
// interfaces each entry point replaces).  Host side only: validation, device allocation, table upload,
// template dispatch and launches.  No CPU compute path exists here: without a CUDA device every entry
// point that would compute fails with SCG_ECUDA.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "scg.h"
#include "scg_kernels.cuh"
#include "scg_rollout_launch.cuh"

struct scg_env {
    scg_config cfg;
    int nx, nu, obs_dim, n_inertial, n_off_slots, P, D;
    int adv_dim = 0;             // dim of the adversary action (0 = no adversary)
    int device;
    // device buffers
    scg_config* d_cfg = nullptr;
    double* d_xgoal = nullptr; int L = 0;
    int n_con = 0; int32_t *d_con_kind = nullptr, *d_con_strict = nullptr, *d_con_ptr = nullptr, *d_con_col = nullptr;
    double *d_con_val = nullptr, *d_con_b = nullptr;
    double* d_ep_table = nullptr; int K = 0;
    double* d_noise_table = nullptr; int Tn = 0;
    void* d_state = nullptr; size_t state_bytes = 0;
    DevTables tab;
    long long noise_cursor = 0;
    bool did_reset = false;
    bool con_single = true;      // every constraint row has exactly one non-zero (bounds): the in-kernel fast path
    uint64_t launches = 0;
    std::string err;
};

static thread_local std::string g_create_err;

static int sys_nx(int s) { return s == SCG_SYS_CARTPOLE ? 4 : s == SCG_SYS_QUAD1D ? 2 : s == SCG_SYS_QUAD2D ? 6 : 12; }
static int sys_nu(int s) { return s == SCG_SYS_CARTPOLE ? 1 : s == SCG_SYS_QUAD1D ? 1 : s == SCG_SYS_QUAD2D ? 2 : 4; }

void scg_internal_set_error(const std::string& msg) { g_create_err = msg; }   // handle-free entry points (scg_ppo.cu)

static int fail(scg_env* e, int code, const std::string& msg) {
    if (e) e->err = msg; else g_create_err = msg;
    return code;
}
#define CUDA_TRY(env, call)                                                                                         \
    do {                                                                                                            \
        cudaError_t _st = (call);                                                                                   \
        if (_st != cudaSuccess) return fail(env, SCG_ECUDA, std::string(#call) + ": " + cudaGetErrorString(_st));  \
    } while (0)

// layout of the state blob: doubles first, then 4-byte sections
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static void carve_state(scg_env* e) {
    const size_t N = (size_t)e->cfg.num_envs;
    size_t off = 0;
    char* base = (char*)e->d_state;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    size_t o_phys = take(sizeof(double) * SCG_PS_MAX * N);
    size_t o_in = take(sizeof(double) * SCG_MAX_INERTIAL * N);
    size_t o_obs = take(sizeof(float) * N * e->obs_dim);
    size_t o_cs = take(sizeof(int32_t) * N);
    size_t o_ep = take(sizeof(int32_t) * N);
    size_t o_ts = take(sizeof(uint32_t) * N);
    size_t o_ir = take(sizeof(int32_t) * N);
    size_t o_do = take(sizeof(int32_t) * 3 * SCG_MAX_DIST * N);
    size_t o_acc = take(sizeof(float) * 3 * N);
    size_t o_adv = take(sizeof(float) * SCG_MAX_ADV * N);
    size_t o_as = take(sizeof(int32_t) * N);
    e->state_bytes = off;
    if (!base) return;
    e->tab.st.phys = (double*)(base + o_phys);
    e->tab.st.inertial = (double*)(base + o_in);
    e->tab.st.cur_obs = (float*)(base + o_obs);
    e->tab.st.ctrl_step = (int32_t*)(base + o_cs);
    e->tab.st.episode = (int32_t*)(base + o_ep);
    e->tab.st.total_steps = (uint32_t*)(base + o_ts);
    e->tab.st.initial_reset = (int32_t*)(base + o_ir);
    e->tab.st.dist_off = (int32_t*)(base + o_do);
    e->tab.st.ep_acc = (float*)(base + o_acc);
    e->tab.st.adv = (float*)(base + o_adv);
    e->tab.st.adv_set = (int32_t*)(base + o_as);
}

static int obs_dim_of(const scg_config* c) {
    const int nx = sys_nx(c->system);
    if (c->cost == SCG_COST_RL_REWARD && c->obs_goal_horizon > 0) {   // benchmark_env.py:432-443
        if (c->task == SCG_TASK_TRAJ_TRACKING) return nx * (1 + c->obs_goal_horizon);
        return nx * 2;
    }
    return nx;
}

extern "C" {

int scg_abi_version(void) { return SCG_ABI_VERSION; }

const char* scg_last_error(const scg_env* env) { return env ? env->err.c_str() : g_create_err.c_str(); }

int scg_create(const scg_config* cfg, scg_env** out) {
    if (!cfg || !out) return fail(nullptr, SCG_EINVAL, "scg_create: null argument");
    *out = nullptr;
    if (cfg->struct_size != (int32_t)sizeof(scg_config) || cfg->abi_version != SCG_ABI_VERSION)
        return fail(nullptr, SCG_EINVAL, "scg_create: scg_config size / ABI version mismatch");
    if (cfg->system < 0 || cfg->system > 3) return fail(nullptr, SCG_EINVAL, "scg_create: unknown system");
    if (cfg->num_envs <= 0) return fail(nullptr, SCG_EINVAL, "scg_create: num_envs must be positive");
    if (cfg->ctrl_freq <= 0 || cfg->pyb_freq <= 0 || cfg->pyb_freq % cfg->ctrl_freq != 0)   // benchmark_env.py:141-142
        return fail(nullptr, SCG_EINVAL, "[ERROR] in BenchmarkEnv.__init__(), pyb_freq is not divisible by env_freq.");
    if (cfg->precision != SCG_PREC_F64 && cfg->precision != SCG_PREC_F32) return fail(nullptr, SCG_EINVAL, "scg_create: bad precision");
    for (int m = 0; m < 3; ++m) {
        if (cfg->n_dist[m] < 0 || cfg->n_dist[m] > SCG_MAX_DIST) return fail(nullptr, SCG_EINVAL, "scg_create: too many disturbances in one mode");
        for (int j = 0; j < cfg->n_dist[m]; ++j)
            if (cfg->dist[m][j].dim < 0 || cfg->dist[m][j].dim > SCG_MAX_DIST_DIM) return fail(nullptr, SCG_EINVAL, "scg_create: disturbance dim out of range");
    }
    if (obs_dim_of(cfg) > SCG_MAX_OBS) return fail(nullptr, SCG_EINVAL, "scg_create: observation too wide (obs_goal_horizon too large)");
    if (cfg->physics < SCG_PHYS_PYB || cfg->physics > SCG_PHYS_PYB_GND_DRAG_DW) return fail(nullptr, SCG_EINVAL, "scg_create: unknown physics");
    if (cfg->physics != SCG_PHYS_PYB && cfg->system == SCG_SYS_CARTPOLE) return fail(nullptr, SCG_EINVAL, "scg_create: physics variants are quadrotor-only");
    if ((cfg->physics == SCG_PHYS_PYB_GND || cfg->physics == SCG_PHYS_PYB_GND_DRAG_DW) && !(cfg->gnd_eff_h_clip > 0 && cfg->prop_radius > 0))
        return fail(nullptr, SCG_EINVAL, "scg_create: ground effect needs prop_radius and gnd_eff_h_clip > 0");
    if (cfg->adversary_mode < SCG_ADV_NONE || cfg->adversary_mode > SCG_ADV_DYNAMICS) return fail(nullptr, SCG_EINVAL, "scg_create: unknown adversary_mode");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(nullptr, SCG_ECUDA, "scg_create: no CUDA device available (this engine has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, SCG_EINVAL, "scg_create: bad device ordinal");
    scg_env* e = new (std::nothrow) scg_env();
    if (!e) return fail(nullptr, SCG_ENOMEM, "scg_create: out of host memory");
    e->cfg = *cfg;
    e->device = cfg->device;
    e->nx = sys_nx(cfg->system); e->nu = sys_nu(cfg->system); e->obs_dim = obs_dim_of(cfg);
    e->n_inertial = cfg->system == SCG_SYS_CARTPOLE ? 3 : 4;
    e->adv_dim = cfg->adversary_mode == SCG_ADV_ACTION ? e->nu                      // benchmark_env.py:289-295: dim of the driven disturbance mode
                 : cfg->adversary_mode == SCG_ADV_DYNAMICS ? (cfg->system == SCG_SYS_CARTPOLE ? 2 : cfg->system == SCG_SYS_QUAD1D ? 1 : cfg->system == SCG_SYS_QUAD2D ? 2 : 3) : 0;
    e->n_off_slots = 0; e->D = 0;
    for (int m = 0; m < 3; ++m)
        for (int j = 0; j < cfg->n_dist[m]; ++j) {
            const int k = cfg->dist[m][j].kind;
            if ((k == SCG_DIST_IMPULSE || k == SCG_DIST_STEP) && cfg->dist[m][j].step_offset < 0) e->n_off_slots += 1;
            if (k == SCG_DIST_UNIFORM || k == SCG_DIST_WHITE || k == SCG_DIST_PERIODIC) e->D += cfg->dist[m][j].dim;
        }
    e->P = e->n_off_slots + SCG_MAX_INERTIAL + e->nx;
    memset(&e->tab, 0, sizeof(e->tab));
    cudaError_t st = cudaSetDevice(e->device);
    if (st == cudaSuccess) st = cudaMalloc((void**)&e->d_cfg, sizeof(scg_config));
    if (st == cudaSuccess) st = cudaMemcpy(e->d_cfg, cfg, sizeof(scg_config), cudaMemcpyHostToDevice);
    carve_state(e);   // sizes only
    if (st == cudaSuccess) st = cudaMalloc(&e->d_state, e->state_bytes);
    if (st == cudaSuccess) st = cudaMemset(e->d_state, 0, e->state_bytes);
    if (st != cudaSuccess) {
        std::string msg = std::string("scg_create: ") + cudaGetErrorString(st);
        cudaFree(e->d_cfg); cudaFree(e->d_state); delete e;
        return fail(nullptr, SCG_ECUDA, msg);
    }
    carve_state(e);
    e->tab.cfg = e->d_cfg; e->tab.N = cfg->num_envs; e->tab.obs_dim = e->obs_dim; e->tab.P = e->P; e->tab.D = e->D; e->tab.n_off_slots = e->n_off_slots;
    *out = e;
    return SCG_OK;
}

void scg_destroy(scg_env* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaFree(e->d_cfg); cudaFree(e->d_xgoal); cudaFree(e->d_con_kind); cudaFree(e->d_con_strict); cudaFree(e->d_con_ptr); cudaFree(e->d_con_col);
    cudaFree(e->d_con_val); cudaFree(e->d_con_b); cudaFree(e->d_ep_table); cudaFree(e->d_noise_table); cudaFree(e->d_state);
    delete e;
}

int scg_dims(const scg_env* e, int32_t* nx, int32_t* nu, int32_t* obs_dim, int32_t* n_con, int32_t* pw, int32_t* dw) {
    if (!e) return SCG_EINVAL;
    if (nx) *nx = e->nx; if (nu) *nu = e->nu; if (obs_dim) *obs_dim = e->obs_dim; if (n_con) *n_con = e->n_con;
    if (pw) *pw = e->P; if (dw) *dw = e->D;
    return SCG_OK;
}

extern "C++" {
template <typename Tp> static cudaError_t upload(Tp** dst, const Tp* src, size_t n) {
    cudaFree(*dst); *dst = nullptr;
    if (n == 0) return cudaSuccess;
    cudaError_t st = cudaMalloc((void**)dst, sizeof(Tp) * n);
    if (st != cudaSuccess) return st;
    return cudaMemcpy(*dst, src, sizeof(Tp) * n, cudaMemcpyHostToDevice);
}
}  // extern C++

int scg_set_reference(scg_env* e, const double* x_goal, int32_t L) {
    if (!e || !x_goal || L <= 0) return fail(e, SCG_EINVAL, "scg_set_reference: bad argument");
    CUDA_TRY(e, cudaSetDevice(e->device));
    CUDA_TRY(e, cudaDeviceSynchronize());
    CUDA_TRY(e, upload(&e->d_xgoal, x_goal, (size_t)L * e->nx));
    e->L = L; e->tab.x_goal = e->d_xgoal; e->tab.L = L;
    return SCG_OK;
}

int scg_set_constraints(scg_env* e, int32_t n_rows, const int32_t* kind, const int32_t* strict, const int32_t* row_ptr, const int32_t* col,
                        const double* val, const double* b) {
    if (!e || n_rows < 0 || n_rows > SCG_MAX_CON_ROWS) return fail(e, SCG_EINVAL, "scg_set_constraints: bad row count");
    CUDA_TRY(e, cudaSetDevice(e->device));
    CUDA_TRY(e, cudaDeviceSynchronize());
    if (n_rows == 0) { e->n_con = 0; e->tab.n_con = 0; return SCG_OK; }
    if (!kind || !strict || !row_ptr || !col || !val || !b) return fail(e, SCG_EINVAL, "scg_set_constraints: null argument");
    const int nnz = row_ptr[n_rows];
    if (nnz < 0 || nnz > SCG_MAX_CON_NNZ) return fail(e, SCG_EINVAL, "scg_set_constraints: too many non-zeros");
    for (int r = 0; r < n_rows; ++r) {
        if (kind[r] < SCG_CON_LINEAR || kind[r] > SCG_CON_QUAD) return fail(e, SCG_EINVAL, "scg_set_constraints: unknown row kind");
        for (int i = row_ptr[r]; i < row_ptr[r + 1]; ++i) {
            const int c0 = kind[r] == SCG_CON_QUAD ? (col[i] & 0xff) : col[i], c1 = kind[r] == SCG_CON_QUAD ? (col[i] >> 8) : 0;
            if (c0 < 0 || c0 >= e->nx + e->nu || c1 < 0 || c1 >= e->nx + e->nu) return fail(e, SCG_EINVAL, "scg_set_constraints: column index out of range");
        }
    }
    CUDA_TRY(e, upload(&e->d_con_kind, kind, (size_t)n_rows));
    CUDA_TRY(e, upload(&e->d_con_strict, strict, (size_t)n_rows));
    CUDA_TRY(e, upload(&e->d_con_ptr, row_ptr, (size_t)n_rows + 1));
    CUDA_TRY(e, upload(&e->d_con_col, col, (size_t)(nnz > 0 ? nnz : 1)));
    CUDA_TRY(e, upload(&e->d_con_val, val, (size_t)(nnz > 0 ? nnz : 1)));
    CUDA_TRY(e, upload(&e->d_con_b, b, (size_t)n_rows));
    e->n_con = n_rows;
    e->con_single = true;
    for (int r = 0; r < n_rows; ++r) if (row_ptr[r + 1] - row_ptr[r] != 1 || kind[r] == SCG_CON_QUAD) e->con_single = false;
    e->tab.n_con = n_rows; e->tab.con_kind = e->d_con_kind; e->tab.con_strict = e->d_con_strict; e->tab.con_ptr = e->d_con_ptr;
    e->tab.con_col = e->d_con_col; e->tab.con_val = e->d_con_val; e->tab.con_b = e->d_con_b;
    return SCG_OK;
}

int scg_set_episode_table(scg_env* e, const double* tab, int32_t K) {
    if (!e || !tab || K <= 0) return fail(e, SCG_EINVAL, "scg_set_episode_table: bad argument");
    CUDA_TRY(e, cudaSetDevice(e->device));
    CUDA_TRY(e, cudaDeviceSynchronize());
    CUDA_TRY(e, upload(&e->d_ep_table, tab, (size_t)K * e->cfg.num_envs * e->P));
    e->K = K; e->tab.ep_table = e->d_ep_table; e->tab.K = K;
    return SCG_OK;
}

int scg_set_noise_table(scg_env* e, const double* tab, int32_t Tn) {
    if (!e || !tab || Tn <= 0) return fail(e, SCG_EINVAL, "scg_set_noise_table: bad argument");
    if (e->D == 0) return fail(e, SCG_EINVAL, "scg_set_noise_table: this configuration draws no per-step noise");
    CUDA_TRY(e, cudaSetDevice(e->device));
    CUDA_TRY(e, cudaDeviceSynchronize());
    CUDA_TRY(e, upload(&e->d_noise_table, tab, (size_t)Tn * e->cfg.num_envs * e->D));
    e->Tn = Tn; e->tab.noise_table = e->d_noise_table; e->tab.Tn = Tn; e->noise_cursor = 0;
    return SCG_OK;
}

static int check_ready(scg_env* e, const char* who) {
    if (!e) return SCG_EINVAL;
    if (!e->d_xgoal) return fail(e, SCG_EINVAL, std::string(who) + ": scg_set_reference() was never called (X_GOAL missing)");
    return SCG_OK;
}

#define DISPATCH_SYS_PREC(env, FN)                                                                 \
    do {                                                                                           \
        const int _s = (env)->cfg.system; const bool _d = (env)->cfg.precision == SCG_PREC_F64;    \
        if (_s == SCG_SYS_CARTPOLE) { if (_d) FN(SCG_SYS_CARTPOLE, double); else FN(SCG_SYS_CARTPOLE, float); } \
        else if (_s == SCG_SYS_QUAD1D) { if (_d) FN(SCG_SYS_QUAD1D, double); else FN(SCG_SYS_QUAD1D, float); }  \
        else if (_s == SCG_SYS_QUAD2D) { if (_d) FN(SCG_SYS_QUAD2D, double); else FN(SCG_SYS_QUAD2D, float); }  \
        else { if (_d) FN(SCG_SYS_QUAD3D, double); else FN(SCG_SYS_QUAD3D, float); }               \
    } while (0)

int scg_reset(scg_env* e, const uint8_t* d_mask, float* d_obs, double* d_obs_f64, double* d_state_f64, void* stream) {
    int rc = check_ready(e, "scg_reset"); if (rc) return rc;
    CUDA_TRY(e, cudaSetDevice(e->device));
    ResetArgs A{d_mask, d_obs, d_obs_f64, d_state_f64};
    const int N = e->cfg.num_envs, B = 128, G = (N + B - 1) / B;
    cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH_RESET(SYS, R) reset_kernel<SYS, R><<<G, B, 0, s>>>(e->tab, A)
    DISPATCH_SYS_PREC(e, LAUNCH_RESET);
#undef LAUNCH_RESET
    CUDA_TRY(e, cudaGetLastError());
    e->launches += 1; e->did_reset = true;
    return SCG_OK;
}

int scg_step(scg_env* e, const float* d_act, const double* d_act_f64, const scg_step_out* out, void* stream) {
    int rc = check_ready(e, "scg_step"); if (rc) return rc;
    if (!out || (!d_act && !d_act_f64)) return fail(e, SCG_EINVAL, "scg_step: null argument");
    if (!e->did_reset)   // benchmark_env.py:230-235
        return fail(e, SCG_ESTATE, "[ERROR] You must call env.reset() at least once before using env.step().");
    CUDA_TRY(e, cudaSetDevice(e->device));
    StepArgs A{d_act, d_act_f64, *out, e->noise_cursor};
    const int N = e->cfg.num_envs, B = 128, G = (N + B - 1) / B;
    cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH_STEP(SYS, R) step_kernel<SYS, R><<<G, B, 0, s>>>(e->tab, A)
    DISPATCH_SYS_PREC(e, LAUNCH_STEP);
#undef LAUNCH_STEP
    CUDA_TRY(e, cudaGetLastError());
    e->launches += 1; e->noise_cursor += 1;
    return SCG_OK;
}

int scg_adversary_dim(const scg_env* e) { return e ? e->adv_dim : 0; }

int scg_set_adversary_control(scg_env* e, const float* d_adv, void* stream) {
    if (!e || !d_adv) return fail(e, SCG_EINVAL, "scg_set_adversary_control: null argument");
    if (e->adv_dim == 0)   // benchmark_env.py:225-228
        return fail(e, SCG_ESTATE, "[ERROR] adversary_disturbance does not exist, env.set_adversary_control() cannot be called.");
    CUDA_TRY(e, cudaSetDevice(e->device));
    const int N = e->cfg.num_envs, B = 128, G = (N + B - 1) / B;
    adversary_set_kernel<<<G, B, 0, (cudaStream_t)stream>>>(e->tab, d_adv, e->adv_dim);
    CUDA_TRY(e, cudaGetLastError());
    e->launches += 1;
    return SCG_OK;
}

int scg_rollout(scg_env* e, const scg_rollout_args* a, const scg_rollout_out* out, void* stream) {
    int rc = check_ready(e, "scg_rollout"); if (rc) return rc;
    if (!a || !out) return fail(e, SCG_EINVAL, "scg_rollout: null argument");
    if (!e->did_reset) return fail(e, SCG_ESTATE, "[ERROR] You must call env.reset() at least once before using env.step().");
    if (a->obs_dim != e->obs_dim || a->act_dim != e->nu) return fail(e, SCG_EINVAL, "scg_rollout: policy dims do not match the env");
    if (a->T <= 0) return fail(e, SCG_EINVAL, "scg_rollout: T must be positive");
    if (!a->d_policy) return fail(e, SCG_EINVAL, "scg_rollout: null policy blob");
    if (((uintptr_t)a->d_policy & 15) != 0) return fail(e, SCG_EINVAL, "scg_rollout: policy blob must be 16-byte aligned");
    if (!out->d_obs || !out->d_act || !out->d_rew || !out->d_mask || !out->d_v || !out->d_logp || !out->d_terminal_v)
        return fail(e, SCG_EINVAL, "scg_rollout: obs/act/rew/mask/v/logp/terminal_v buffers are mandatory");
    if (a->activation < 0 || a->activation > 2) return fail(e, SCG_EINVAL, "scg_rollout: unknown activation");
    if (!a->d_eps && !a->deterministic && e->cfg.rng_mode != SCG_RNG_PHILOX)
        return fail(e, SCG_EINVAL, "scg_rollout: d_eps is required unless rng_mode is SCG_RNG_PHILOX or deterministic is set");
    CUDA_TRY(e, cudaSetDevice(e->device));
    RolloutParams P;
    P.T = a->T; P.activation = a->activation; P.use_gae = a->use_gae; P.deterministic = a->deterministic;
    P.gamma = (float)a->gamma; P.lam = (float)a->gae_lambda; P.gamma_d = a->gamma; P.lam_d = a->gae_lambda;
    P.policy = a->d_policy; P.eps = a->d_eps; P.noise_cursor = e->noise_cursor; P.out = *out; P.prof = nullptr;
    if (a->n_peers < 0 || a->n_peers > SCG_MAX_PEERS) return fail(e, SCG_EINVAL, "scg_rollout: n_peers out of range");
    P.n_peers = a->n_peers;
    for (int g = 0; g < SCG_MAX_PEERS; ++g) {
        P.peer_off[g] = g < a->n_peers ? a->peer_byte_offset[g] : 0;
        if (P.peer_off[g] & 15) return fail(e, SCG_EINVAL, "scg_rollout: peer_byte_offset must be a multiple of 16");
    }
    if (a->multicast_byte_offset & 15) return fail(e, SCG_EINVAL, "scg_rollout: multicast_byte_offset must be a multiple of 16");
    P.mc_off = a->multicast_byte_offset;
    P.norm = a->d_norm; P.obs0 = a->d_obs0;
    if (a->safety_constraints < 0 || a->safety_constraints > SCG_MAX_SAFETY || (a->safety_constraints > 0 && (!a->d_safety || a->safety_hidden < 1 || a->safety_hidden > 256)))
        return fail(e, SCG_EINVAL, "scg_rollout: bad safety layer (safety_constraints in [0, 16], safety_hidden in [1, 256], d_safety non-null)");
    if (a->safety_constraints > e->n_con) return fail(e, SCG_EINVAL, "scg_rollout: safety_constraints exceeds the number of constraint rows of the env");
    if (a->safety_constraints > 0 && !e->con_single) return fail(e, SCG_EINVAL, "scg_rollout: the safety layer needs single-variable (bound) constraint rows");
    P.safety_C = a->safety_constraints; P.safety_H = a->safety_hidden; P.safety = a->d_safety;
    for (int i = 0; i < SCG_MAX_SAFETY; ++i) P.safety_slack[i] = (float)a->safety_slack[i];
    if (a->d_adv_policy) {
        if (e->adv_dim == 0) return fail(e, SCG_ESTATE, "[ERROR] adversary_disturbance does not exist, env.set_adversary_control() cannot be called.");
        if (a->adv_act_dim != e->adv_dim) return fail(e, SCG_EINVAL, "scg_rollout: adv_act_dim does not match the env's adversary action space");
        if (((uintptr_t)a->d_adv_policy & 15) != 0) return fail(e, SCG_EINVAL, "scg_rollout: adversary blob must be 16-byte aligned");
        if (!a->d_eps_adv && !a->deterministic && e->cfg.rng_mode != SCG_RNG_PHILOX)
            return fail(e, SCG_EINVAL, "scg_rollout: d_eps_adv is required unless rng_mode is SCG_RNG_PHILOX or deterministic is set");
    }
    P.adv_policy = a->d_adv_policy; P.eps_adv = a->d_eps_adv; P.adv_dim = a->d_adv_policy ? a->adv_act_dim : 0; P.adv_record = a->d_adv_policy ? a->adv_record : 0;
    cudaStream_t s = (cudaStream_t)stream;
    const bool dbl = e->cfg.precision == SCG_PREC_F64;
    const bool ext_phys = e->cfg.physics == SCG_PHYS_PYB_GND || e->cfg.physics == SCG_PHYS_PYB_DRAG || e->cfg.physics == SCG_PHYS_PYB_GND_DRAG_DW;
    // rollout_kernel's FAST-schedule test, evaluated here: use_cand (no randomised disturbance offsets), no observation disturbance, no
    // done_on_violation, no constraint penalty
    const bool fast_ok = e->n_off_slots == 0 && e->cfg.n_dist[SCG_MODE_OBSERVATION] == 0 && !e->cfg.done_on_violation && !e->cfg.use_constraint_penalty;
    RolloutLaunch rl{e->cfg.num_envs, e->obs_dim, e->nu, e->nx, e->L, e->tab, (ext_phys || e->adv_dim > 0) ? 1 : 0, fast_ok ? 1 : 0};
    RolloutLaunchResult r;
    switch (e->cfg.system) {
        case SCG_SYS_CARTPOLE: r = dbl ? scg_rollout_cartpole_f64(rl, a, P, s) : scg_rollout_cartpole_f32(rl, a, P, s); break;
        case SCG_SYS_QUAD1D: r = dbl ? scg_rollout_quad1d_f64(rl, a, P, s) : scg_rollout_quad1d_f32(rl, a, P, s); break;
        case SCG_SYS_QUAD2D: r = dbl ? scg_rollout_quad2d_f64(rl, a, P, s) : scg_rollout_quad2d_f32(rl, a, P, s); break;
        case SCG_SYS_QUAD3D: r = dbl ? scg_rollout_quad3d_f64(rl, a, P, s) : scg_rollout_quad3d_f32(rl, a, P, s); break;
        default: return fail(e, SCG_EINVAL, "scg_rollout: unknown system");
    }
    if (r.code == SCG_ECUDA) return fail(e, SCG_ECUDA, std::string(r.msg) + ": " + cudaGetErrorString(r.cuda));
    if (r.code != SCG_OK) return fail(e, r.code, r.msg);
    e->launches += 1; e->noise_cursor += a->T;
    return SCG_OK;
}

int scg_policy_blob_floats(int32_t obs_dim, int32_t act_dim, int32_t hidden_dim) {
    if (obs_dim <= 0 || act_dim <= 0 || hidden_dim <= 0) return SCG_EINVAL;
    return blob_layout(obs_dim, act_dim, hidden_dim).total;
}

int scg_policy_pack(int32_t obs_dim, int32_t act_dim, int32_t hid, const float* const aw[3], const float* const ab[3], const float* logstd,
                    const float* const cw[3], const float* const cb[3], float* blob) {
    if (!aw || !ab || !cw || !cb || !logstd || !blob) return SCG_EINVAL;
    const BlobLayout L = blob_layout(obs_dim, act_dim, hid);
    memset(blob, 0, sizeof(float) * L.total);
    if ((obs_dim & 1) || (hid & 3)) return SCG_EINVAL;   // pair-major layout: even input width, hidden multiple of 4
    for (int k = 0; k < obs_dim; ++k)
        for (int u = 0; u < hid; ++u) {   // W1P [k/2][2H][2]; nn.Linear weight is [out][in]
            blob[L.w1t + ((k >> 1) * 2 * hid + u) * 2 + (k & 1)] = aw[0][(size_t)u * obs_dim + k];
            blob[L.w1t + ((k >> 1) * 2 * hid + hid + u) * 2 + (k & 1)] = cw[0][(size_t)u * obs_dim + k];
        }
    for (int u = 0; u < hid; ++u) { blob[L.b1 + u] = ab[0][u]; blob[L.b1 + hid + u] = cb[0][u]; }
    for (int k = 0; k < hid; ++k)
        for (int u = 0; u < hid; ++u) {   // W2P [net][k/2][H][2]
            blob[L.w2t + ((k >> 1) * hid + u) * 2 + (k & 1)] = aw[1][(size_t)u * hid + k];
            blob[L.w2t + hid * hid + ((k >> 1) * hid + u) * 2 + (k & 1)] = cw[1][(size_t)u * hid + k];
        }
    for (int u = 0; u < hid; ++u) { blob[L.b2 + u] = ab[1][u]; blob[L.b2 + hid + u] = cb[1][u]; }
    for (int o = 0; o < act_dim; ++o)
        for (int k = 0; k < hid; ++k) blob[L.w3 + o * hid + k] = aw[2][(size_t)o * hid + k];
    for (int k = 0; k < hid; ++k) blob[L.w3 + act_dim * hid + k] = cw[2][k];
    for (int o = 0; o < act_dim; ++o) { blob[L.b3 + o] = ab[2][o]; blob[L.logstd + o] = logstd[o]; }
    blob[L.b3 + act_dim] = cb[2][0];
    return SCG_OK;
}

// checkpoint blob = [int64 noise_cursor][int64 did_reset][device state blob]
int scg_state_bytes(const scg_env* e, size_t* bytes) {
    if (!e || !bytes) return SCG_EINVAL;
    *bytes = e->state_bytes + 16;
    return SCG_OK;
}
int scg_get_state(scg_env* e, void* h_blob, size_t bytes, void* stream) {
    if (!e || !h_blob || bytes != e->state_bytes + 16) return fail(e, SCG_EINVAL, "scg_get_state: bad blob size");
    CUDA_TRY(e, cudaSetDevice(e->device));
    CUDA_TRY(e, cudaStreamSynchronize((cudaStream_t)stream));
    long long hdr[2] = {e->noise_cursor, e->did_reset ? 1 : 0};
    memcpy(h_blob, hdr, 16);
    CUDA_TRY(e, cudaMemcpy((char*)h_blob + 16, e->d_state, e->state_bytes, cudaMemcpyDeviceToHost));
    return SCG_OK;
}
int scg_set_state(scg_env* e, const void* h_blob, size_t bytes, void* stream) {
    if (!e || !h_blob || bytes != e->state_bytes + 16) return fail(e, SCG_EINVAL, "scg_set_state: bad blob size");
    CUDA_TRY(e, cudaSetDevice(e->device));
    CUDA_TRY(e, cudaStreamSynchronize((cudaStream_t)stream));
    long long hdr[2]; memcpy(hdr, h_blob, 16);
    e->noise_cursor = hdr[0]; e->did_reset = hdr[1] != 0;
    CUDA_TRY(e, cudaMemcpy(e->d_state, (const char*)h_blob + 16, e->state_bytes, cudaMemcpyHostToDevice));
    return SCG_OK;
}

int scg_launch_count(const scg_env* e, uint64_t* launches) {
    if (!e || !launches) return SCG_EINVAL;
    *launches = e->launches;
    return SCG_OK;
}

int scg_philox_selftest(int32_t device, const uint32_t* h_ctr, const uint32_t* h_key, uint32_t* h_out, int32_t n) {
    if (!h_ctr || !h_key || !h_out || n <= 0) return SCG_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(nullptr, SCG_ECUDA, "scg_philox_selftest: no CUDA device");
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, SCG_ECUDA, "scg_philox_selftest: cudaSetDevice failed");
    uint32_t *dc = nullptr, *dk = nullptr, *dout = nullptr;
    cudaError_t st = cudaMalloc((void**)&dc, 16 * (size_t)n);
    if (st == cudaSuccess) st = cudaMalloc((void**)&dk, 8 * (size_t)n);
    if (st == cudaSuccess) st = cudaMalloc((void**)&dout, 16 * (size_t)n);
    if (st == cudaSuccess) st = cudaMemcpy(dc, h_ctr, 16 * (size_t)n, cudaMemcpyHostToDevice);
    if (st == cudaSuccess) st = cudaMemcpy(dk, h_key, 8 * (size_t)n, cudaMemcpyHostToDevice);
    if (st == cudaSuccess) { philox_selftest_kernel<<<(n + 127) / 128, 128>>>(dc, dk, dout, n); st = cudaGetLastError(); }
    if (st == cudaSuccess) st = cudaMemcpy(h_out, dout, 16 * (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(dc); cudaFree(dk); cudaFree(dout);
    if (st != cudaSuccess) return fail(nullptr, SCG_ECUDA, std::string("scg_philox_selftest: ") + cudaGetErrorString(st));
    return SCG_OK;
}

}  // extern "C"
