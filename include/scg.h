/*
 * scg.h -- C-ABI of the B200-native batched rollout engine for safe-control-gym's
 *          CartPole / Quadrotor-2D / Quadrotor-3D `env.step()` hot path and the PPO
 *          rollout collector.
 *
 * The reference (utiasDSL/safe-control-gym) is pure Python: there is no FFI on this
 * path.  The entry points below are what a ctypes binding inside the reference would
 * call to replace the following reference interfaces (paths relative to the reference
 * root):
 *
 *   scg_create / scg_destroy      <-  Quadrotor.__init__  safe_control_gym/envs/gym_pybullet_drones/quadrotor.py:150-326
 *                                     CartPole.__init__   safe_control_gym/envs/gym_control/cartpole.py:125-236
 *                                     BenchmarkEnv.__init__ safe_control_gym/envs/benchmark_env.py:54-191
 *   scg_set_reference             <-  X_GOAL generation   benchmark_env.py:504-713, quadrotor.py:261-323
 *   scg_set_constraints           <-  create_constraint_list  safe_control_gym/envs/constraints.py:647-665
 *   scg_set_episode_table         <-  BenchmarkEnv._randomize_values_by_info  benchmark_env.py:237-268
 *   scg_set_noise_table           <-  DisturbanceList.apply   safe_control_gym/envs/disturbances.py:54-62
 *   scg_reset                     <-  Quadrotor.reset quadrotor.py:328-392 / CartPole.reset cartpole.py:266-352
 *                                     DummyVecEnv.reset  safe_control_gym/envs/env_wrappers/vectorized_env/dummy_vec_env.py:43-48
 *   scg_step                      <-  Quadrotor.step quadrotor.py:394-445 / CartPole.step cartpole.py:238-264
 *                                     DummyVecEnv.step_wait dummy_vec_env.py:29-41 (auto-reset)
 *   scg_rollout                   <-  PPO.train_step collector loop  safe_control_gym/controllers/ppo/ppo.py:259-297
 *                                     MLPActorCritic.step  safe_control_gym/controllers/ppo/ppo_utils.py:224-231
 *                                     compute_returns_and_advantages  ppo_utils.py:374-400
 *   scg_set_adversary_control     <-  BenchmarkEnv.set_adversary_control  benchmark_env.py:216-228 (RARL rarl.py:365-368)
 *   scg_get_state / scg_set_state <-  PPO.save/load env state  ppo.py:112-148, dummy_vec_env.py:68-74
 *
 * Conventions
 *   - plain C, no C++ types, no torch types; every pointer named d_* is a DEVICE pointer
 *     owned by the caller (e.g. a torch.cuda tensor); the library only borrows it for the
 *     duration of the call.  Pointers named h_* / without prefix in setters are HOST pointers
 *     that are copied during the call.
 *   - all work is enqueued on the caller supplied CUDA stream (`stream` = cudaStream_t cast to
 *     void*, NULL = legacy default stream); no hidden synchronisation except in
 *     scg_get_state / scg_set_state / setters (which copy host tables).
 *   - return value 0 = SCG_OK, negative = error; scg_last_error() gives the message.
 *     No C++ exception crosses the ABI.  A handle is not thread-safe.
 *   - there is NO CPU fallback: scg_create fails with SCG_ECUDA when no CUDA device is usable.
 */
#ifndef SCG_H_
#define SCG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCG_ABI_VERSION 4

#define SCG_MAX_NX 12          /* max state dim (Quadrotor-3D) */
#define SCG_MAX_NU 4           /* max action dim */
#define SCG_MAX_OBS 48         /* max observation dim = nx * (1 + obs_goal_horizon) */
#define SCG_MAX_INERTIAL 4     /* quad: M,Ixx,Iyy,Izz ; cartpole: pole_length,cart_mass,pole_mass */
#define SCG_MAX_DIST 4         /* disturbances per mode */
#define SCG_MAX_DIST_DIM 12    /* dim of a disturbance target */
#define SCG_MAX_CON_ROWS 96    /* scalar constraint rows */
#define SCG_MAX_CON_NNZ 512    /* non-zeros over all constraint rows */
#define SCG_MAX_ADV 4          /* max dim of an adversary action (= dim of the disturbance mode it drives) */

/* error codes */
#define SCG_OK 0
#define SCG_EINVAL (-1)        /* bad argument / unsupported configuration (ValueError in the reference) */
#define SCG_ECUDA (-2)         /* CUDA runtime failure or no device */
#define SCG_ESTATE (-3)        /* step() before reset() -- benchmark_env.py:230-235 RuntimeError */
#define SCG_ENOMEM (-4)

/* systems (QuadType, quadrotor_utils.py:8-13; cartpole = 0) */
#define SCG_SYS_CARTPOLE 0
#define SCG_SYS_QUAD1D 1
#define SCG_SYS_QUAD2D 2
#define SCG_SYS_QUAD3D 3

/* Task / Cost (benchmark_env.py:21-32) */
#define SCG_TASK_STABILIZATION 0
#define SCG_TASK_TRAJ_TRACKING 1
#define SCG_COST_RL_REWARD 0
#define SCG_COST_QUADRATIC 1

/* state precision of the integrator (outputs are fp32 / fp64 as documented per buffer) */
#define SCG_PREC_F64 0         /* float64 state + env arithmetic, as the reference (numpy / PyBullet double) */
#define SCG_PREC_F32 1         /* float32 state + env arithmetic (throughput mode) */

/* source of random numbers */
#define SCG_RNG_TABLE 0        /* host-provided tables (scg_set_episode_table / scg_set_noise_table / d_eps) */
#define SCG_RNG_PHILOX 1       /* in-kernel Philox4x32-10 keyed by (seed, global env index) */

/* disturbance kinds (disturbances.py:277-282) */
#define SCG_DIST_IMPULSE 0
#define SCG_DIST_STEP 1
#define SCG_DIST_UNIFORM 2
#define SCG_DIST_WHITE 3
#define SCG_DIST_PERIODIC 4

/* disturbance modes, in the per-step application order action -> dynamics -> observation
 * (quadrotor.py:735-736, 419-421, 806-807) */
#define SCG_MODE_ACTION 0
#define SCG_MODE_DYNAMICS 1
#define SCG_MODE_OBSERVATION 2

/* physics update of the quadrotor (Physics enum, gym_pybullet_drones/base_aviary.py:32-40; dispatch :243-268) */
#define SCG_PHYS_PYB 0
#define SCG_PHYS_PYB_GND 1          /* + ground effect  (base_aviary.py:386-418) */
#define SCG_PHYS_PYB_DRAG 2         /* + drag           (base_aviary.py:420-440) */
#define SCG_PHYS_PYB_DW 3           /* + downwash: with ONE drone delta_z = pos[2] - pos[2] = 0, so no force is ever applied (base_aviary.py:442-460) */
#define SCG_PHYS_PYB_GND_DRAG_DW 4

/* adversary disturbance set from outside the env (BenchmarkEnv.set_adversary_control, benchmark_env.py:216-228) */
#define SCG_ADV_NONE 0
#define SCG_ADV_ACTION 1            /* added to the physical action after the passive action disturbances (quadrotor.py:737-738, cartpole.py:494-495) */
#define SCG_ADV_DYNAMICS 2          /* added to the external force / tap force, consumed by the step (quadrotor.py:416-425, cartpole.py:543-551) */

/* distribution kinds for in-kernel (Philox) reset randomisation */
#define SCG_RAND_NONE 0
#define SCG_RAND_UNIFORM 1     /* a = low, b = high */
#define SCG_RAND_NORMAL 2      /* a = loc, b = scale */

/* hidden activations of the policy MLP (neural_networks.py:8-9 get_activation) */
#define SCG_ACT_TANH 0
#define SCG_ACT_RELU 1
#define SCG_ACT_LEAKY_RELU 2   /* negative slope 0.01 (torch default) */

/* constraint row kinds */
#define SCG_CON_LINEAR 0       /* g = a.z - b           (constraints.py:274) */
#define SCG_CON_ABS 1          /* g = |a.z| - b         (SymmetricStateConstraint, constraints.py:445-447) */
#define SCG_CON_QUAD 2         /* g = sum_ij q_ij z_i z_j - b  (QuadraticConstraint x'F'PFx - b, constraints.py:222); every non-zero of the
                                * row carries the PAIR of z indices as col = i | (j << 8) */

typedef struct scg_disturbance {
    int32_t kind;              /* SCG_DIST_* */
    int32_t dim;               /* dim of the target this disturbance is added to */
    int32_t has_mask;
    int32_t step_offset;       /* impulse/step: fixed offset, or -1 = drawn per episode (disturbances.py:102-106) */
    double magnitude;          /* impulse / step */
    double duration;           /* impulse */
    double decay_rate;         /* impulse */
    double scale;              /* periodic */
    double frequency;          /* periodic */
    double low[SCG_MAX_DIST_DIM];   /* uniform */
    double high[SCG_MAX_DIST_DIM];  /* uniform */
    double std[SCG_MAX_DIST_DIM];   /* white noise */
    double mask[SCG_MAX_DIST_DIM];  /* float32-rounded 0/1 mask (disturbances.py:18) */
} scg_disturbance;

/* Flat POD description of one batch of identical environments.  Every field mirrors a
 * constructor kwarg / YAML key of the reference env (quadrotor.yaml:1-39, cartpole.yaml:1-37)
 * after host-side resolution (float32 casts already applied where the reference stores float32). */
typedef struct scg_config {
    int32_t struct_size;       /* = sizeof(scg_config), ABI guard */
    int32_t abi_version;       /* = SCG_ABI_VERSION */
    int32_t system;            /* SCG_SYS_* */
    int32_t num_envs;          /* N, envs in this handle (this rank's shard) */
    int32_t env_index_offset;  /* global index of env 0 (multi-GPU sharding: seed + global index) */
    int32_t device;            /* CUDA device ordinal */
    int32_t precision;         /* SCG_PREC_* */
    int32_t rng_mode;          /* SCG_RNG_* */
    uint64_t seed;             /* Philox key seed */

    /* timing (benchmark_env.py:139-148) */
    int32_t ctrl_freq;
    int32_t pyb_freq;
    int32_t ctrl_steps;        /* CTRL_STEPS = episode_len_sec * ctrl_freq */

    /* task */
    int32_t task;              /* SCG_TASK_* */
    int32_t cost;              /* SCG_COST_* */
    int32_t obs_goal_horizon;
    int32_t obs_wrap_angle;    /* cartpole only (cartpole.py:598-599) */
    int32_t normalized_action; /* normalized_rl_action_space */
    int32_t rew_exponential;
    int32_t done_on_out_of_bound;
    int32_t done_on_violation;
    int32_t use_constraint_penalty;
    int32_t auto_reset;        /* 1 = VecEnv semantics (dummy_vec_env.py:33-38), 0 = single env */
    int32_t _pad0;
    double constraint_penalty;
    double norm_act_scale;     /* quadrotor.py:155 */
    double goal_tolerance;     /* TASK_INFO['stabilization_goal_tolerance'] */
    double rew_state_weight[SCG_MAX_NX];
    double rew_act_weight[SCG_MAX_NU];
    double mse_weight[SCG_MAX_NX];          /* info_mse_metric_state_weight */
    double Q[SCG_MAX_NX];      /* diagonal of Q (quadratic cost), lqr_utils.py:77-99 */
    double R[SCG_MAX_NU];      /* diagonal of R */
    double u_goal[SCG_MAX_NU]; /* U_GOAL (quadrotor.py:262, cartpole.py:216) */
    double act_low[SCG_MAX_NU];   /* physical_action_bounds, float32-rounded for the quadrotor (quadrotor.py:625-626) */
    double act_high[SCG_MAX_NU];
    double hover_thrust;       /* quadrotor.py:630 (URDF mass) ; cartpole: action_scale (=10, cartpole.py:441) */
    double state_low[SCG_MAX_NX];  /* state_space bounds, float32-rounded (quadrotor.py:697, cartpole.py:461) */
    double state_high[SCG_MAX_NX];
    double x_threshold;        /* cartpole.py:452 */
    double theta_threshold;    /* cartpole.py:454 */

    /* physics constants */
    double gravity;            /* 9.8 */
    double kf, km;             /* cf2x.urdf:5 */
    double arm;                /* prop link CoM offset 0.028 (cf2x.urdf:42,54,66,78) */
    double pwm2rpm_scale, pwm2rpm_const, pwm_min, pwm_max;
    double inertial_nominal[SCG_MAX_INERTIAL];  /* quad: M,Ixx,Iyy,Izz ; cartpole: pole_length,cart_mass,pole_mass */
    double pole_width;         /* collision box side 0.05 (cartpole_template.urdf:58-66); pole inertia = m/12((2l)^2+w^2) */

    /* initial state (additive randomisation, benchmark_env.py:267) */
    double init_nominal[SCG_MAX_NX];            /* in INIT_STATE_LABELS order (quadrotor.py:209-214) */
    int32_t randomized_init;
    int32_t randomized_inertial;
    int32_t init_rand_kind[SCG_MAX_NX];         /* SCG_RAND_* , used with SCG_RNG_PHILOX */
    int32_t inertial_rand_kind[SCG_MAX_INERTIAL];
    double init_rand_a[SCG_MAX_NX], init_rand_b[SCG_MAX_NX];
    double inertial_rand_a[SCG_MAX_INERTIAL], inertial_rand_b[SCG_MAX_INERTIAL];

    /* disturbances, per mode (SCG_MODE_*) */
    int32_t n_dist[3];
    int32_t _pad1;
    scg_disturbance dist[3][SCG_MAX_DIST];

    /* non-default quadrotor physics (base_aviary.py:386-460; coefficients from cf2x.urdf:5) */
    int32_t physics;           /* SCG_PHYS_* */
    /* adversary (RARL / RAP): adv_action = clip(a, -1, 1) * scale + offset (benchmark_env.py:223-224); its dim is the dim of the
     * disturbance mode it drives (action: nu; dynamics: cartpole 2, quad 1D/2D/3D 1/2/3; benchmark_env.py:289-295) */
    int32_t adversary_mode;    /* SCG_ADV_* */
    double adversary_scale, adversary_offset;
    double gnd_eff_coeff, prop_radius, gnd_eff_h_clip;   /* base_aviary.py:130-131,406-408 */
    double drag_coeff_xy, drag_coeff_z;                  /* base_aviary.py:433-435 */
} scg_config;

/* Outputs of one control step for N envs.  Any pointer may be NULL (skipped).
 * "f64" buffers carry the environment's own arithmetic (float64 in SCG_PREC_F64). */
typedef struct scg_step_out {
    float*   d_obs;            /* [N][obs_dim]  observation returned by step (reset obs where done && auto_reset) */
    double*  d_obs_f64;        /* [N][obs_dim]  same, float64 */
    double*  d_state_f64;      /* [N][nx]       env.state after the step (post auto-reset where done) */
    float*   d_rew;            /* [N] */
    double*  d_rew_f64;        /* [N] */
    uint8_t* d_done;           /* [N] */
    uint8_t* d_truncated;      /* [N] info['TimeLimit.truncated'] (benchmark_env.py:499-501) */
    float*   d_cviol;          /* [N] info['constraint_violation'] 0/1 */
    double*  d_cvals;          /* [N][n_con] info['constraint_values'] */
    double*  d_mse;            /* [N] info['mse'] */
    uint8_t* d_out_of_bounds;  /* [N] */
    uint8_t* d_goal_reached;   /* [N] */
    int32_t* d_current_step;   /* [N] info['current_step'] of the transition just made */
    float*   d_terminal_obs;   /* [N][obs_dim] info['terminal_observation'] (valid where done, auto_reset) */
    double*  d_terminal_obs_f64;
    double*  d_terminal_state_f64; /* [N][nx] state before the auto-reset */
} scg_step_out;

/* policy-MLP engines of the fused collector */
#define SCG_MLP_AUTO 0         /* library default */
#define SCG_MLP_FFMA2 1        /* CUDA cores: register-tiled FP32 GEMMs with packed FFMA2 */
#define SCG_MLP_TCGEN05 2      /* 5th-gen tensor cores: tcgen05.mma kind::f16 on fp16 hi + scaled-lo operands (22 significant bits), FP32 accumulators in
                                * tensor memory; hidden 128: two M128 tiles per layer, hidden 64: actor + critic fused into one width-128 net */
#define SCG_MLP_TCGEN05_TF32 3 /* hidden 64 only: tcgen05.mma kind::tf32, 3xTF32 hi/lo split (twice the MMAs of the fp16 form; kept for comparison) */

/* Output buffers of the fused rollout, PPOBuffer scheme (ppo_utils.py:266-295): [T][N][.] float32. */
#define SCG_MAX_PEERS 8
#define SCG_MAX_SAFETY 16        /* constraint models of a safety layer (SafeExplorer) */

typedef struct scg_rollout_out {
    float* d_obs;              /* [T][N][obs_dim] observation fed to the policy at step t */
    float* d_act;              /* [T][N][act_dim] */
    float* d_rew;              /* [T][N]  (after scg_rollout: rew + gamma*terminal_v, ppo_utils.py:389, if ret/adv requested) */
    float* d_mask;             /* [T][N]  1 - done */
    float* d_v;                /* [T][N] */
    float* d_logp;             /* [T][N] */
    float* d_terminal_v;       /* [T][N] */
    float* d_ret;              /* [T][N] nullable */
    float* d_adv;              /* [T][N] nullable (un-normalised; ppo.py:300 normalisation is a batch reduction done by the caller) */
    float* d_cviol;            /* [T][N] nullable, info['constraint_violation'] */
    float* d_cvals;            /* [T][N][n_con] nullable */
    float* d_mse;              /* [T][N] nullable */
    float* d_last_obs;         /* [N][obs_dim] nullable: observation after the last step */
    float* d_last_val;         /* [N] nullable: critic(last_obs) (ppo.py:289) */
    /* episode statistics of episodes that FINISHED inside this rollout (record_episode_statistics.py:139-166) */
    float*   d_ep_return_sum;  /* [N] nullable */
    float*   d_ep_length_sum;  /* [N] nullable */
    float*   d_ep_cviol_sum;   /* [N] nullable */
    int32_t* d_ep_count;       /* [N] nullable */
    /* with obs / reward normalisation (scg_rollout_args.d_norm): what the running normalisers of the controller must be updated with
     * after the rollout (normalization.py:106-113,148-159) */
    float* d_c;                /* [T][N][safety_constraints] nullable: the constraint values fed to the safety layer at step t (SafePPOBuffer 'c') */
    float* d_obs_raw;          /* [T][N][obs_dim] nullable: the UN-normalised observation env.step returned at step t (reset obs where done) */
    float* d_rew_raw;          /* [T][N] nullable: the UN-normalised reward */
} scg_rollout_out;

/* Arguments of the fused rollout. */
typedef struct scg_rollout_args {
    int32_t T;                 /* rollout_steps */
    int32_t obs_dim, act_dim;  /* must match the env */
    int32_t hidden_dim;        /* 32, 64 or 128 */
    int32_t activation;        /* SCG_ACT_* */
    int32_t use_gae;           /* ppo.yaml:11 */
    int32_t deterministic;     /* 1 = act() = mean (ppo_utils.py:233-238), logp of the mean */
    int32_t envs_per_block;    /* 0 = library default; bits 0-15 = envs per CTA, bits 16-31 = threads per CTA (0 = 8 per env) */
    double gamma, gae_lambda;
    const float* d_policy;     /* packed policy blob, layout of scg_policy_pack() */
    const float* d_eps;        /* [T][N][act_dim] standard-normal draws, or NULL -> Philox */
    /* Fused rollout + all-gather over NVLink (the single collective of the path): every store to the PPOBuffer keys
     * obs / act / rew / mask / v / logp / terminal_v / ret / adv is repeated at `pointer + peer_byte_offset[g]` for g < n_peers,
     * i.e. into the same slot of each peer GPU's gathered buffer (peer-mapped memory, e.g. torch symmetric memory).  Those keys
     * must then live in ONE allocation per GPU so that one offset per peer fits them all.  n_peers = 0: local stores only.  The
     * caller orders the peers' reads after the kernel with its own barrier (stream-ordered). */
    int32_t n_peers;
    int32_t mlp_engine;        /* SCG_MLP_*: which units run the policy MLP of the fused collector */
    int64_t peer_byte_offset[SCG_MAX_PEERS];
    /* NVSwitch multicast variant of the same gather: if non-zero, `pointer + multicast_byte_offset` is the multicast mapping of the
     * gathered buffer (one `multimem.st` lands in EVERY rank's copy, this rank's included), and it replaces the n_peers unicast
     * copies: one extra store per record entry instead of n_peers.  The plain local store is still issued (the kernel reads its own
     * record back for the GAE scan through the unicast address). */
    int64_t multicast_byte_offset;
    /* Observation / reward normalisers of the controller (MeanStdNormalizer / RewardStdNormalizer, math_and_models/normalization.py:88-159;
     * applied in ppo.py:270-271) with statistics FROZEN for the duration of the rollout (the controller updates them once per rollout
     * from d_obs_raw / d_rew_raw).  d_norm = NULL: off.  Layout (float64): mean[obs_dim], inv_std[obs_dim] = 1/sqrt(var + eps),
     * clip_obs, obs_on (0/1), rew_scale = 1/sqrt(var_ret + eps), clip_rew, rew_on (0/1).  The terminal observation fed to the critic
     * stays un-normalised, as in the reference (ppo.py:279-282). */
    const double* d_norm;
    const float* d_obs0;       /* [N][obs_dim] nullable: policy input of step 0 (default: the observation the engine carried over) */
    /* Safety layer of the SafeExplorer collector (safe_control_gym/controllers/safe_explorer/safe_explorer_utils.py:15-176, applied inside the
     * actor: safe_ppo_utils.py:88-104): one constraint model per state-constraint row i < safety_constraints, each an MLP
     * obs -> relu(hidden) -> act_dim; the mean action is projected  mu_safe = mu - lambda* g_i*(obs),
     * lambda_i = relu((g_i . mu + c_i + slack_i) / (g_i . g_i + 1e-8)), i* = argmax lambda (Dalal et al. 2018, eq. 5-6), c = the first
     * `safety_constraints` constraint values of the current state; sampling and log-prob use mu_safe.  0 = no safety layer.
     * d_safety: per model { W1[hidden][obs_dim], b1[hidden], W2[act_dim][hidden], b2[act_dim] } (nn.Linear layout), models back to back. */
    int32_t safety_constraints;
    int32_t safety_hidden;
    const float* d_safety;
    double safety_slack[SCG_MAX_SAFETY];
    /* Adversary of the RARL / RAP collectors (safe_control_gym/controllers/rarl/rarl.py:349-418, rap.py): a second actor-critic that sees the
     * same observation; its sampled action goes through set_adversary_control (clip, scale, offset) into the env step of the same control
     * step.  d_adv_policy = NULL: no adversary.  The blob has the scg_policy_pack() layout with act_dim = adv_act_dim (which must equal
     * scg_adversary_dim()) and the same hidden_dim / activation as the protagonist.  adv_record = 0: the record is the protagonist's
     * (act, v, logp, rew, terminal_v / bootstrap from its critic); 1: the adversary's (its act [T][N][adv_act_dim], v, logp, rew = -rew,
     * terminal_v / bootstrap from ITS critic; rarl.py:391-399).  d_eps_adv: [T][N][adv_act_dim] standard normals or NULL -> Philox. */
    const float* d_adv_policy;
    const float* d_eps_adv;
    int32_t adv_act_dim;
    int32_t adv_record;
} scg_rollout_args;

typedef struct scg_env scg_env;   /* opaque, one per (device, batch) */

/* lifecycle */
int  scg_abi_version(void);
int  scg_create(const scg_config* cfg, scg_env** out);
void scg_destroy(scg_env* env);
const char* scg_last_error(const scg_env* env);   /* env may be NULL: last error of a failed scg_create */

/* dimensions resolved from the config */
int  scg_dims(const scg_env* env, int32_t* nx, int32_t* nu, int32_t* obs_dim, int32_t* n_con,
              int32_t* episode_table_width, int32_t* noise_table_width);

/* host tables (copied to the device during the call) */
int  scg_set_reference(scg_env* env, const double* x_goal /*[L][nx]*/, int32_t L);
int  scg_set_constraints(scg_env* env, int32_t n_rows, const int32_t* row_kind /*[n_rows] SCG_CON_* */,
                         const int32_t* row_strict /*[n_rows]*/, const int32_t* row_ptr /*[n_rows+1] CSR*/,
                         const int32_t* col /*[nnz] index into z=[state;input]*/, const double* val /*[nnz]*/,
                         const double* b /*[n_rows]*/);
int  scg_set_episode_table(scg_env* env, const double* tab /*[K][N][P]*/, int32_t K);
int  scg_set_noise_table(scg_env* env, const double* tab /*[Tn][N][D]*/, int32_t Tn);

/* env API */
int  scg_reset(scg_env* env, const uint8_t* d_mask /*[N] or NULL = all*/, float* d_obs /*[N][obs_dim] nullable*/,
               double* d_obs_f64 /*nullable*/, double* d_state_f64 /*nullable*/, void* stream);
int  scg_step(scg_env* env, const float* d_act /*[N][nu] nullable*/, const double* d_act_f64 /*[N][nu] nullable*/,
              const scg_step_out* out, void* stream);
int  scg_rollout(scg_env* env, const scg_rollout_args* args, const scg_rollout_out* out, void* stream);
/* BenchmarkEnv.set_adversary_control (benchmark_env.py:216-228) for all N envs: d_adv [N][scg_adversary_dim()] raw adversary actions.
 * The clipped, scaled action is kept on the device and applied by the next scg_step (dynamics mode: consumed by that step; action mode:
 * kept until the env resets, like the reference attribute).  SCG_ESTATE ("adversary_disturbance does not exist") without an adversary. */
int  scg_set_adversary_control(scg_env* env, const float* d_adv, void* stream);
int  scg_adversary_dim(const scg_env* env);   /* 0 = the env has no adversary */

/* policy blob helpers (host side, pure functions) */
int  scg_policy_blob_floats(int32_t obs_dim, int32_t act_dim, int32_t hidden_dim);
int  scg_policy_pack(int32_t obs_dim, int32_t act_dim, int32_t hidden_dim,
                     const float* const actor_w[3], const float* const actor_b[3], const float* logstd,
                     const float* const critic_w[3], const float* const critic_b[3],
                     float* blob_out /* host, scg_policy_blob_floats() floats */);

/* checkpoint / resume of the device env state (synchronises the stream) */
int  scg_state_bytes(const scg_env* env, size_t* bytes);
int  scg_get_state(scg_env* env, void* h_blob, size_t bytes, void* stream);
int  scg_set_state(scg_env* env, const void* h_blob, size_t bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused PPO update: PPOAgent.update, safe_control_gym/controllers/ppo/ppo_utils.py:82-146 --
 * opt_epochs x num_mini_batch sequential minibatch steps, each = actor loss (clipped surrogate + entropy bonus, :56-69),
 * KL gate `approx_kl <= 1.5 * target_kl` (:128), critic loss (:71-80), backward through both MLPs and two Adam steps
 * (torch.optim.Adam defaults, :44-45) -- in ONE cooperative kernel launch.  Handle-free: runs on the current device.
 *
 * Training parameter blob ("train layout", floats; every section starts at a multiple of 4):
 *   actor : W1[H][obs] b1[H] W2[H][H] b2[H] W3[A][H] b3[A] logstd[A]      (nn.Linear layout, ppo_utils.py:149-177)
 *   critic: W1[H][obs] b1[H] W2[H][H] b2[H] W3[1][H] b3[1]                 (ppo_utils.py:180-194)
 * scg_ppo_param_layout() returns the 13 section offsets + [13] = first critic float + [14] = total.
 * ------------------------------------------------------------------------------------------- */
#define SCG_PPO_KEY_OBS 0
#define SCG_PPO_KEY_ACT 1
#define SCG_PPO_KEY_LOGP 2
#define SCG_PPO_KEY_ADV 3
#define SCG_PPO_KEY_RET 4
#define SCG_PPO_KEY_V 5
typedef struct scg_ppo_args {
    int32_t obs_dim, act_dim, hidden_dim, activation;
    int32_t mini_batch_size, num_mini_batch, opt_epochs, use_clipped_value;   /* ppo.yaml:13-22 */
    /* the record: row r of key k lives at d_key[k] + (r / block_rows) * block_stride[k] + (r % block_rows) * width(k)
     * (one block = one rank's [T*n][width] slab of the all-gathered rollout record; a single-GPU record is one block) */
    int64_t block_rows;
    int64_t block_stride[6];
    const float* d_key[6];         /* SCG_PPO_KEY_*: obs[.][obs_dim], act[.][act_dim], logp, adv (already normalised, ppo.py:300), ret, v */
    const int32_t* d_perm;         /* [opt_epochs][perm_stride] row indices; minibatch i of epoch e = perm[e][i*mb .. (i+1)*mb) (ppo_utils.py:342-355) */
    int64_t perm_stride;
    double clip_param, target_kl, entropy_coef;
    double actor_lr, critic_lr, beta1, beta2, adam_eps;
    float* d_params;               /* train layout, updated in place */
    float* d_exp_avg;              /* Adam first moments, train layout */
    float* d_exp_avg_sq;           /* Adam second moments, train layout */
    int64_t* d_adam_steps;         /* [2] optimizer step counts: actor (advances only when the KL gate passes), critic */
    float* d_stats;                /* [8] += { sum policy_loss, sum value_loss, sum entropy_loss, sum approx_kl, actor steps taken, minibatches, 0, 0 } */
    float* d_workspace;            /* scg_ppo_workspace_bytes() bytes of device scratch */
    size_t workspace_bytes;
    float* d_grad_out;             /* nullable: reduced gradient of the LAST minibatch in train layout (tests) */
} scg_ppo_args;

int    scg_ppo_param_layout(int32_t obs_dim, int32_t act_dim, int32_t hidden_dim, int32_t offsets_out[15]);
size_t scg_ppo_workspace_bytes(const scg_ppo_args* args);
int    scg_ppo_update(const scg_ppo_args* args, void* stream);
/* train layout (device) -> packed rollout blob of scg_policy_pack() (device), one small launch */
int    scg_policy_pack_device(int32_t obs_dim, int32_t act_dim, int32_t hidden_dim, const float* d_params, float* d_blob, void* stream);

/* introspection used by the benchmark: how many kernels this handle launched so far */
int  scg_launch_count(const scg_env* env, uint64_t* launches);

/* Philox4x32-10 known-answer hook (tests): runs the device generator for n counters */
int  scg_philox_selftest(int32_t device, const uint32_t* h_ctr /*[n][4]*/, const uint32_t* h_key /*[n][2]*/,
                         uint32_t* h_out /*[n][4]*/, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* SCG_H_ */
