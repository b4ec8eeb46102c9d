/*
 * scg_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C float64 restatement of what the reference (utiasDSL/safe-control-gym) computes on the
 * env.step() + PPO-rollout hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load this library; the product (safe_control_gym_b200)
 * never does and has no CPU fallback.
 *
 * Parity status of this oracle:
 *   - CartPole physics: PINNED against 5208 real PyBullet transitions shipped inside the reference
 *     (examples/cbf/models/cbf_nn_ppo_model_cartpole.pt, produced by safety_filters/cbf/cbf_nn.py:324-380),
 *     committed as tests/golden/cartpole_pybullet_transitions.npz; max |err| 1e-7 (float32 storage).
 *   - X_GOAL tables: PINNED against the observations stored in examples/rl/models/ppo/*.pt.
 *   - MLP forward: PINNED against torch on the shipped checkpoints.
 *   - Env logic around the integrator (both systems): PINNED against the reference's OWN code run here -- quadrotor.py, base_aviary.py,
 *     cartpole.py, benchmark_env.py, constraints.py, disturbances.py, unmodified, with PyBullet replaced by a stand-in (tools/ref_harness.py):
 *     24 recorded cases (tests/golden/refloop_*.npz, tests/test_refloop_golden.py) reproduced to 2e-9 -- observations, state, reward, done,
 *     truncation, mse, constraint values, flags, X_GOAL; randomisation order and arithmetic; all five disturbance kinds; the four non-default
 *     physics variants; the adversary input.
 *   - Collector loops: PINNED against the buffers the reference's own PPO.train_step(), RARL.collect_rollouts() and SafeExplorerPPO.train_step()
 *     filled (tests/golden/{ppo,rarl,safe}_loop_*.npz, tests/test_{ppo,rarl,safe}_loop_golden.py), PPO.run() evaluation results, GAE against the reference function, sampling against MLPActorCritic.step (tests/test_ppo_reference_golden.py).
 *   - Quadrotor physics vs PyBullet: UNPINNED by data (no PyBullet wheel, no recorded quadrotor
 *     transitions in the reference tree).  It restates the same Bullet btMultiBody semantics that the
 *     cartpole pin validates (symplectic Euler per sub-step, forces re-applied every sub-step) and is
 *     checked against analytic invariants (hover, free fall, torque-free precession) and the
 *     convergence of the sub-stepped integrator to the reference's own symbolic ODE.
 *
 * Reference files followed (paths relative to the reference root):
 *   BE = safe_control_gym/envs/benchmark_env.py          QD = safe_control_gym/envs/gym_pybullet_drones/quadrotor.py
 *   BA = safe_control_gym/envs/gym_pybullet_drones/base_aviary.py   QU = .../quadrotor_utils.py
 *   CP = safe_control_gym/envs/gym_control/cartpole.py   CO = safe_control_gym/envs/constraints.py
 *   DI = safe_control_gym/envs/disturbances.py           VE = safe_control_gym/envs/env_wrappers/vectorized_env/dummy_vec_env.py
 *   PU = safe_control_gym/controllers/ppo/ppo_utils.py   PPO = safe_control_gym/controllers/ppo/ppo.py
 * PyBullet (pybullet ^3.2, pyproject.toml:22) is a third-party dependency that is not vendored; its
 * btMultiBody step is restated from the published Bullet3 algorithm (Featherstone ABA + symplectic
 * Euler, exponential-map base orientation update).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "scg.h"

#ifdef _OPENMP
#include <omp.h>
#endif

#define PI_D 3.14159265358979323846

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11 "Parallel random numbers: as easy as 1, 2, 3").          */
/* Stream definition shared with the CUDA engine ("random numbers").                 */
/* ------------------------------------------------------------------------------------------ */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void scg_oracle_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out, int n) {
    for (int i = 0; i < n; ++i) philox4x32_10(ctr + 4 * i, key + 2 * i, out + 4 * i);
}

#define STREAM_EPISODE 1u
#define STREAM_STEP 2u
#define STREAM_POLICY 3u
#define STREAM_RESET 4u
#define STREAM_ADVERSARY 5u   /* sampling noise of the RARL / RAP adversary policy */

/* draw slot j of (stream, idx) = words 2(j&1), 2(j&1)+1 of Philox block j>>1: counter = (j>>1, idx, global env, stream) */
static void draw_words(uint64_t seed, uint32_t stream, uint32_t genv, uint32_t idx, uint32_t slot, uint32_t* w0, uint32_t* w1) {
    uint32_t ctr[4] = {slot >> 1, idx, genv, stream};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t x[4];
    philox4x32_10(ctr, key, x);
    *w0 = (slot & 1u) ? x[2] : x[0]; *w1 = (slot & 1u) ? x[3] : x[1];
}
static double u01(uint32_t x) { return ((double)x + 0.5) * (1.0 / 4294967296.0); }
static double draw_uniform(uint64_t seed, uint32_t stream, uint32_t genv, uint32_t idx, uint32_t slot) {
    uint32_t w0, w1; draw_words(seed, stream, genv, idx, slot, &w0, &w1); return u01(w0);
}
static double draw_normal(uint64_t seed, uint32_t stream, uint32_t genv, uint32_t idx, uint32_t slot) {
    uint32_t w0, w1; draw_words(seed, stream, genv, idx, slot, &w0, &w1);
    return sqrt(-2.0 * log(u01(w0))) * cos(2.0 * PI_D * u01(w1));
}
static double draw_randint(uint64_t seed, uint32_t stream, uint32_t genv, uint32_t idx, uint32_t slot, uint32_t n) {
    uint32_t w0, w1; draw_words(seed, stream, genv, idx, slot, &w0, &w1);
    return (double)(uint32_t)(((uint64_t)w0 * (uint64_t)n) >> 32);
}

/* ------------------------------------------------------------------------------------------ */
typedef struct {
    /* rigid body (quadrotor, Bullet world frame) */
    double pos[3], quat[4] /* x,y,z,w */, vel[3], omega[3] /* world */, rpy[3];
    /* cartpole generalized coordinates */
    double q[2], qd[2];
    /* per-episode physical parameters */
    double inertial[SCG_MAX_INERTIAL];
    /* env bookkeeping */
    double state[SCG_MAX_NX];
    double obs[SCG_MAX_OBS];
    double noisy_act[SCG_MAX_NU];   /* current_noisy_physical_action */
    double clipped_act[SCG_MAX_NU]; /* current_clipped_action */
    int ctrl_step, pyb_step;
    int initial_reset;
    int episode;                    /* number of resets done so far == index of the next episode row */
    uint32_t total_steps;           /* control steps taken since creation */
    int dist_offset[3][SCG_MAX_DIST];
    double ep_return, ep_length, ep_cviol;
    double adv_action[SCG_MAX_ADV]; /* self.adv_action (BE:224); adv_set = 0 <=> None */
    int adv_set;
} oenv;

typedef struct scg_oracle {
    scg_config cfg;
    int nx, nu, obs_dim, n_inertial;
    int L; double* x_goal;                       /* [L][nx]; stabilisation: L = 1 */
    /* constraints (CSR rows over z = [state; input]) */
    int n_con; int *row_kind, *row_strict, *row_ptr, *col; double *val, *b;
    /* tables */
    int K, P; double* ep_table;                  /* [K][N][P] */
    int Tn, D; double* noise_table;              /* [Tn][N][D] */
    int stale_force_point;                       /* see scg_oracle_set_stale_force_point */
    int n_off_slots;                             /* number of randomised step offsets */
    long noise_cursor;
    oenv* envs;
} scg_oracle;

static int sys_nx(int s) { return s == SCG_SYS_CARTPOLE ? 4 : s == SCG_SYS_QUAD1D ? 2 : s == SCG_SYS_QUAD2D ? 6 : 12; }
static int sys_nu(int s) { return s == SCG_SYS_CARTPOLE ? 1 : s == SCG_SYS_QUAD1D ? 1 : s == SCG_SYS_QUAD2D ? 2 : 4; }

static int obs_dim_of(const scg_config* c) {
    int nx = sys_nx(c->system);
    /* BE:432-443, QD:700-708 */
    if (c->cost == SCG_COST_RL_REWARD && c->obs_goal_horizon > 0) {
        if (c->task == SCG_TASK_TRAJ_TRACKING) return nx * (1 + c->obs_goal_horizon);
        return nx * 2;
    }
    return nx;
}

static int noise_dims(const scg_config* c) {
    int d = 0;
    for (int m = 0; m < 3; ++m)
        for (int j = 0; j < c->n_dist[m]; ++j) {
            int k = c->dist[m][j].kind;
            if (k == SCG_DIST_UNIFORM || k == SCG_DIST_WHITE || k == SCG_DIST_PERIODIC) d += c->dist[m][j].dim;
        }
    return d;
}
static int offset_slots(const scg_config* c) {
    int d = 0;
    for (int m = 0; m < 3; ++m)
        for (int j = 0; j < c->n_dist[m]; ++j) {
            int k = c->dist[m][j].kind;
            if ((k == SCG_DIST_IMPULSE || k == SCG_DIST_STEP) && c->dist[m][j].step_offset < 0) d += 1;
        }
    return d;
}

scg_oracle* scg_oracle_create(const scg_config* cfg) {
    if (!cfg || cfg->struct_size != (int32_t)sizeof(scg_config)) return NULL;
    scg_oracle* o = (scg_oracle*)calloc(1, sizeof(scg_oracle));
    o->cfg = *cfg;
    o->nx = sys_nx(cfg->system); o->nu = sys_nu(cfg->system);
    o->obs_dim = obs_dim_of(cfg);
    o->n_inertial = cfg->system == SCG_SYS_CARTPOLE ? 3 : 4;
    o->n_off_slots = offset_slots(cfg);
    o->P = o->n_off_slots + SCG_MAX_INERTIAL + o->nx;
    o->D = noise_dims(cfg);
    o->envs = (oenv*)calloc((size_t)cfg->num_envs, sizeof(oenv));
    return o;
}
void scg_oracle_destroy(scg_oracle* o) {
    if (!o) return;
    free(o->x_goal); free(o->row_kind); free(o->row_strict); free(o->row_ptr); free(o->col); free(o->val); free(o->b);
    free(o->ep_table); free(o->noise_table); free(o->envs); free(o);
}
void scg_oracle_dims(const scg_oracle* o, int* nx, int* nu, int* obs_dim, int* n_con, int* P, int* D) {
    *nx = o->nx; *nu = o->nu; *obs_dim = o->obs_dim; *n_con = o->n_con; *P = o->P; *D = o->D;
}
void scg_oracle_set_reference(scg_oracle* o, const double* xg, int L) {
    free(o->x_goal); o->L = L; o->x_goal = (double*)malloc(sizeof(double) * (size_t)L * o->nx);
    memcpy(o->x_goal, xg, sizeof(double) * (size_t)L * o->nx);
}
void scg_oracle_set_constraints(scg_oracle* o, int n_rows, const int32_t* kind, const int32_t* strict, const int32_t* row_ptr,
                                const int32_t* col, const double* val, const double* b) {
    free(o->row_kind); free(o->row_strict); free(o->row_ptr); free(o->col); free(o->val); free(o->b);
    o->n_con = n_rows;
    int nnz = n_rows ? row_ptr[n_rows] : 0;
    o->row_kind = (int*)malloc(sizeof(int) * (n_rows + 1)); o->row_strict = (int*)malloc(sizeof(int) * (n_rows + 1));
    o->row_ptr = (int*)malloc(sizeof(int) * (n_rows + 1)); o->col = (int*)malloc(sizeof(int) * (nnz + 1));
    o->val = (double*)malloc(sizeof(double) * (nnz + 1)); o->b = (double*)malloc(sizeof(double) * (n_rows + 1));
    for (int i = 0; i < n_rows; ++i) { o->row_kind[i] = kind[i]; o->row_strict[i] = strict[i]; o->b[i] = b[i]; }
    for (int i = 0; i <= n_rows; ++i) o->row_ptr[i] = n_rows ? row_ptr[i] : 0;
    for (int i = 0; i < nnz; ++i) { o->col[i] = col[i]; o->val[i] = val[i]; }
}
void scg_oracle_set_episode_table(scg_oracle* o, const double* tab, int K) {
    free(o->ep_table); o->K = K; size_t n = (size_t)K * o->cfg.num_envs * o->P;
    o->ep_table = (double*)malloc(sizeof(double) * n); memcpy(o->ep_table, tab, sizeof(double) * n);
}
/* Reference quirk, OFF by default (the CUDA path does not have it; tests/test_refloop_golden.py switches it on to show that the restatement
 * reproduces the reference with it).  BA:268-279 applies the dynamics disturbance force with flags=WORLD_FRAME at posObj = self.pos, and under
 * physics 'pyb' self.pos is only refreshed at the END of a control step (BA:246-251 refreshes per sub-step for the other physics variants only):
 * from the second sub-step on the force acts at the position the drone had when the control step began, i.e. with the lever arm
 * (pos_at_step_start - pos_now) about the centre of mass -- a small torque the "force through the CoM" reading does not have. */
void scg_oracle_set_stale_force_point(scg_oracle* o, int on) { o->stale_force_point = on; }

/* Snapshot / restore of everything a step changes (all envs + the noise-table cursor): the counterpart of scg_get_state / scg_set_state, used by the
 * tests that replay a rollout from the same start (e.g. the RAP composition, tests/test_rap_reference.py). */
size_t scg_oracle_state_bytes(const scg_oracle* o) { return sizeof(oenv) * (size_t)o->cfg.num_envs + sizeof(long long); }
void scg_oracle_get_state(const scg_oracle* o, void* buf) {
    memcpy(buf, o->envs, sizeof(oenv) * (size_t)o->cfg.num_envs);
    long long cur = (long long)o->noise_cursor; memcpy((char*)buf + sizeof(oenv) * (size_t)o->cfg.num_envs, &cur, sizeof(cur));
}
void scg_oracle_set_state(scg_oracle* o, const void* buf) {
    memcpy(o->envs, buf, sizeof(oenv) * (size_t)o->cfg.num_envs);
    long long cur; memcpy(&cur, (const char*)buf + sizeof(oenv) * (size_t)o->cfg.num_envs, sizeof(cur)); o->noise_cursor = cur;
}

void scg_oracle_set_noise_table(scg_oracle* o, const double* tab, int Tn) {
    free(o->noise_table); o->Tn = Tn; size_t n = (size_t)Tn * o->cfg.num_envs * (o->D > 0 ? o->D : 1);
    o->noise_table = (double*)malloc(sizeof(double) * n); memcpy(o->noise_table, tab, sizeof(double) * (size_t)Tn * o->cfg.num_envs * o->D);
    o->noise_cursor = 0;
}

/* ------------------------------------------------------------------------------------------ */
/* small math helpers                                                                           */
/* ------------------------------------------------------------------------------------------ */
/* normalization.py:8-10 ; numpy `%` is a floored modulo */
static double normalize_angle(double x) {
    double y = fmod(x + PI_D, 2.0 * PI_D);
    if (y < 0) y += 2.0 * PI_D;
    return y - PI_D;
}
/* np.round(x, 8) (CO:109): numpy computes rint(x * 1e8) / 1e8 */
static double round8(double x) { return rint(x * 1e8) / 1e8; }

/* Bullet btQuaternion::setEulerZYX as used by pybullet.getQuaternionFromEuler (QD:381) */
static void quat_from_rpy(const double rpy[3], double q[4]) {
    double cr = cos(rpy[0] * 0.5), sr = sin(rpy[0] * 0.5);
    double cp = cos(rpy[1] * 0.5), sp = sin(rpy[1] * 0.5);
    double cy = cos(rpy[2] * 0.5), sy = sin(rpy[2] * 0.5);
    q[0] = sr * cp * cy - cr * sp * sy;
    q[1] = cr * sp * cy + sr * cp * sy;
    q[2] = cr * cp * sy - sr * sp * cy;
    q[3] = cr * cp * cy + sr * sp * sy;
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] /= n;
}
/* Bullet btMatrix3x3::setRotation: body -> world rotation, row-major (pybullet.getMatrixFromQuaternion, QD:795) */
static void rot_from_quat(const double q[4], double R[9]) {
    double d = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    double s = 2.0 / d;
    double xs = q[0] * s, ys = q[1] * s, zs = q[2] * s;
    double wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
    double xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs;
    double yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
    R[0] = 1.0 - (yy + zz); R[1] = xy - wz;          R[2] = xz + wy;
    R[3] = xy + wz;          R[4] = 1.0 - (xx + zz); R[5] = yz - wx;
    R[6] = xz - wy;          R[7] = yz + wx;          R[8] = 1.0 - (xx + yy);
}
/* Bullet btQuaternion::getEulerZYX as used by pybullet.getEulerFromQuaternion (BA:329) */
static void rpy_from_quat(const double q[4], double rpy[3]) {
    double sqx = q[0] * q[0], sqy = q[1] * q[1], sqz = q[2] * q[2], squ = q[3] * q[3];
    double sarg = -2.0 * (q[0] * q[2] - q[3] * q[1]);
    if (sarg <= -0.99999) { rpy[1] = -0.5 * PI_D; rpy[0] = 0; rpy[2] = 2 * atan2(q[0], -q[1]); }
    else if (sarg >= 0.99999) { rpy[1] = 0.5 * PI_D; rpy[0] = 0; rpy[2] = 2 * atan2(-q[0], q[1]); }
    else {
        rpy[1] = asin(sarg);
        rpy[0] = atan2(2 * (q[1] * q[2] + q[3] * q[0]), squ - sqx - sqy + sqz);
        rpy[2] = atan2(2 * (q[0] * q[1] + q[3] * q[2]), squ + sqx - sqy - sqz);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* physics: one Bullet stepSimulation() of dt for the free-flying Crazyflie (BA:252-282, 364-384) */
/* ------------------------------------------------------------------------------------------ */
static void quad_substep(const scg_config* c, oenv* e, const double rpm[4], const double fdist[3], int has_dist, double dt, const double* force_point) {
    const double m = e->inertial[0];
    const double J[3] = {e->inertial[1], e->inertial[2], e->inertial[3]};
    double R[9]; rot_from_quat(e->quat, R);
    /* BA:370-372: per-motor thrust along link z, yaw torque about body z */
    double f[4], tq[4];
    for (int i = 0; i < 4; ++i) { f[i] = rpm[i] * rpm[i] * c->kf; tq[i] = rpm[i] * rpm[i] * c->km; }
    double z_torque = -tq[0] + tq[1] - tq[2] + tq[3];
    /* prop link CoM offsets (cf2x.urdf:42,54,66,78): r x F with F = (0,0,f) in the body frame */
    const double a = c->arm;
    const double rx[4] = {a, -a, -a, a}, ry[4] = {a, a, -a, -a};
    const int gnd = c->physics == SCG_PHYS_PYB_GND || c->physics == SCG_PHYS_PYB_GND_DRAG_DW;
    const int drag = c->physics == SCG_PHYS_PYB_DRAG || c->physics == SCG_PHYS_PYB_GND_DRAG_DW;
    /* Physics.PYB_DW: _downwash() compares the drone with ITSELF (delta_z = pos[2] - pos[2] = 0, BA:448-451): never a force. */
    if (gnd) {
        /* _ground_effect (BA:386-418): kinematic info is refreshed at the start of every sub-step (BA:246-251); per-propeller force
         * KF rpm^2 GND_EFF_COEFF (PROP_RADIUS / (4 h_i))^2 along the LINK z axis at the prop link, h_i = world height of prop link i
         * clipped from below, only while |roll| and |pitch| < pi/2.  Same application point and direction as the thrust: f_i grows. */
        double rpy[3]; rpy_from_quat(e->quat, rpy);
        if (fabs(rpy[0]) < PI_D / 2 && fabs(rpy[1]) < PI_D / 2) {
            for (int i = 0; i < 4; ++i) {
                double h = e->pos[2] + R[6] * rx[i] + R[7] * ry[i];
                if (h < c->gnd_eff_h_clip) h = c->gnd_eff_h_clip;
                const double q = c->prop_radius / (4 * h);
                f[i] += rpm[i] * rpm[i] * c->kf * c->gnd_eff_coeff * q * q;
            }
        }
    }
    double tau_b[3] = {0, 0, z_torque}, thrust = 0;
    for (int i = 0; i < 4; ++i) { tau_b[0] += ry[i] * f[i]; tau_b[1] += -rx[i] * f[i]; thrust += f[i]; }
    /* world force: thrust along body z + gravity + disturbance (BA:271-279, WORLD_FRAME through the CoM) */
    double F[3] = {R[2] * thrust, R[5] * thrust, R[8] * thrust - m * c->gravity};
    if (drag) {
        /* _drag (BA:420-440): drag_factors = -DRAG_COEFF * sum(2 pi rpm / 60); the reference rotates (drag_factors * vel_world) by the
         * base rotation and then applies the result as a LINK_FRAME force at the CoM link, i.e. it is rotated once more: R R (k v). */
        double srpm = 0; for (int i = 0; i < 4; ++i) srpm += 2 * PI_D * rpm[i] / 60;
        const double k[3] = {-c->drag_coeff_xy * srpm, -c->drag_coeff_xy * srpm, -c->drag_coeff_z * srpm};
        const double kv[3] = {k[0] * e->vel[0], k[1] * e->vel[1], k[2] * e->vel[2]};
        double d1[3], d2[3];
        for (int r = 0; r < 3; ++r) d1[r] = R[3 * r] * kv[0] + R[3 * r + 1] * kv[1] + R[3 * r + 2] * kv[2];
        for (int r = 0; r < 3; ++r) d2[r] = R[3 * r] * d1[0] + R[3 * r + 1] * d1[1] + R[3 * r + 2] * d1[2];
        F[0] += d2[0]; F[1] += d2[1]; F[2] += d2[2];
    }
    if (has_dist) { F[0] += fdist[0]; F[1] += fdist[1]; F[2] += fdist[2]; }
    if (has_dist && force_point) {                     /* stale application point (scg_oracle_set_stale_force_point): tau_w = (p0 - p) x f, into the body frame */
        const double ar[3] = {force_point[0] - e->pos[0], force_point[1] - e->pos[1], force_point[2] - e->pos[2]};
        const double tw[3] = {ar[1] * fdist[2] - ar[2] * fdist[1], ar[2] * fdist[0] - ar[0] * fdist[2], ar[0] * fdist[1] - ar[1] * fdist[0]};
        for (int r = 0; r < 3; ++r) tau_b[r] += R[r] * tw[0] + R[3 + r] * tw[1] + R[6 + r] * tw[2];
    }
    /* body rates and Euler's equation with the gyroscopic term (btMultiBody m_useGyroTerm = true) */
    double wb[3] = {R[0] * e->omega[0] + R[3] * e->omega[1] + R[6] * e->omega[2],
                    R[1] * e->omega[0] + R[4] * e->omega[1] + R[7] * e->omega[2],
                    R[2] * e->omega[0] + R[5] * e->omega[1] + R[8] * e->omega[2]};
    double Jw[3] = {J[0] * wb[0], J[1] * wb[1], J[2] * wb[2]};
    double gyro[3] = {wb[1] * Jw[2] - wb[2] * Jw[1], wb[2] * Jw[0] - wb[0] * Jw[2], wb[0] * Jw[1] - wb[1] * Jw[0]};
    double wdot_b[3] = {(tau_b[0] - gyro[0]) / J[0], (tau_b[1] - gyro[1]) / J[1], (tau_b[2] - gyro[2]) / J[2]};
    double wdot_w[3] = {R[0] * wdot_b[0] + R[1] * wdot_b[1] + R[2] * wdot_b[2],
                        R[3] * wdot_b[0] + R[4] * wdot_b[1] + R[5] * wdot_b[2],
                        R[6] * wdot_b[0] + R[7] * wdot_b[1] + R[8] * wdot_b[2]};
    /* symplectic Euler: velocities first ... */
    for (int i = 0; i < 3; ++i) { e->vel[i] += F[i] / m * dt; e->omega[i] += wdot_w[i] * dt; }
    /* ... then positions with the NEW velocities; orientation by the exponential map (btMultiBody::stepPositionsMultiDof) */
    for (int i = 0; i < 3; ++i) e->pos[i] += e->vel[i] * dt;
    double w[3] = {e->omega[0], e->omega[1], e->omega[2]};
    double ang = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    if (ang * dt > 0.5 * (0.5 * PI_D)) ang = 0.5 * (0.5 * PI_D) / dt;   /* ANGULAR_MOTION_THRESHOLD */
    double ax[3], cw;
    if (ang < 0.001) {
        double k = 0.5 * dt - dt * dt * dt * 0.020833333333 * ang * ang;
        ax[0] = w[0] * k; ax[1] = w[1] * k; ax[2] = w[2] * k;
    } else {
        double k = sin(0.5 * ang * dt) / ang;
        ax[0] = w[0] * k; ax[1] = w[1] * k; ax[2] = w[2] * k;
    }
    cw = cos(ang * dt * 0.5);
    /* q_new = dq (x) q  with dq = (ax, cw) */
    const double* q = e->quat; double n[4];
    n[0] = cw * q[0] + ax[0] * q[3] + ax[1] * q[2] - ax[2] * q[1];
    n[1] = cw * q[1] + ax[1] * q[3] + ax[2] * q[0] - ax[0] * q[2];
    n[2] = cw * q[2] + ax[2] * q[3] + ax[0] * q[1] - ax[1] * q[0];
    n[3] = cw * q[3] - ax[0] * q[0] - ax[1] * q[1] - ax[2] * q[2];
    double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2] + n[3] * n[3]);
    for (int i = 0; i < 4; ++i) e->quat[i] = n[i] / nn;
}

/* one Bullet stepSimulation() of dt for the 2-DoF cart-pole chain (CP:552-583).
 * Pole inertia about its CoM: PyBullet recomputes it from the collision box when changeDynamics(mass=)
 * is called (CP:318-322) because the URDF is loaded without URDF_USE_INERTIA_FROM_FILE (CP:301-304):
 * I_c = m/12 * ((2 l)^2 + w^2), w = 0.05.  Pinned by tests/golden/cartpole_pybullet_transitions.npz. */
static void cartpole_substep(const scg_config* c, oenv* e, double force, const double tap[2], int has_tap, double dt) {
    const double l = e->inertial[0], M = e->inertial[1], m = e->inertial[2], g = c->gravity;
    const double Ic = m / 12.0 * ((2 * l) * (2 * l) + c->pole_width * c->pole_width);
    double th = e->q[1], thd = e->qd[1];
    double ct = cos(th), st = sin(th);
    double Qx = force, Qth = 0.0;
    if (has_tap) { Qx += tap[0]; Qth += tap[0] * l * ct - tap[1] * l * st; }
    double a11 = M + m, a12 = m * l * ct, a22 = Ic + m * l * l;
    double b1 = Qx + m * l * st * thd * thd, b2 = Qth + m * g * l * st;
    double det = a11 * a22 - a12 * a12;
    double xdd = (a22 * b1 - a12 * b2) / det, thdd = (a11 * b2 - a12 * b1) / det;
    e->qd[0] += xdd * dt; e->qd[1] += thdd * dt;
    e->q[0] += e->qd[0] * dt; e->q[1] += e->qd[1] * dt;
}

/* ------------------------------------------------------------------------------------------ */
/* env logic                                                                                    */
/* ------------------------------------------------------------------------------------------ */
static const double* goal_row(const scg_oracle* o, int idx) {
    if (o->cfg.task == SCG_TASK_STABILIZATION) return o->x_goal;
    if (idx > o->L - 1) idx = o->L - 1;
    return o->x_goal + (size_t)idx * o->nx;
}

/* noise source: column `slot` of this env's row for the current control step */
typedef struct { const scg_oracle* o; int env; const double* row; uint32_t stream; uint32_t idx; int slot; } noise_src;
static double next_u01(noise_src* s) {
    double v;
    if (s->o->cfg.rng_mode == SCG_RNG_TABLE) v = s->row ? s->row[s->slot] : 0.0;
    else v = draw_uniform(s->o->cfg.seed, s->stream, (uint32_t)(s->o->cfg.env_index_offset + s->env), s->idx, (uint32_t)s->slot);
    s->slot++; return v;
}
static double next_normal(noise_src* s) {
    double v;
    if (s->o->cfg.rng_mode == SCG_RNG_TABLE) v = s->row ? s->row[s->slot] : 0.0;
    else v = draw_normal(s->o->cfg.seed, s->stream, (uint32_t)(s->o->cfg.env_index_offset + s->env), s->idx, (uint32_t)s->slot);
    s->slot++; return v;
}

/* DisturbanceList.apply (DI:54-62) for one mode */
static void apply_disturbances(const scg_oracle* o, const oenv* e, int mode, double* target, int tdim, noise_src* ns) {
    const scg_config* c = &o->cfg;
    for (int j = 0; j < c->n_dist[mode]; ++j) {
        const scg_disturbance* d = &c->dist[mode][j];
        double noise[SCG_MAX_DIST_DIM];
        switch (d->kind) {
            case SCG_DIST_IMPULSE: { /* DI:108-123 */
                double nz = 0;
                int off = d->step_offset >= 0 ? d->step_offset : e->dist_offset[mode][j];
                int peak = (int)(off + d->duration / 2);
                if (e->ctrl_step >= off) {
                    int po = abs(e->ctrl_step - peak);
                    double decay = (po < d->duration / 2) ? pow(d->decay_rate, po) : 0.0;
                    nz = d->magnitude * decay;
                }
                for (int k = 0; k < d->dim; ++k) noise[k] = nz;
            } break;
            case SCG_DIST_STEP: { /* DI:153-163 */
                int off = d->step_offset >= 0 ? d->step_offset : e->dist_offset[mode][j];
                double nz = e->ctrl_step >= off ? d->magnitude : 0.0;
                for (int k = 0; k < d->dim; ++k) noise[k] = nz;
            } break;
            case SCG_DIST_UNIFORM: /* DI:187-192: low + (high-low)*U[0,1) */
                for (int k = 0; k < d->dim; ++k) noise[k] = d->low[k] + (d->high[k] - d->low[k]) * next_u01(ns);
                break;
            case SCG_DIST_WHITE: /* DI:215-223 */
                for (int k = 0; k < d->dim; ++k) noise[k] = d->std[k] * next_normal(ns);
                break;
            case SCG_DIST_PERIODIC: { /* DI:249-259: random phase each call, t = pyb_step_counter * dt_pyb ; mask dropped (DI:244) */
                double t = e->pyb_step * (1.0 / c->pyb_freq);
                for (int k = 0; k < d->dim; ++k) {
                    double phase = -PI_D + 2.0 * PI_D * next_u01(ns);
                    noise[k] = d->scale * sin(2 * PI_D * d->frequency * t + phase);
                }
            } break;
            default: for (int k = 0; k < d->dim; ++k) noise[k] = 0;
        }
        for (int k = 0; k < tdim && k < d->dim; ++k) {
            double nz = noise[k];
            if (d->has_mask && d->kind != SCG_DIST_PERIODIC) nz *= d->mask[k];
            target[k] += nz;
        }
    }
}

/* state vector from the rigid body / joints (QD:784-802, CP:255-257) */
static void extract_state(const scg_oracle* o, oenv* e) {
    switch (o->cfg.system) {
        case SCG_SYS_CARTPOLE:
            e->state[0] = e->q[0]; e->state[1] = e->qd[0]; e->state[2] = e->q[1]; e->state[3] = e->qd[1]; break;
        case SCG_SYS_QUAD1D:
            e->state[0] = e->pos[2]; e->state[1] = e->vel[2]; break;
        case SCG_SYS_QUAD2D:
            rpy_from_quat(e->quat, e->rpy);
            e->state[0] = e->pos[0]; e->state[1] = e->vel[0]; e->state[2] = e->pos[2]; e->state[3] = e->vel[2];
            e->state[4] = e->rpy[1]; e->state[5] = e->omega[1]; break;
        default: {
            rpy_from_quat(e->quat, e->rpy);
            double R[9]; rot_from_quat(e->quat, R);
            e->state[0] = e->pos[0]; e->state[1] = e->vel[0]; e->state[2] = e->pos[1]; e->state[3] = e->vel[1];
            e->state[4] = e->pos[2]; e->state[5] = e->vel[2];
            e->state[6] = e->rpy[0]; e->state[7] = e->rpy[1]; e->state[8] = e->rpy[2];
            for (int i = 0; i < 3; ++i) e->state[9 + i] = R[0 + i] * e->omega[0] + R[3 + i] * e->omega[1] + R[6 + i] * e->omega[2];
        }
    }
}

/* _get_observation (QD:777-817, CP:585-609) + extend_obs (BE:422-445) */
static void make_obs(const scg_oracle* o, oenv* e, int at_reset, noise_src* ns, double* obs) {
    const scg_config* c = &o->cfg; const int nx = o->nx;
    for (int i = 0; i < nx; ++i) obs[i] = e->state[i];
    if (c->n_dist[SCG_MODE_OBSERVATION] > 0) apply_disturbances(o, e, SCG_MODE_OBSERVATION, obs, nx, ns);
    if (c->system == SCG_SYS_CARTPOLE && c->obs_wrap_angle) obs[2] = normalize_angle(obs[2]);
    if (c->cost == SCG_COST_RL_REWARD && c->obs_goal_horizon > 0) {
        if (c->task == SCG_TASK_TRAJ_TRACKING) {
            int next_step = at_reset ? 1 : e->ctrl_step + 2;
            for (int h = 0; h < c->obs_goal_horizon; ++h) {
                const double* g = goal_row(o, next_step + h);
                for (int i = 0; i < nx; ++i) obs[nx * (1 + h) + i] = g[i];
            }
        } else {
            for (int i = 0; i < nx; ++i) obs[nx + i] = o->x_goal[i];
        }
    }
}

/* reset() of one env (QD:328-392, CP:266-352, BE:320-359) */
static void env_reset(scg_oracle* o, int i) {
    const scg_config* c = &o->cfg; oenv* e = &o->envs[i]; const int nx = o->nx;
    const uint32_t genv = (uint32_t)(c->env_index_offset + i);
    const int k = e->episode;
    const double* row = (c->rng_mode == SCG_RNG_TABLE && o->ep_table) ? o->ep_table + ((size_t)(k % o->K) * c->num_envs + i) * o->P : NULL;
    e->initial_reset = 1; e->ctrl_step = 0; e->pyb_step = 0;
    e->adv_set = 0;                                                               /* BE:338-339: self.adv_action = None */
    /* disturbance reset draws (BE:336-337, DI:102-106,148-151) */
    int slot = 0;
    for (int m = 0; m < 3; ++m)
        for (int j = 0; j < c->n_dist[m]; ++j) {
            const scg_disturbance* d = &c->dist[m][j];
            e->dist_offset[m][j] = d->step_offset;
            if ((d->kind == SCG_DIST_IMPULSE || d->kind == SCG_DIST_STEP) && d->step_offset < 0) {
                double v = row ? row[slot] : (c->rng_mode == SCG_RNG_PHILOX ? draw_randint(c->seed, STREAM_EPISODE, genv, (uint32_t)k, (uint32_t)slot, (uint32_t)c->ctrl_steps) : 0.0);
                e->dist_offset[m][j] = (int)v; slot++;
            }
        }
    /* inertial properties, additive randomisation (BE:267, QD:346-358, CP:287-294) */
    for (int j = 0; j < SCG_MAX_INERTIAL; ++j) {
        double v = c->inertial_nominal[j];
        if (c->randomized_inertial && j < o->n_inertial) {
            if (row) v += row[slot + j];
            else if (c->rng_mode == SCG_RNG_PHILOX) {
                if (c->inertial_rand_kind[j] == SCG_RAND_UNIFORM)
                    v += c->inertial_rand_a[j] + (c->inertial_rand_b[j] - c->inertial_rand_a[j]) * draw_uniform(c->seed, STREAM_EPISODE, genv, (uint32_t)k, (uint32_t)(slot + j));
                else if (c->inertial_rand_kind[j] == SCG_RAND_NORMAL)
                    v += c->inertial_rand_a[j] + c->inertial_rand_b[j] * draw_normal(c->seed, STREAM_EPISODE, genv, (uint32_t)k, (uint32_t)(slot + j));
            }
        }
        e->inertial[j] = v;
    }
    slot += SCG_MAX_INERTIAL;
    /* initial state, additive randomisation (QD:369-372, CP:324-326) */
    double init[SCG_MAX_NX];
    for (int j = 0; j < nx; ++j) {
        double v = c->init_nominal[j];
        if (c->randomized_init) {
            if (row) v += row[slot + j];
            else if (c->rng_mode == SCG_RNG_PHILOX) {
                if (c->init_rand_kind[j] == SCG_RAND_UNIFORM)
                    v += c->init_rand_a[j] + (c->init_rand_b[j] - c->init_rand_a[j]) * draw_uniform(c->seed, STREAM_EPISODE, genv, (uint32_t)k, (uint32_t)(slot + j));
                else if (c->init_rand_kind[j] == SCG_RAND_NORMAL)
                    v += c->init_rand_a[j] + c->init_rand_b[j] * draw_normal(c->seed, STREAM_EPISODE, genv, (uint32_t)k, (uint32_t)(slot + j));
            }
        }
        init[j] = v;
    }
    e->episode = k + 1;
    if (c->system == SCG_SYS_CARTPOLE) {
        e->q[0] = init[0]; e->qd[0] = init[1]; e->q[1] = init[2]; e->qd[1] = init[3];
    } else {
        double rpy[3] = {0, 0, 0};
        for (int j = 0; j < 3; ++j) { e->pos[j] = 0; e->vel[j] = 0; e->omega[j] = 0; }
        if (c->system == SCG_SYS_QUAD1D) { e->pos[0] = init[0]; e->vel[0] = init[1]; }   /* quirk: the 1D labels are init_x, init_x_dot (QD:210) and QD:373-374 route them to x, z stays 0 */
        else if (c->system == SCG_SYS_QUAD2D) {
            e->pos[0] = init[0]; e->vel[0] = init[1]; e->pos[2] = init[2]; e->vel[2] = init[3];
            rpy[1] = init[4]; e->omega[1] = init[5];                                       /* QD:376-377 */
        } else {
            e->pos[0] = init[0]; e->vel[0] = init[1]; e->pos[1] = init[2]; e->vel[1] = init[3]; e->pos[2] = init[4]; e->vel[2] = init[5];
            rpy[0] = init[6]; rpy[1] = init[7]; rpy[2] = init[8];
            e->omega[0] = init[9]; e->omega[1] = init[10]; e->omega[2] = init[11];         /* QD:379: body rates passed as WORLD angular velocity */
        }
        quat_from_rpy(rpy, e->quat);
    }
    extract_state(o, e);
    /* reset-time observation noise (QD:806-807 at reset): Philox stream RESET keyed by the episode; not tabulated in table mode */
    noise_src ns = {o, i, NULL, STREAM_RESET, (uint32_t)k, 0};
    make_obs(o, e, 1, &ns, e->obs);
}

/* constraint values g (CO:97-109, 274, 445-447) ; z = [state ; current_noisy_physical_action] (CO:166-171) */
static int eval_constraints(const scg_oracle* o, const oenv* e, double* g_out) {
    int violated = 0; const int nx = o->nx;
    int rows = o->n_con;
    for (int r = 0; r < rows; ++r) {
        double acc = 0;
        for (int p = o->row_ptr[r]; p < o->row_ptr[r + 1]; ++p) {
            if (o->row_kind[r] == SCG_CON_QUAD) {      /* x'F'PFx (CO:222): the non-zero carries the pair of z indices */
                int c0 = o->col[p] & 0xff, c1 = o->col[p] >> 8;
                double z0 = c0 < nx ? e->state[c0] : e->noisy_act[c0 - nx], z1 = c1 < nx ? e->state[c1] : e->noisy_act[c1 - nx];
                acc += o->val[p] * z0 * z1;
                continue;
            }
            int cidx = o->col[p];
            double z = cidx < nx ? e->state[cidx] : e->noisy_act[cidx - nx];
            acc += o->val[p] * z;
        }
        if (o->row_kind[r] == SCG_CON_ABS) acc = fabs(acc);
        double g = round8(acc - o->b[r]);
        if (g_out) g_out[r] = g;
        if (o->row_strict[r] ? (g >= 0.0) : (g > 0.0)) violated = 1;   /* CO:126-129 */
    }
    return violated;
}

typedef struct {
    double rew; int done; int truncated; int cviol; double mse; int out_of_bounds; int goal_reached; int current_step;
} step_info;

/* step() of one env WITHOUT auto-reset (QD:394-445, CP:238-264, BE:400-502) */
static void env_step(scg_oracle* o, int i, const double* action, const double* noise_row, double* obs_out, step_info* info, double* cvals) {
    const scg_config* c = &o->cfg; oenv* e = &o->envs[i]; const int nx = o->nx, nu = o->nu;
    noise_src ns = {o, i, noise_row, STREAM_STEP, e->total_steps, 0};
    const int steps = c->pyb_freq / c->ctrl_freq;
    const double dt = 1.0 / c->pyb_freq;
    double u[SCG_MAX_NU];
    if (c->system == SCG_SYS_CARTPOLE) {
        /* CP:479-502 */
        for (int k = 0; k < nu; ++k) u[k] = c->normalized_action ? c->hover_thrust * action[k] : action[k];
        apply_disturbances(o, e, SCG_MODE_ACTION, u, nu, &ns);
        if (c->adversary_mode == SCG_ADV_ACTION && e->adv_set) for (int k = 0; k < nu; ++k) u[k] += e->adv_action[k];   /* CP:494-495 (not cleared) */
        for (int k = 0; k < nu; ++k) {
            e->noisy_act[k] = u[k];
            double f = u[k]; if (f < c->act_low[k]) f = c->act_low[k]; if (f > c->act_high[k]) f = c->act_high[k];
            e->clipped_act[k] = f;
        }
        double tap[2] = {0, 0}; int has_tap = c->n_dist[SCG_MODE_DYNAMICS] > 0;
        if (has_tap) apply_disturbances(o, e, SCG_MODE_DYNAMICS, tap, 2, &ns);
        if (c->adversary_mode == SCG_ADV_DYNAMICS) {                              /* CP:543-551: added, then cleared */
            has_tap = 1;
            if (e->adv_set) { tap[0] += e->adv_action[0]; tap[1] += e->adv_action[1]; e->adv_set = 0; }
        }
        for (int s = 0; s < steps; ++s) cartpole_substep(c, e, e->clipped_act[0], tap, has_tap, dt);
    } else {
        /* QD:722-747 */
        for (int k = 0; k < nu; ++k) u[k] = c->normalized_action ? (1 + c->norm_act_scale * action[k]) * c->hover_thrust : action[k];
        apply_disturbances(o, e, SCG_MODE_ACTION, u, nu, &ns);
        if (c->adversary_mode == SCG_ADV_ACTION && e->adv_set) for (int k = 0; k < nu; ++k) u[k] += e->adv_action[k];   /* QD:737-738 */
        double thrust[SCG_MAX_NU];
        for (int k = 0; k < nu; ++k) {
            e->noisy_act[k] = u[k];
            double t = u[k]; if (t < c->act_low[k]) t = c->act_low[k]; if (t > c->act_high[k]) t = c->act_high[k];
            e->clipped_act[k] = t; thrust[k] = t;
        }
        /* cmd2pwm / pwm2rpm (QU:16-60) */
        int n_motor = 4 / nu; double pwm[4], rpm[4], mp[SCG_MAX_NU];
        for (int k = 0; k < nu; ++k) {
            double t = thrust[k] < 0 ? 0 : thrust[k];
            mp[k] = (sqrt(t / n_motor / c->kf) - c->pwm2rpm_const) / c->pwm2rpm_scale;
        }
        if (nu == 1) { pwm[0] = pwm[1] = pwm[2] = pwm[3] = mp[0]; }
        else if (nu == 2) { pwm[0] = mp[0]; pwm[1] = mp[1]; pwm[2] = mp[1]; pwm[3] = mp[0]; }   /* QU:39 */
        else { for (int k = 0; k < 4; ++k) pwm[k] = mp[k]; }
        for (int k = 0; k < 4; ++k) {
            if (pwm[k] < c->pwm_min) pwm[k] = c->pwm_min; if (pwm[k] > c->pwm_max) pwm[k] = c->pwm_max;
            rpm[k] = c->pwm2rpm_scale * pwm[k] + c->pwm2rpm_const;
        }
        /* dynamics disturbance force (QD:413-435) */
        double fd[3] = {0, 0, 0}; int has_dist = c->n_dist[SCG_MODE_DYNAMICS] > 0;
        if (has_dist) {
            double t[3] = {0, 0, 0};
            int ddim = c->system == SCG_SYS_QUAD1D ? 1 : c->system == SCG_SYS_QUAD2D ? 2 : 3;
            apply_disturbances(o, e, SCG_MODE_DYNAMICS, t, ddim, &ns);
            if (ddim == 1) { fd[2] = t[0]; } else if (ddim == 2) { fd[0] = t[0]; fd[2] = t[1]; } else { fd[0] = t[0]; fd[1] = t[1]; fd[2] = t[2]; }
        }
        if (c->adversary_mode == SCG_ADV_DYNAMICS) {                              /* QD:416-425: added to the force, then cleared */
            has_dist = 1;
            if (e->adv_set) {
                const double* av = e->adv_action;
                if (c->system == SCG_SYS_QUAD1D) fd[2] += av[0];
                else if (c->system == SCG_SYS_QUAD2D) { fd[0] += av[0]; fd[2] += av[1]; }
                else { fd[0] += av[0]; fd[1] += av[1]; fd[2] += av[2]; }
                e->adv_set = 0;
            }
        }
        const double pos0[3] = {e->pos[0], e->pos[1], e->pos[2]};
        const double* fp = (o->stale_force_point && c->physics == SCG_PHYS_PYB && steps > 1) ? pos0 : NULL;
        for (int s = 0; s < steps; ++s) quad_substep(c, e, rpm, fd, has_dist, dt, fp);
    }
    extract_state(o, e);
    make_obs(o, e, 0, &ns, obs_out);

    /* reward (QD:819-862, CP:611-652) */
    double rew;
    {
        double st[SCG_MAX_NX]; for (int k = 0; k < nx; ++k) st[k] = e->state[k];
        if (c->cost == SCG_COST_RL_REWARD) {
            if (c->system == SCG_SYS_CARTPOLE) st[2] = normalize_angle(st[2]);   /* CP:620 */
            const double* g = goal_row(o, e->ctrl_step + 1);                     /* QD:837 (clamped) */
            double dist = 0;
            for (int k = 0; k < nx; ++k) { double er = st[k] - g[k]; dist += c->rew_state_weight[k] * er * er; }
            for (int k = 0; k < nu; ++k) {
                double ae = (c->system == SCG_SYS_CARTPOLE) ? e->noisy_act[k] : (e->noisy_act[k] - c->u_goal[k]);  /* CP:625 vs QD:829 */
                dist += c->rew_act_weight[k] * ae * ae;
            }
            rew = -dist; if (c->rew_exponential) rew = exp(rew);
        } else {
            /* symbolic.loss: 0.5 e'Qe + 0.5 (u-Ur)'R(u-Ur) with the clipped action (QD:848-862, CP:636-652) */
            int idx = (c->system == SCG_SYS_CARTPOLE) ? e->ctrl_step : e->ctrl_step + 1;   /* CP:648 vs QD:858, unclamped */
            const double* g = (c->task == SCG_TASK_STABILIZATION) ? o->x_goal : o->x_goal + (size_t)(idx < o->L ? idx : o->L - 1) * nx;
            double l = 0;
            for (int k = 0; k < nx; ++k) { double er = st[k] - g[k]; l += 0.5 * c->Q[k] * er * er; }
            for (int k = 0; k < nu; ++k) { double ae = e->clipped_act[k] - c->u_goal[k]; l += 0.5 * c->R[k] * ae * ae; }
            rew = -l;
        }
    }
    /* done (QD:864-894, CP:654-672) */
    int done = 0, goal_reached = 0, oob = 0;
    if (c->task == SCG_TASK_STABILIZATION) {
        double n2 = 0; for (int k = 0; k < nx; ++k) { double er = e->state[k] - o->x_goal[k]; n2 += er * er; }
        goal_reached = sqrt(n2) < c->goal_tolerance;
        if (goal_reached) done = 1;
    }
    if (!done && c->done_on_out_of_bound) {
        if (c->system == SCG_SYS_CARTPOLE) {
            double x = e->state[0], th = e->state[2];
            if (x < -c->x_threshold || x > c->x_threshold || th < -c->theta_threshold || th > c->theta_threshold) oob = 1;
        } else {
            static const int m1[2] = {1, 0}, m2[6] = {1, 0, 1, 0, 1, 0}, m3[12] = {1, 0, 1, 0, 1, 0, 1, 1, 1, 0, 0, 0};
            const int* mk = c->system == SCG_SYS_QUAD1D ? m1 : c->system == SCG_SYS_QUAD2D ? m2 : m3;
            for (int k = 0; k < nx; ++k) if (mk[k] && (e->state[k] < c->state_low[k] || e->state[k] > c->state_high[k])) oob = 1;
        }
        if (oob) done = 1;
    }
    /* info['mse'] (QD:907-922, CP:685-695) */
    double mse = 0;
    {
        double st[SCG_MAX_NX]; for (int k = 0; k < nx; ++k) st[k] = e->state[k];
        const double* g;
        if (c->task == SCG_TASK_STABILIZATION) g = o->x_goal;
        else {
            if (c->system == SCG_SYS_CARTPOLE) st[2] = normalize_angle(st[2]);
            else if (c->system == SCG_SYS_QUAD2D) st[4] = normalize_angle(st[4]);
            else if (c->system == SCG_SYS_QUAD3D) { st[6] = normalize_angle(st[6]); st[7] = normalize_angle(st[7]); st[8] = normalize_angle(st[8]); }
            g = goal_row(o, e->ctrl_step + 1);
        }
        for (int k = 0; k < nx; ++k) { double er = (st[k] - g[k]) * c->mse_weight[k]; mse += er * er; }
    }
    /* after_step (BE:447-502) */
    e->pyb_step += steps; e->ctrl_step += 1; e->total_steps += 1;
    int cviol = 0;
    if (o->n_con > 0) {
        cviol = eval_constraints(o, e, cvals);
        if (cviol && c->done_on_violation) { done = 1; if (c->cost == SCG_COST_RL_REWARD && c->use_constraint_penalty) rew = 0; }
    }
    if (c->cost == SCG_COST_RL_REWARD && o->n_con > 0 && c->use_constraint_penalty && cviol) {
        if (c->rew_exponential) { rew = log(rew); rew -= c->constraint_penalty; rew = exp(rew); }
        else rew -= c->constraint_penalty;
    }
    int truncated = 0;
    if (e->ctrl_step >= c->ctrl_steps) { truncated = !done; done = 1; }
    info->rew = rew; info->done = done; info->truncated = truncated; info->cviol = cviol; info->mse = mse;
    info->out_of_bounds = oob; info->goal_reached = goal_reached; info->current_step = e->ctrl_step;
}

/* ------------------------------------------------------------------------------------------ */
/* public batch API                                                                             */
/* ------------------------------------------------------------------------------------------ */
void scg_oracle_reset(scg_oracle* o, const uint8_t* mask, double* obs /*[N][obs_dim]*/, double* state /*[N][nx]*/) {
    const int N = o->cfg.num_envs;
    for (int i = 0; i < N; ++i) {
        if (mask && !mask[i]) continue;
        env_reset(o, i);
        o->envs[i].ep_return = 0; o->envs[i].ep_length = 0; o->envs[i].ep_cviol = 0;
    }
    for (int i = 0; i < N; ++i) {
        if (obs) memcpy(obs + (size_t)i * o->obs_dim, o->envs[i].obs, sizeof(double) * o->obs_dim);
        if (state) memcpy(state + (size_t)i * o->nx, o->envs[i].state, sizeof(double) * o->nx);
    }
}

typedef struct {
    double* obs; double* state; double* rew; uint8_t* done; uint8_t* truncated; double* cviol; double* cvals; double* mse;
    uint8_t* out_of_bounds; uint8_t* goal_reached; int32_t* current_step; double* terminal_obs; double* terminal_state;
} oracle_step_out;

/* returns SCG_ESTATE if an env was never reset (BE:230-235) */
int scg_oracle_step(scg_oracle* o, const double* act /*[N][nu]*/, const oracle_step_out* out) {
    const int N = o->cfg.num_envs, nx = o->nx, nu = o->nu, od = o->obs_dim, nc = o->n_con;
    for (int i = 0; i < N; ++i) if (!o->envs[i].initial_reset) return SCG_ESTATE;
    const double* nrow0 = (o->cfg.rng_mode == SCG_RNG_TABLE && o->noise_table && o->D > 0) ? o->noise_table + (size_t)(o->noise_cursor % o->Tn) * N * o->D : NULL;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        double obs[SCG_MAX_OBS], cv[SCG_MAX_CON_ROWS]; step_info inf;
        env_step(o, i, act + (size_t)i * nu, nrow0 ? nrow0 + (size_t)i * o->D : NULL, obs, &inf, cv);
        oenv* e = &o->envs[i];
        if (out->terminal_obs) memcpy(out->terminal_obs + (size_t)i * od, obs, sizeof(double) * od);
        if (out->terminal_state) memcpy(out->terminal_state + (size_t)i * nx, e->state, sizeof(double) * nx);
        if (inf.done && o->cfg.auto_reset) { env_reset(o, i); memcpy(obs, e->obs, sizeof(double) * od); }   /* VE:33-38 */
        else memcpy(e->obs, obs, sizeof(double) * od);
        if (out->obs) memcpy(out->obs + (size_t)i * od, obs, sizeof(double) * od);
        if (out->state) memcpy(out->state + (size_t)i * nx, e->state, sizeof(double) * nx);
        if (out->rew) out->rew[i] = inf.rew;
        if (out->done) out->done[i] = (uint8_t)inf.done;
        if (out->truncated) out->truncated[i] = (uint8_t)inf.truncated;
        if (out->cviol) out->cviol[i] = inf.cviol;
        if (out->cvals && nc > 0) memcpy(out->cvals + (size_t)i * nc, cv, sizeof(double) * nc);
        if (out->mse) out->mse[i] = inf.mse;
        if (out->out_of_bounds) out->out_of_bounds[i] = (uint8_t)inf.out_of_bounds;
        if (out->goal_reached) out->goal_reached[i] = (uint8_t)inf.goal_reached;
        if (out->current_step) out->current_step[i] = inf.current_step;
    }
    o->noise_cursor += 1;
    return SCG_OK;
}


/* BenchmarkEnv.set_adversary_control (BE:216-228); adversary dim = dim of the disturbance mode it drives (BE:289-295) */
int scg_oracle_adversary_dim(const scg_oracle* o) {
    const scg_config* c = &o->cfg;
    if (c->adversary_mode == SCG_ADV_ACTION) return o->nu;
    if (c->adversary_mode == SCG_ADV_DYNAMICS) return c->system == SCG_SYS_CARTPOLE ? 2 : c->system == SCG_SYS_QUAD1D ? 1 : c->system == SCG_SYS_QUAD2D ? 2 : 3;
    return 0;
}
static void set_adversary(const scg_oracle* o, oenv* e, const double* a) {
    const scg_config* c = &o->cfg; const int d = scg_oracle_adversary_dim(o);
    for (int k = 0; k < d; ++k) {
        /* the adversary's action is a float32 array (torch -> numpy); `clipped * scale + offset` with python-float scale / offset stays
         * float32 (NEP 50 weak scalars), and is promoted to float64 when it is added to the force / action */
        float v = (float)a[k]; v = v < -1.0f ? -1.0f : v > 1.0f ? 1.0f : v;      /* adversary_action_space = Box(-1, 1) */
        v = v * (float)c->adversary_scale; v = v + (float)c->adversary_offset;
        e->adv_action[k] = (double)v;
    }
    e->adv_set = 1;
}
int scg_oracle_set_adversary_control(scg_oracle* o, const double* adv /*[N][dim]*/) {
    const int d = scg_oracle_adversary_dim(o);
    if (d == 0) return SCG_ESTATE;                                               /* RuntimeError: adversary_disturbance does not exist */
    for (int i = 0; i < o->cfg.num_envs; ++i) set_adversary(o, &o->envs[i], adv + (size_t)i * d);
    return SCG_OK;
}

/* direct state injection for unit tests of the physics (bypasses reset randomisation) */
void scg_oracle_set_env_state(scg_oracle* o, int i, const double* state, const double* inertial) {
    oenv* e = &o->envs[i]; const scg_config* c = &o->cfg;
    for (int j = 0; j < SCG_MAX_INERTIAL; ++j) e->inertial[j] = inertial ? inertial[j] : c->inertial_nominal[j];
    e->initial_reset = 1;
    if (c->system == SCG_SYS_CARTPOLE) { e->q[0] = state[0]; e->qd[0] = state[1]; e->q[1] = state[2]; e->qd[1] = state[3]; }
    else {
        double rpy[3] = {0, 0, 0};
        for (int j = 0; j < 3; ++j) { e->pos[j] = 0; e->vel[j] = 0; e->omega[j] = 0; }
        if (c->system == SCG_SYS_QUAD1D) { e->pos[2] = state[0]; e->vel[2] = state[1]; }
        else if (c->system == SCG_SYS_QUAD2D) { e->pos[0] = state[0]; e->vel[0] = state[1]; e->pos[2] = state[2]; e->vel[2] = state[3]; rpy[1] = state[4]; e->omega[1] = state[5]; }
        else {
            e->pos[0] = state[0]; e->vel[0] = state[1]; e->pos[1] = state[2]; e->vel[1] = state[3]; e->pos[2] = state[4]; e->vel[2] = state[5];
            rpy[0] = state[6]; rpy[1] = state[7]; rpy[2] = state[8]; e->omega[0] = state[9]; e->omega[1] = state[10]; e->omega[2] = state[11];
        }
        quat_from_rpy(rpy, e->quat);
    }
    extract_state(o, e);
    noise_src ns = {o, i, NULL, STREAM_RESET, (uint32_t)e->episode, 0};
    make_obs(o, e, e->ctrl_step == 0, &ns, e->obs);
}

/* Overwrites the physical state of env i with an OBSERVED env.state (the inverse of extract_state): used by the teacher-forced
 * replay below.  Counters, episode index and the per-episode inertial properties are left alone. */
static void inject_observed_state(scg_oracle* o, int i, const double* st) {
    oenv* e = &o->envs[i]; const scg_config* c = &o->cfg;
    if (c->system == SCG_SYS_CARTPOLE) { e->q[0] = st[0]; e->qd[0] = st[1]; e->q[1] = st[2]; e->qd[1] = st[3]; }
    else if (c->system == SCG_SYS_QUAD1D) { e->pos[2] = st[0]; e->vel[2] = st[1]; }
    else if (c->system == SCG_SYS_QUAD2D) {
        double rpy[3] = {0, st[4], 0};
        e->pos[0] = st[0]; e->vel[0] = st[1]; e->pos[1] = 0; e->vel[1] = 0; e->pos[2] = st[2]; e->vel[2] = st[3];
        quat_from_rpy(rpy, e->quat); e->omega[0] = 0; e->omega[1] = st[5]; e->omega[2] = 0;
    } else {
        double rpy[3] = {st[6], st[7], st[8]}, R[9];
        e->pos[0] = st[0]; e->vel[0] = st[1]; e->pos[1] = st[2]; e->vel[1] = st[3]; e->pos[2] = st[4]; e->vel[2] = st[5];
        quat_from_rpy(rpy, e->quat); rot_from_quat(e->quat, R);
        for (int r = 0; r < 3; ++r) e->omega[r] = R[3 * r + 0] * st[9] + R[3 * r + 1] * st[10] + R[3 * r + 2] * st[11];   /* omega_world = R . body rates */
    }
    extract_state(o, e);
}

/* ------------------------------------------------------------------------------------------ */
/* policy: MLPActorCritic (PU:149-238, neural_networks.py:18-54, distributions.py:9-33), float32 */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int obs_dim, act_dim, hidden, activation;
    const float *aw[3], *ab[3], *cw[3], *cb[3], *logstd;   /* nn.Linear layout [out][in] */
} oracle_policy;

static float act_fn(float x, int kind) {
    if (kind == SCG_ACT_TANH) return tanhf(x);
    if (kind == SCG_ACT_RELU) return x > 0 ? x : 0;
    return x > 0 ? x : 0.01f * x;
}
static void mlp_forward(const float* const w[3], const float* const b[3], int in_dim, int hid, int out_dim, int actv, const float* x, float* y) {
    float h1[256], h2[256];
    for (int j = 0; j < hid; ++j) { float s = b[0][j]; for (int k = 0; k < in_dim; ++k) s += w[0][(size_t)j * in_dim + k] * x[k]; h1[j] = act_fn(s, actv); }
    for (int j = 0; j < hid; ++j) { float s = b[1][j]; for (int k = 0; k < hid; ++k) s += w[1][(size_t)j * hid + k] * h1[k]; h2[j] = act_fn(s, actv); }
    for (int j = 0; j < out_dim; ++j) { float s = b[2][j]; for (int k = 0; k < hid; ++k) s += w[2][(size_t)j * hid + k] * h2[k]; y[j] = s; }
}
void scg_oracle_policy_forward(const oracle_policy* p, const float* obs, int n, float* mean /*[n][A]*/, float* value /*[n]*/) {
    for (int i = 0; i < n; ++i) {
        mlp_forward(p->aw, p->ab, p->obs_dim, p->hidden, p->act_dim, p->activation, obs + (size_t)i * p->obs_dim, mean + (size_t)i * p->act_dim);
        mlp_forward(p->cw, p->cb, p->obs_dim, p->hidden, 1, p->activation, obs + (size_t)i * p->obs_dim, value + i);
    }
}

typedef struct {
    float *obs, *act, *rew, *mask, *v, *logp, *terminal_v, *ret, *adv, *cviol, *cvals, *mse, *last_obs, *last_val;
    float *ep_return_sum, *ep_length_sum, *ep_cviol_sum; int32_t* ep_count;
    double* state_trace;   /* [T][N][nx] env.state after each step before auto-reset (debug) */
    /* TEACHER-FORCED REPLAY of somebody else's record (the CUDA collector's): before every step the env's physical state is
     * overwritten with the state part of replay_obs[t] and the env is stepped with replay_act[t]; an episode ends when
     * replay_mask[t] says so.  Every output of this struct is then the oracle's answer to the RECORD's own inputs of that step --
     * one-step predictions, free of the closed-loop amplification of float32 rounding differences.  `act` holds mean + sd * eps
     * (what the sampler should have produced for the recorded observation); `next_obs` the observation after the step
     * (terminal observation when done; compare with replay_obs[t + 1] where the record did not reset). */
    const float *replay_obs, *replay_act, *replay_mask;
    float* next_obs;       /* [T][N][obs_dim] */
    /* frozen observation / reward normalisers (normalization.py:88-159 as applied in PPO:270-271; layout of scg_rollout_args.d_norm) */
    const double* norm;
    const float* obs0;     /* [N][obs_dim] policy input of step 0 (nullable) */
    float *obs_raw, *rew_raw;
    /* SafeExplorer safety layer (safe_explorer_utils.py:15-176; layout of scg_rollout_args.d_safety) */
    int safety_C, safety_H;
    const float* safety;
    const double* safety_slack;
    float* c;              /* [T][N][safety_C] constraint values fed to the safety layer */
    /* RARL / RAP adversary (rarl.py:349-418): second actor-critic on the same observation; adv_record: the record is the adversary's */
    const oracle_policy* adv_policy;
    const float* eps_adv;  /* [T][N][adv_dim] or NULL -> Philox stream ADVERSARY */
    int adv_record;
} oracle_rollout_out;

/* PPO.train_step collector loop (PPO:259-297) + compute_returns_and_advantages (PU:374-400) */
int scg_oracle_rollout(scg_oracle* o, const oracle_policy* p, const float* eps /*[T][N][A] or NULL*/, int T, int deterministic,
                       double gamma, double lam, int use_gae, const oracle_rollout_out* out) {
    const scg_config* c = &o->cfg; const int N = c->num_envs, nx = o->nx, A = o->nu, od = o->obs_dim, nc = o->n_con;
    if (p->obs_dim != od || p->act_dim != A) return SCG_EINVAL;
    for (int i = 0; i < N; ++i) if (!o->envs[i].initial_reset) return SCG_ESTATE;
    if (out->ep_count) for (int i = 0; i < N; ++i) { out->ep_count[i] = 0; out->ep_return_sum[i] = 0; out->ep_length_sum[i] = 0; out->ep_cviol_sum[i] = 0; }
    const float log_sqrt_2pi = (float)log(sqrt(2 * PI_D));
    const double* nm = out->norm;
    const int norm_obs = nm && nm[2 * od + 1] != 0.0, norm_rew = nm && nm[2 * od + 4] != 0.0;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        oenv* e = &o->envs[i];
        const uint32_t genv = (uint32_t)(c->env_index_offset + i);
        if (out->obs0) for (int k = 0; k < od; ++k) e->obs[k] = (double)out->obs0[(size_t)i * od + k];
        for (int t = 0; t < T; ++t) {
            size_t tn = (size_t)t * N + i;
            float obs_f[SCG_MAX_OBS], mean[SCG_MAX_NU], val, act[SCG_MAX_NU];
            const int replay = out->replay_obs != NULL;
            if (replay) {
                double st[SCG_MAX_NX];
                for (int k = 0; k < nx; ++k) st[k] = (double)out->replay_obs[tn * od + k];
                inject_observed_state(o, i, st);
                for (int k = 0; k < od; ++k) e->obs[k] = (double)out->replay_obs[tn * od + k];
            }
            for (int k = 0; k < od; ++k) obs_f[k] = (float)e->obs[k];        /* torch.FloatTensor(obs) PPO:268 */
            mlp_forward(p->aw, p->ab, od, p->hidden, A, p->activation, obs_f, mean);
            mlp_forward(p->cw, p->cb, od, p->hidden, 1, p->activation, obs_f, &val);
            if (out->safety_C > 0) {
                /* c = info['constraint_values'][:C] of the current state (safe_ppo.py:123,344); mu_safe = mu - lambda* g_i* (safe_explorer_utils.py:120-176) */
                double cv[SCG_MAX_CON_ROWS];
                eval_constraints(o, e, cv);
                const int C = out->safety_C, H = out->safety_H, per = H * od + H + A * H + A;
                float best = -1.0f, gbest[SCG_MAX_NU] = {0};
                for (int ci = 0; ci < C; ++ci) {
                    const float* w1 = out->safety + (size_t)ci * per; const float* b1 = w1 + H * od; const float* w2 = b1 + H; const float* b2 = w2 + A * H;
                    float g[SCG_MAX_NU];
                    for (int a = 0; a < A; ++a) g[a] = b2[a];
                    for (int j = 0; j < H; ++j) {
                        float hj = b1[j];
                        for (int k = 0; k < od; ++k) hj += w1[j * od + k] * obs_f[k];
                        hj = hj > 0 ? hj : 0;
                        for (int a = 0; a < A; ++a) g[a] += w2[a * H + j] * hj;
                    }
                    const float cf = (float)cv[ci];
                    if (out->c) out->c[tn * C + ci] = cf;
                    float num = 0, den = 0;
                    for (int a = 0; a < A; ++a) { num += g[a] * mean[a]; den += g[a] * g[a]; }
                    num = (num + cf) + (float)out->safety_slack[ci];
                    float m = num / (den + 1e-8f); m = m > 0 ? m : 0;
                    if (m > best) { best = m; for (int a = 0; a < A; ++a) gbest[a] = g[a]; }
                }
                for (int a = 0; a < A; ++a) mean[a] = mean[a] - best * gbest[a];
            }
            float logp = 0;
            for (int k = 0; k < A; ++k) {
                float sd = expf(p->logstd[k]);
                float z = 0;
                if (!deterministic) z = eps ? eps[tn * A + k] : (float)draw_normal(c->seed, STREAM_POLICY, genv, e->total_steps, (uint32_t)k);
                act[k] = mean[k] + sd * z;                                    /* Normal.sample PU:228 */
                float d = act[k] - mean[k];
                logp += -(d * d) / (2 * sd * sd) - p->logstd[k] - log_sqrt_2pi; /* distributions.py:12-21 */
            }
            /* adversary action of the same observation (rarl.py:362-367): sampled like the protagonist's, then set_adversary_control */
            const oracle_policy* q = out->adv_policy; const int A2 = q ? q->act_dim : 0;
            float act2[SCG_MAX_ADV], val2 = 0, logp2 = 0;
            if (q) {
                float mean2[SCG_MAX_ADV];
                mlp_forward(q->aw, q->ab, od, q->hidden, A2, q->activation, obs_f, mean2);
                mlp_forward(q->cw, q->cb, od, q->hidden, 1, q->activation, obs_f, &val2);
                double advd[SCG_MAX_ADV];
                for (int k = 0; k < A2; ++k) {
                    float sd = expf(q->logstd[k]), z = 0;
                    if (!deterministic) z = out->eps_adv ? out->eps_adv[tn * A2 + k] : (float)draw_normal(c->seed, STREAM_ADVERSARY, genv, e->total_steps, (uint32_t)k);
                    act2[k] = mean2[k] + sd * z;
                    float d = act2[k] - mean2[k];
                    logp2 += -(d * d) / (2 * sd * sd) - q->logstd[k] - log_sqrt_2pi;
                    advd[k] = (double)act2[k];
                }
                set_adversary(o, e, advd);
            }
            double actd[SCG_MAX_NU]; for (int k = 0; k < A; ++k) actd[k] = replay ? (double)out->replay_act[tn * A + k] : (double)act[k];
            const double* nrow = (c->rng_mode == SCG_RNG_TABLE && o->noise_table && o->D > 0)
                                 ? o->noise_table + ((size_t)((o->noise_cursor + t) % o->Tn) * N + i) * o->D : NULL;
            double nobs[SCG_MAX_OBS], cv[SCG_MAX_CON_ROWS]; step_info inf;
            env_step(o, i, actd, nrow, nobs, &inf, cv);
            if (out->state_trace) memcpy(out->state_trace + tn * nx, e->state, sizeof(double) * nx);
            if (out->next_obs) for (int k = 0; k < od; ++k) out->next_obs[tn * od + k] = (float)nobs[k];
            e->ep_return += inf.rew; e->ep_length += 1; e->ep_cviol += inf.cviol;
            float term_v = 0;
            const int own_done = inf.done;
            if (replay) inf.done = out->replay_mask[tn] == 0.f;            /* follow the record's episode boundaries */
            if (inf.done) {
                if (inf.truncated) {                                          /* PPO:274-283 */
                    float tobs[SCG_MAX_OBS]; for (int k = 0; k < od; ++k) tobs[k] = (float)nobs[k];
                    if (q && out->adv_record) mlp_forward(q->cw, q->cb, od, q->hidden, 1, q->activation, tobs, &term_v);   /* rarl.py:385-386 */
                    else mlp_forward(p->cw, p->cb, od, p->hidden, 1, p->activation, tobs, &term_v);
                }
                if (out->ep_count) { out->ep_count[i] += 1; out->ep_return_sum[i] += (float)e->ep_return; out->ep_length_sum[i] += (float)e->ep_length; out->ep_cviol_sum[i] += (float)e->ep_cviol; }
                e->ep_return = 0; e->ep_length = 0; e->ep_cviol = 0;
                env_reset(o, i);                                              /* VE:33-38 */
            } else memcpy(e->obs, nobs, sizeof(double) * od);
            /* PPO:270-271: next_obs = obs_normalizer(next_obs); rew = reward_normalizer(rew, done) -- statistics frozen for the rollout */
            if (out->obs_raw) for (int k = 0; k < od; ++k) out->obs_raw[tn * od + k] = (float)e->obs[k];
            if (norm_obs) for (int k = 0; k < od; ++k) {
                double y = ((double)(float)e->obs[k] - nm[k]) * nm[od + k];
                y = y < -nm[2 * od] ? -nm[2 * od] : y; y = y > nm[2 * od] ? nm[2 * od] : y;
                e->obs[k] = (double)(float)y;
            }
            if (out->rew_raw) out->rew_raw[tn] = (float)inf.rew;
            float rew_rec = (float)inf.rew;
            if (norm_rew) { double y = (double)rew_rec * nm[2 * od + 2]; y = y < -nm[2 * od + 3] ? -nm[2 * od + 3] : y; y = y > nm[2 * od + 3] ? nm[2 * od + 3] : y; rew_rec = (float)y; }
            for (int k = 0; k < od; ++k) out->obs[tn * od + k] = obs_f[k];
            if (q && out->adv_record) {                                       /* rarl.py:391-399: act_adv, -rew, v_adv, logp_adv */
                for (int k = 0; k < A2; ++k) out->act[tn * A2 + k] = act2[k];
                rew_rec = -rew_rec; val = val2; logp = logp2;
            } else for (int k = 0; k < A; ++k) out->act[tn * A + k] = act[k];
            out->rew[tn] = rew_rec; out->mask[tn] = (replay ? own_done : inf.done) ? 0.f : 1.f; out->v[tn] = val; out->logp[tn] = logp; out->terminal_v[tn] = term_v;
            if (out->cviol) out->cviol[tn] = (float)inf.cviol;
            if (out->cvals && nc > 0) for (int k = 0; k < nc; ++k) out->cvals[tn * nc + k] = (float)cv[k];
            if (out->mse) out->mse[tn] = (float)inf.mse;
        }
        float lobs[SCG_MAX_OBS], lval; for (int k = 0; k < od; ++k) lobs[k] = (float)e->obs[k];
        if (out->adv_policy && out->adv_record) mlp_forward(out->adv_policy->cw, out->adv_policy->cb, od, out->adv_policy->hidden, 1, out->adv_policy->activation, lobs, &lval);   /* rarl.py:407-408 */
        else mlp_forward(p->cw, p->cb, od, p->hidden, 1, p->activation, lobs, &lval);
        if (out->last_obs) for (int k = 0; k < od; ++k) out->last_obs[(size_t)i * od + k] = lobs[k];
        if (out->last_val) out->last_val[i] = lval;
        if (out->ret && out->adv) {
            /* PU:374-400.  numpy dtypes: rews/vals/masks/last_val are float32 arrays and gamma/lambda python floats, so
             * `ret` and `td_error` are float32 recursions; `adv` starts as a float64 zeros array, so the GAE recursion
             * accumulates in float64 (use_gae) or is the float32 difference ret - v (otherwise). */
            const float g32 = (float)gamma;
            float ret = lval, vnext = lval; double adv = 0;
            for (int t = T - 1; t >= 0; --t) {
                size_t tn = (size_t)t * N + i;
                float r = out->rew[tn] + g32 * out->terminal_v[tn];             /* rews += gamma*terminal_v, in place (PU:389) */
                out->rew[tn] = r;
                float m = out->mask[tn], v = out->v[tn];
                ret = r + g32 * m * ret;
                if (!use_gae) adv = (double)(ret - v);
                else { float td = r + g32 * m * vnext - v; adv = adv * lam * gamma * (double)m + (double)td; }
                out->ret[tn] = ret; out->adv[tn] = (float)adv; vnext = v;
            }
        }
    }
    o->noise_cursor += T;
    return SCG_OK;
}

int scg_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void scg_oracle_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
