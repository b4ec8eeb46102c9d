'''Analytic invariants and cross-checks of the oracle's quadrotor physics (unpinned by PyBullet data, see oracle header).'''
import numpy as np
import pytest

import helpers as H


def _quad(quad_type, **kw):
    base = dict(quad_type=quad_type, ctrl_freq=50, pyb_freq=1000, task='stabilization', randomized_init=False,
                task_info={'stabilization_goal': [0, 1] if quad_type == 2 else [0, 0, 1], 'stabilization_goal_tolerance': 0.0},
                done_on_out_of_bound=False, cost='rl_reward')
    base.update(kw)
    return H.make_env('quadrotor', **base)


@pytest.mark.parametrize('qt', [2, 3])
def test_hover_is_an_equilibrium(qt):
    '''thrust = M g  =>  zero acceleration.'''
    env = _quad(qt)
    nx, nu = env.state_dim, env.action_dim
    o = H.make_oracle(env, 1, auto_reset=False)
    s0 = np.zeros(nx); s0[2 if qt == 2 else 4] = 1.0
    o.set_env_state(0, s0)
    hover = np.full((1, nu), 0.027 * 9.8 / nu)
    for _ in range(100):
        out = o.step(hover)
    # cmd2pwm -> rpm -> kf rpm^2 round-trips the thrust to ~1e-16 relative; the residual acceleration integrates to < 1e-12
    np.testing.assert_allclose(out['state'][0], s0, atol=1e-11)


def test_free_fall_matches_symplectic_euler_closed_form():
    '''Zero thrust is clipped to the minimum motor thrust; with M such that min thrust is negligible use the closed form
    z_k = z0 - a dt^2 k(k+1)/2 of semi-implicit Euler with a = g - T_min/m.'''
    env = _quad(2)
    o = H.make_oracle(env, 1, auto_reset=False)
    o.set_env_state(0, np.array([0, 0, 1.0, 0, 0, 0.0]))
    out = o.step(np.zeros((1, 2)))
    tmin = 2 * float(np.float32(env.physical_action_bounds[0][0]))
    # thrust after the pwm round trip equals the clipped command up to rounding
    a = 9.8 - tmin / 0.027
    k, dt = 20, 1e-3
    np.testing.assert_allclose(out['state'][0][2], 1.0 - a * dt * dt * k * (k + 1) / 2, rtol=0, atol=1e-12)
    np.testing.assert_allclose(out['state'][0][3], -a * dt * k, rtol=0, atol=1e-12)


def test_planar_3d_run_equals_2d_run():
    '''A 3D quadrotor started in the x-z plane with f = [F0, F1, F1, F0] must reproduce the 2D env (quadrotor_utils.py:39).'''
    e2, e3 = _quad(2), _quad(3)
    o2, o3 = H.make_oracle(e2, 1, auto_reset=False), H.make_oracle(e3, 1, auto_reset=False)
    s2 = np.array([0.1, 0.2, 1.0, -0.1, 0.15, 0.3])
    s3 = np.zeros(12); s3[[0, 1, 4, 5, 7, 10]] = s2
    o2.set_env_state(0, s2); o3.set_env_state(0, s3)
    rng = np.random.default_rng(0)
    for _ in range(20):
        t = rng.uniform(0.1318, 0.1328, size=2)     # total thrust of motor pairs (0,3) and (1,2); pitch stays far from +-90 deg
        r2 = o2.step(t.reshape(1, 2))
        r3 = o3.step(np.array([[t[0] / 2, t[1] / 2, t[1] / 2, t[0] / 2]]))
        np.testing.assert_allclose(r3['state'][0][[0, 1, 4, 5, 7, 10]], r2['state'][0], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(r3['state'][0][[2, 3, 6, 8, 9, 11]], 0, atol=1e-13)


def test_substepped_integrator_converges_to_the_symbolic_ode():
    '''As pyb_freq -> inf the oracle converges (first order) to an RK integration of the reference's symbolic 2D model
    (quadrotor.py:506-509) evaluated with the same lever arm (0.028).'''
    from scipy.integrate import solve_ivp
    m, Iyy, g, arm = 0.027, 1.4e-5, 9.8, 0.028
    T1, T2 = 0.14, 0.135

    def f(t, x):
        return [x[1], np.sin(x[4]) * (T1 + T2) / m, x[3], np.cos(x[4]) * (T1 + T2) / m - g, x[5], arm * (T2 - T1) / Iyy]
    s0 = np.array([0, 0, 1.0, 0, 0.05, 0.0])
    ref = solve_ivp(f, (0, 0.2), s0, rtol=1e-12, atol=1e-14).y[:, -1]
    errs = []
    for pyb in (500, 1000, 2000, 4000):
        env = _quad(2, pyb_freq=pyb)
        o = H.make_oracle(env, 1, auto_reset=False)
        o.set_env_state(0, s0)
        for _ in range(10):
            out = o.step(np.array([[T1, T2]]))
        errs.append(np.abs(out['state'][0] - ref).max())
    ratios = np.asarray(errs[:-1]) / np.asarray(errs[1:])
    assert np.all(ratios > 1.8) and np.all(ratios < 2.2), (errs, ratios)     # O(dt)


def test_torque_free_rotation_conserves_angular_momentum():
    '''3D, hover thrust, initial body rates: |J w| stays constant to O(dt) (gyroscopic term present).'''
    env = _quad(3)
    o = H.make_oracle(env, 1, auto_reset=False)
    s = np.zeros(12); s[4] = 1.0; s[9:12] = [2.0, -1.0, 3.0]
    o.set_env_state(0, s)
    J = np.array([1.4e-5, 1.4e-5, 2.17e-5])
    L0 = np.linalg.norm(J * s[9:12])
    hover = np.full((1, 4), 0.027 * 9.8 / 4)
    for _ in range(25):
        out = o.step(hover)
    L1 = np.linalg.norm(J * out['state'][0][9:12])
    assert abs(L1 - L0) / L0 < 2e-3


def test_cartpole_energy_drift_is_small_without_input():
    env = H.make_env('cartpole', ctrl_freq=50, pyb_freq=1000, randomized_init=False, done_on_out_of_bound=False,
                     task_info={'stabilization_goal': [0], 'stabilization_goal_tolerance': 0.0})
    o = H.make_oracle(env, 1, auto_reset=False)
    s = np.array([0.0, 0.3, 0.4, -0.2])
    o.set_env_state(0, s)
    M, m, l, g = 1.0, 0.1, 0.5, 9.8
    Ic = m / 12 * (1.0 + 0.05**2)

    def energy(x):
        return 0.5 * (M + m) * x[1]**2 + m * l * np.cos(x[2]) * x[1] * x[3] + 0.5 * (Ic + m * l * l) * x[3]**2 + m * g * l * np.cos(x[2])
    e0 = energy(s)
    for _ in range(100):
        out = o.step(np.zeros((1, 1)))
    assert abs(energy(out['state'][0]) - e0) / abs(e0) < 5e-3


def test_3d_integrator_converges_to_the_reference_symbolic_model():
    '''As pyb_freq -> inf the 3-D oracle (quaternion exponential map, world-frame angular velocity, gyroscopic term) converges at
    first order to an RK integration of the reference's symbolic 3-D model -- Z-Y-X Euler angles, body rates, Euler-rate kinematics
    (quadrotor.py:533-562) -- evaluated with the URDF lever arm 0.028 (cf2x.urdf:42-78) instead of L/sqrt(2).'''
    from scipy.integrate import solve_ivp
    m, g, arm = 0.027, 9.8, 0.028
    J = np.array([1.4e-5, 1.4e-5, 2.17e-5])
    gamma = 7.94e-12 / 3.16e-10
    f = np.array([0.0672, 0.0660, 0.0655, 0.0668])                  # unequal motor thrusts: roll, pitch and yaw torques

    def rhs(t, x):
        phi, th, psi = x[6:9]
        w = x[9:12]
        cph, sph, cth, sth, cps, sps = np.cos(phi), np.sin(phi), np.cos(th), np.sin(th), np.cos(psi), np.sin(psi)
        Rob = np.array([[cps * cth, cps * sth * sph - sps * cph, cps * sth * cph + sps * sph],
                        [sps * cth, sps * sth * sph + cps * cph, sps * sth * cph - cps * sph],
                        [-sth, cth * sph, cth * cph]])                  # RotZ(psi) RotY(theta) RotX(phi) (transformations.py:173-187)
        acc = Rob @ np.array([0, 0, f.sum()]) / m - np.array([0, 0, g])
        Mb = np.array([arm * (f[0] + f[1] - f[2] - f[3]), arm * (-f[0] + f[1] + f[2] - f[3]), gamma * (-f[0] + f[1] - f[2] + f[3])])
        wdot = (Mb - np.cross(w, J * w)) / J
        W = np.array([[1, sph * np.tan(th), cph * np.tan(th)], [0, cph, -sph], [0, sph / cth, cph / cth]])
        ang = W @ w
        return [x[1], acc[0], x[3], acc[1], x[5], acc[2], ang[0], ang[1], ang[2], wdot[0], wdot[1], wdot[2]]
    s0 = np.zeros(12); s0[4] = 1.0; s0[6:9] = [0.08, -0.05, 0.3]        # zero initial rates (quadrotor.py:379 passes p,q,r as WORLD rates)
    ref = solve_ivp(rhs, (0, 0.2), s0, rtol=1e-12, atol=1e-14).y[:, -1]
    errs = []
    for pyb in (500, 1000, 2000, 4000):
        env = _quad(3, pyb_freq=pyb)
        o = H.make_oracle(env, 1, auto_reset=False)
        o.set_env_state(0, s0)
        for _ in range(10):
            out = o.step(f.reshape(1, 4))
        errs.append(np.abs(out['state'][0] - ref).max())
    ratios = np.asarray(errs[:-1]) / np.asarray(errs[1:])
    assert errs[-1] < 2e-3 and np.all(ratios > 1.7) and np.all(ratios < 2.3), (errs, ratios)     # O(dt)


def test_quadratic_constraint_rows_match_numpy():
    '''QuadraticConstraint (constraints.py:181-231): the oracle's CSR evaluation of z'Qz - b against the reference formula
    np.round(x.T @ F.T @ P @ F @ x - b, 8) evaluated with numpy on the oracle's own states.'''
    rng = np.random.default_rng(0)
    P6 = rng.normal(size=(6, 6)); P6 = P6 @ P6.T
    P4 = np.eye(4) + 0.2
    cons = [dict(constraint_form='quadratic_constraint', constrained_variable='state', P=P6.tolist(), b=3.0),
            dict(constraint_form='quadratic_constraint', constrained_variable='input_and_state', active_dims=[0, 2, 6, 7], P=P4.tolist(), b=2.0)]
    env = H.make_env('quadrotor_2D_track', constraints=cons)
    N = 16
    o = H.make_oracle(env, N, auto_reset=False)
    o.set_episode_table(H.episode_table(env, N, K=2, seed=1))
    o.reset()
    for t in range(20):
        act = rng.uniform(-1, 1, size=(N, 2))
        s = o.step(act)
        x = s['state']
        g0 = np.round(np.einsum('ni,ij,nj->n', x, P6, x) - 3.0, 8)
        np.testing.assert_allclose(s['cvals'][:, 0], g0, rtol=0, atol=2e-8)
        assert ((s['cvals'] > 0).any(1) == (s['cviol'] > 0)).all()


# ------------------------------------------------------------------------------------------------------------------
# non-default quadrotor physics (base_aviary.py:386-460) and the adversary input (benchmark_env.py:216-228)
# ------------------------------------------------------------------------------------------------------------------
def _motor(env, thrust_per_input):
    '''thrust command -> (rpm per motor, force per motor) through cmd2pwm / pwm2rpm (quadrotor_utils.py:16-60).'''
    n_mot = 4 // env.action_dim
    pwm = (np.sqrt(thrust_per_input / n_mot / env.KF) - env.PWM2RPM_CONST) / env.PWM2RPM_SCALE
    rpm = env.PWM2RPM_SCALE * np.clip(pwm, env.MIN_PWM, env.MAX_PWM) + env.PWM2RPM_CONST
    return rpm, env.KF * rpm**2


def test_downwash_of_a_single_drone_is_no_force():
    '''_downwash compares the drone with itself (delta_z = pos[2] - pos[2] = 0, base_aviary.py:448-451): pyb_dw == pyb, bit for bit.'''
    outs = []
    for phys in ('pyb', 'pyb_dw'):
        env = _quad(3, physics=phys)
        o = H.make_oracle(env, 1, auto_reset=False)
        s0 = np.zeros(12); s0[[0, 2, 4]] = [0.1, -0.2, 0.3]; s0[[1, 3, 5]] = [0.5, -0.4, 0.2]; s0[6:9] = [0.1, -0.2, 0.3]
        o.set_env_state(0, s0)
        for _ in range(10):
            out = o.step(np.array([[0.07, 0.06, 0.08, 0.065]]))
        outs.append(out['state'][0].copy())
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize('qt', [1, 2, 3])
def test_ground_effect_of_a_level_drone_matches_the_reference_formula(qt):
    '''Level attitude, equal motor commands: every propeller sits at the CoM height and gains KF rpm^2 GND_EFF_COEFF (PROP_RADIUS / (4 h))^2
    with h clipped at GND_EFF_H_CLIP (base_aviary.py:406-408), re-evaluated every sub-step; the attitude stays level.'''
    env = _quad(qt, physics='pyb_gnd')
    nx, nu = env.state_dim, env.action_dim
    o = H.make_oracle(env, 1, auto_reset=False)
    z0, iz = 0.03, {1: 0, 2: 2, 3: 4}[qt]                       # 3 cm above the world origin: below the clip height of ~3.8 cm at first
    s0 = np.zeros(nx); s0[iz] = z0; s0[iz + 1] = 0.4
    o.set_env_state(0, s0)
    cmd = np.full((1, nu), 0.9 * 0.027 * 9.8 / nu)
    rpm, f = _motor(env, cmd[0, 0])
    z, vz, dt = z0, 0.4, 1.0 / 1000
    assert env.GND_EFF_H_CLIP == pytest.approx(0.25 * env.PROP_RADIUS * np.sqrt(15 * env.GND_EFF_COEFF / 4))
    for _ in range(5):
        out = o.step(cmd)
        for _ in range(20):
            h = max(z, env.GND_EFF_H_CLIP)
            thrust = 4 * f * (1 + env.GND_EFF_COEFF * (env.PROP_RADIUS / (4 * h))**2)
            vz += (thrust - 0.027 * 9.8) / 0.027 * dt
            z += vz * dt
    np.testing.assert_allclose([out['state'][0][iz], out['state'][0][iz + 1]], [z, vz], rtol=1e-12, atol=1e-14)
    assert z > z0 + 0.4 * 0.1 * 0.5                                # sanity: the drone climbed through the clip height
    others = np.delete(out['state'][0], [iz, iz + 1])
    np.testing.assert_allclose(others, 0, atol=1e-13)             # symmetric gains: no torque, no lateral force


def test_drag_of_a_pitched_2d_drone_matches_the_reference_formula():
    '''_drag (base_aviary.py:420-440): -DRAG_COEFF sum(2 pi rpm / 60) * vel is rotated by the base rotation and then applied as a LINK_FRAME
    force, i.e. rotated once more.  Equal motor commands: no torque, the pitch stays constant, numpy restatement with explicit matrices.'''
    env = _quad(2, physics='pyb_drag')
    o = H.make_oracle(env, 1, auto_reset=False)
    th = 0.3
    s0 = np.array([0.1, 1.2, 1.0, -0.7, th, 0.0])
    o.set_env_state(0, s0)
    cmd = np.full((1, 2), 0.5 * 0.027 * 9.8)
    rpm, f = _motor(env, cmd[0, 0])
    Rm = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
    pos, vel, dt = np.array([0.1, 0.0, 1.0]), np.array([1.2, 0.0, -0.7]), 1.0 / 1000
    k = -env.DRAG_COEFF * np.sum(2 * np.pi * np.full(4, rpm) / 60)
    for _ in range(3):
        out = o.step(cmd)
        for _ in range(20):
            F = Rm @ np.array([0, 0, 4 * f]) + np.array([0, 0, -0.027 * 9.8]) + Rm @ (Rm @ (k * vel))
            vel = vel + F / 0.027 * dt
            pos = pos + vel * dt
    np.testing.assert_allclose(out['state'][0], [pos[0], vel[0], pos[2], vel[2], th, 0.0], rtol=1e-11, atol=1e-13)
    # the drag force is of the order of 1e-3 N at 1 m/s: it must have changed the trajectory measurably
    env0 = _quad(2); o0 = H.make_oracle(env0, 1, auto_reset=False); o0.set_env_state(0, s0)
    for _ in range(3):
        ref = o0.step(cmd)
    assert abs(ref['state'][0][1] - out['state'][0][1]) > 1e-4


def test_tilted_3d_ground_effect_produces_the_lever_arm_torque():
    '''Rolled drone near the ground: the lower propellers gain more thrust, which must show up as a restoring roll acceleration of
    a (f0' + f1' - f2' - f3') / Ixx after one sub-step (prop offsets (a,a) (-a,a) (-a,-a) (a,-a), cf2x.urdf:42-78).'''
    env = _quad(3, physics='pyb_gnd', pyb_freq=50)               # one sub-step per control step
    o = H.make_oracle(env, 1, auto_reset=False)
    roll = 0.2
    s0 = np.zeros(12); s0[4] = 0.05; s0[6] = roll
    o.set_env_state(0, s0)
    cmd = np.full((1, 4), 0.027 * 9.8 / 4)
    rpm, f = _motor(env, cmd[0, 0])
    a = 0.028
    # world height of prop i: z + (R [x, y, 0])_z with R = R_x(roll): y sin(roll)
    hs = np.array([0.05 + y * np.sin(roll) for y in (a, a, -a, -a)])
    fe = f * (1 + env.GND_EFF_COEFF * (env.PROP_RADIUS / (4 * np.maximum(hs, env.GND_EFF_H_CLIP)))**2)
    out = o.step(cmd)
    p_rate = out['state'][0][9]                                   # body roll rate after dt = 1/50
    assert p_rate == pytest.approx(a * (fe[0] + fe[1] - fe[2] - fe[3]) / 1.4e-5 / 50, rel=1e-9)
    assert hs[0] > hs[2] and p_rate < 0                           # props 0, 1 are the higher ones for roll > 0: less gain there, restoring torque


@pytest.mark.parametrize('name,mode', [('quadrotor', 'dynamics'), ('quadrotor', 'action'), ('cartpole', 'dynamics'), ('cartpole', 'action')])
def test_adversary_control_equals_the_matching_passive_step_disturbance(name, mode):
    '''set_adversary_control (benchmark_env.py:216-228) adds clip(a) * scale + offset to the same target a passive disturbance of that mode is
    added to (quadrotor.py:416-425,737-738; cartpole.py:494-495,543-551): with equal components it must reproduce a `step` disturbance of
    that magnitude bit for bit (values chosen exactly representable in float32, the adversary's arithmetic).'''
    kw = dict(quad_type=2, ctrl_freq=50, pyb_freq=1000, done_on_out_of_bound=False) if name == 'quadrotor' else dict(ctrl_freq=50, pyb_freq=50, done_on_out_of_bound=False)
    mag = 0.5 * 0.0078125 + 0.001953125                          # a = 0.5, scale = 2^-7, offset = 2^-9
    adv_env = H.make_env(name, randomized_init=False, adversary_disturbance=mode, adversary_disturbance_scale=0.0078125, adversary_disturbance_offset=0.001953125, **kw)
    ref_env = H.make_env(name, randomized_init=False, disturbances={mode: [{'disturbance_func': 'step', 'magnitude': mag, 'step_offset': 0}]}, **kw)
    oa, orf = H.make_oracle(adv_env, 2, auto_reset=False), H.make_oracle(ref_env, 2, auto_reset=False)
    oa.reset(); orf.reset()
    assert oa.adversary_dim == adv_env.adversary_action_space.shape[0] == adv_env.DISTURBANCE_MODES[mode]['dim']
    rng = np.random.default_rng(0)
    for t in range(12):
        act = rng.uniform(-0.3, 0.3, (2, adv_env.action_dim)) + (0.027 * 9.8 / 2 if name == 'quadrotor' else 0.0)
        oa.set_adversary_control(np.full((2, oa.adversary_dim), 0.5))
        a, r = oa.step(act), orf.step(act)
        assert np.array_equal(a['state'], r['state']) and np.array_equal(a['rew'], r['rew'])
    # the dynamics-mode action is consumed by the step (no call -> no force); the action-mode one persists until reset()
    a, r = oa.step(act), orf.step(act)
    assert np.array_equal(a['state'], r['state']) == (mode == 'action')
    # clipping to the adversary action space: 7.0 acts like 1.0
    o1, o2 = H.make_oracle(adv_env, 1, auto_reset=False), H.make_oracle(adv_env, 1, auto_reset=False)
    o1.reset(); o2.reset()
    o1.set_adversary_control(np.full((1, o1.adversary_dim), 7.0)); o2.set_adversary_control(np.full((1, o2.adversary_dim), 1.0))
    assert np.array_equal(o1.step(act[:1])['state'], o2.step(act[:1])['state'])
    # an env without an adversary refuses the call like the reference
    plain = H.make_oracle(H.make_env(name, **kw), 1, auto_reset=False)
    with pytest.raises(RuntimeError, match='adversary_disturbance does not exist'):
        plain.set_adversary_control(np.zeros((1, 1)))
