'''Reference-in-the-loop harness (build container only; needs /root/reference).

Runs the UNMODIFIED reference environments -- `safe_control_gym/envs/gym_pybullet_drones/quadrotor.py` (+ base_aviary.py, benchmark_env.py,
constraints.py, disturbances.py, quadrotor_utils.py, transformations.py, normalization.py) and `envs/gym_control/cartpole.py` -- with the
third-party packages that are NOT installed here replaced by stand-ins:

* `pybullet`: FakeBullet below -- the handful of API calls the envs make, backed by one floating rigid body (quadrotor) or a cart + pole chain
  (cart-pole) integrated the way DESIGN.md 2 / SURVEY.md 8(c) restate Bullet's `stepSimulation` (symplectic Euler, exponential-map orientation
  update, gyroscopic term, forces cleared every step).  The physics inside is therefore NOT a pin of PyBullet (only real PyBullet can be: the
  cart-pole is pinned by tests/golden/cartpole_pybullet_transitions.npz, the quadrotor by tools/record_pybullet_goldens.py when a wheel exists).
  What it buys: EVERYTHING AROUND the integrator -- action maps, PWM / RPM conversion, which link gets which force in which frame, URDF offsets,
  observation assembly (Euler angles, body rates), goal look-ahead, rewards, termination, constraint values, disturbance flow, randomisation
  order, time-limit truncation, reset -- is computed by the reference's own code, and the oracle has to reproduce it.
* `casadi`: an object that accepts any call / attribute / arithmetic (the symbolic model is built by the constructors but is out of scope).
* `gymnasium`: `spaces.Box`, `Env`, `utils.seeding.np_random` (same stream: numpy Generator over PCG64 seeded with SeedSequence(seed)).
* `matplotlib.pyplot`, `pybullet_data`: empty.
`import safe_control_gym` itself would import every subsystem; empty package shells with the right `__path__` let the env modules load from the
files where they lie.  No reference file is modified or copied.
'''
import math
import os
import sys
import types
import xml.etree.ElementTree as etxml

import numpy as np

REF_ROOT = '/root/reference'
REF = os.path.join(REF_ROOT, 'safe_control_gym')


# ------------------------------------------------------------------------------------------------
class _Any:
    '''casadi stand-in: survives any use, means nothing'''
    __array_priority__ = 1000
    __array_ufunc__ = None

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Any()

    def __getattr__(self, n):
        if n.startswith('__') and n.endswith('__'):
            raise AttributeError(n)
        return _Any()

    def __getitem__(self, i):
        return _Any()

    def __setitem__(self, i, v):
        pass

    def __iter__(self):
        return iter(())

    def __len__(self):
        return 0

    def _op(self, *a):
        return _Any()
    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = __truediv__ = __rtruediv__ = __matmul__ = __rmatmul__ = __pow__ = __rpow__ = _op
    __neg__ = __pos__ = _op


class _AnyModule(types.ModuleType):
    def __getattr__(self, n):
        if n.startswith('__') and n.endswith('__'):
            raise AttributeError(n)
        return _Any()


# ------------------------------------------------------------------------------------------------
def _quat_mul(a, b):      # (x, y, z, w)
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz])


def _rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


class FakeBullet(types.ModuleType):
    '''The PyBullet calls of base_aviary.py / quadrotor.py / cartpole.py on ONE body per client.'''
    GUI, DIRECT, LINK_FRAME, WORLD_FRAME = 1, 2, 1, 2
    COV_ENABLE_GUI = COV_ENABLE_RGB_BUFFER_PREVIEW = COV_ENABLE_DEPTH_BUFFER_PREVIEW = COV_ENABLE_SEGMENTATION_MARK_PREVIEW = 0
    ER_TINY_RENDERER = ER_SEGMENTATION_MASK_OBJECT_AND_LINKINDEX = 0
    VELOCITY_CONTROL, TORQUE_CONTROL = 0, 2
    JOINT_PRISMATIC, JOINT_REVOLUTE = 1, 0

    class World:
        def __init__(self):
            self.reset()

        def reset(self):
            self.gravity, self.dt = np.zeros(3), 1.0 / 240
            self.body = None        # ('quad', dict) | ('cartpole', dict)
            self.next_id = 0

    def __init__(self):
        super().__init__('pybullet')
        self.worlds = {}

    def _w(self, k):
        '''the world of the call's physicsClientId (one per p.connect(); a call without it goes to the first client, like PyBullet's default)'''
        cid = k.get('physicsClientId', 0)
        if cid not in self.worlds:
            self.worlds[cid] = FakeBullet.World()
        return self.worlds[cid]

    def __getattr__(self, n):                           # camera / GUI / logging calls: accepted, ignored
        if n.startswith('__') and n.endswith('__'):
            raise AttributeError(n)
        return lambda *a, **k: None

    # -- world -----------------------------------------------------------------------------------
    def connect(self, *a, **k):
        cid = len(self.worlds)
        self.worlds[cid] = FakeBullet.World()
        return cid

    def disconnect(self, *a, **k):
        pass

    def resetSimulation(self, *a, **k):
        self._w(k).reset()

    def setGravity(self, gx, gy, gz, **k):
        self._w(k).gravity = np.array([gx, gy, gz], dtype=np.float64)

    def setTimeStep(self, dt, **k):
        self._w(k).dt = float(dt)

    def setRealTimeSimulation(self, *a, **k):
        pass

    def setAdditionalSearchPath(self, *a, **k):
        pass

    def configureDebugVisualizer(self, *a, **k):
        pass

    def resetDebugVisualizerCamera(self, *a, **k):
        pass

    def setJointMotorControl2(self, bid, jointIndex, controlMode, force=0.0, **k):
        # cartpole.py:310-311 switches the default velocity motors off (VELOCITY_CONTROL, force 0): there are none here.
        # cartpole.py:576-581: TORQUE_CONTROL on the slider joint = the force of the next simulation step
        if controlMode == self.TORQUE_CONTROL:
            assert jointIndex == 0
            self._w(k).body[1]['force'] = float(force)

    def loadURDF(self, path, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1), **k):
        bid = self._w(k).next_id
        self._w(k).next_id += 1
        if os.path.basename(path) == 'plane.urdf':
            return bid
        tree = etxml.parse(path).getroot()
        if tree.attrib.get('name', '').startswith('cf2') or 'cf2' in os.path.basename(path):
            # Bullet link index i = child link of the i-th joint of the file.  LINK_FRAME of applyExternalForce / applyExternalTorque is the
            # link's centre-of-mass (inertial) frame: joint origin + the child link's <inertial><origin> (cf2x.urdf puts the propeller
            # positions there; the joints have no origin), all fixed and un-rotated here.
            inertial_origin = {}
            for l in tree.findall('link'):
                o = l.find('inertial/origin')
                assert o is None or all(float(v) == 0 for v in o.attrib.get('rpy', '0 0 0').split())
                inertial_origin[l.attrib['name']] = np.array([float(v) for v in o.attrib['xyz'].split()]) if o is not None else np.zeros(3)
            offs = []
            for j in tree.findall('joint'):
                o = j.find('origin')
                assert j.attrib['type'] == 'fixed' and (o is None or all(float(v) == 0 for v in o.attrib.get('rpy', '0 0 0').split()))
                jo = np.array([float(v) for v in o.attrib['xyz'].split()]) if o is not None else np.zeros(3)
                offs.append(jo + inertial_origin[j.find('child').attrib['link']])
            base = tree.find('link')
            self._w(k).body = ('quad', dict(id=bid, mass=float(base.find('inertial/mass').attrib['value']),
                                      J=np.array([float(base.find('inertial/inertia').attrib[k2]) for k2 in ('ixx', 'iyy', 'izz')]),
                                      link_off=offs,
                                      pos=np.array(basePosition, dtype=np.float64), quat=np.array(baseOrientation, dtype=np.float64),
                                      vel=np.zeros(3), omega=np.zeros(3), F=np.zeros(3), tau=np.zeros(3)))
        else:                                                      # cartpole_template-derived file written by cartpole.py
            # links of the file: slideBar (base, index -1), cart (index 0, prismatic joint along x), pole (index 1, revolute about y at the cart)
            pole = [l for l in tree.findall('link') if l.attrib['name'] == 'pole'][0]
            cart = [l for l in tree.findall('link') if l.attrib['name'] == 'cart'][0]
            box = np.array([float(v) for v in pole.find('collision/geometry/box').attrib['size'].split()])
            self._w(k).body = ('cartpole', dict(id=bid, q=np.zeros(2), qd=np.zeros(2), force=0.0, Qx=0.0, Qth=0.0,
                                          mass={0: float(cart.find('inertial/mass').attrib['value']), 1: float(pole.find('inertial/mass').attrib['value'])},
                                          l=float(pole.find('inertial/origin').attrib['xyz'].split()[-1]), box=box))
        return bid

    def changeDynamics(self, bid, linkIndex=-1, mass=None, localInertiaDiagonal=None, linearDamping=None, angularDamping=None, **k):
        if self._w(k).body is None:
            return
        kind, b = self._w(k).body
        link = linkIndex
        if kind == 'quad':
            if mass is not None:
                b['mass'] = float(mass)
            if localInertiaDiagonal is not None:
                b['J'] = np.array(localInertiaDiagonal, dtype=np.float64)
        else:
            if mass is not None:                       # link -1 = slider (cart) base?  cartpole.py:318-322: linkIndex 0 cart, 1 pole
                b['mass'][link] = float(mass)

    # -- quadrotor body ----------------------------------------------------------------------------
    def resetBasePositionAndOrientation(self, bid, pos, quat, **k):
        b = self._w(k).body[1]
        b['pos'], b['quat'] = np.array(pos, dtype=np.float64), np.array(quat, dtype=np.float64)

    def resetBaseVelocity(self, bid, linearVelocity=(0, 0, 0), angularVelocity=(0, 0, 0), **k):
        b = self._w(k).body[1]
        b['vel'], b['omega'] = np.array(linearVelocity, dtype=np.float64), np.array(angularVelocity, dtype=np.float64)

    def getBasePositionAndOrientation(self, bid, **k):
        b = self._w(k).body[1]
        return tuple(b['pos']), tuple(b['quat'])

    def getBaseVelocity(self, bid, **k):
        b = self._w(k).body[1]
        return tuple(b['vel']), tuple(b['omega'])

    def getQuaternionFromEuler(self, rpy, **k):          # Bullet: ZYX (yaw, pitch, roll) -> (x, y, z, w)
        r, p, y = (float(v) for v in rpy)
        cr, sr, cp, sp, cy, sy = math.cos(r / 2), math.sin(r / 2), math.cos(p / 2), math.sin(p / 2), math.cos(y / 2), math.sin(y / 2)
        return (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy)

    def getEulerFromQuaternion(self, q, **k):            # Bullet b3Quaternion::getEulerZYX incl. its gimbal branch
        x, y, z, w = (float(v) for v in q)
        sqx, sqy, sqz, sqw = x * x, y * y, z * z, w * w
        sarg = -2.0 * (x * z - w * y)
        if sarg <= -0.99999:
            return (0.0, -0.5 * math.pi, 2 * math.atan2(x, -y))
        if sarg >= 0.99999:
            return (0.0, 0.5 * math.pi, 2 * math.atan2(-x, y))
        return (math.atan2(2 * (y * z + w * x), sqw - sqx - sqy + sqz), math.asin(sarg), math.atan2(2 * (x * y + w * z), sqw + sqx - sqy - sqz))

    def getMatrixFromQuaternion(self, q, **k):
        return tuple(_rot(np.array(q, dtype=np.float64)).reshape(-1))

    def applyExternalForce(self, bid, linkIndex, forceObj, posObj, flags, **k):
        link = linkIndex
        kind, b = self._w(k).body
        f, r = np.array(forceObj, dtype=np.float64), np.array(posObj, dtype=np.float64)
        if kind == 'cartpole':                             # cartpole.py:557-566: a world-frame force on the pole link at a world point
            assert link == 1 and flags == self.WORLD_FRAME
            arm = r - np.array([b['q'][0], 0.0, 0.0])      # from the pivot (on the cart, at the slider coordinate)
            b['Qx'] += f[0]                                # generalised forces of (x, theta about +y)
            b['Qth'] += arm[2] * f[0] - arm[0] * f[2]
            return
        R = _rot(b['quat'])
        if flags == self.LINK_FRAME:
            off = b['link_off'][link] if link >= 0 else np.zeros(3)
            fw, arm = R @ f, R @ (off + r)
        else:
            fw, arm = f, r - b['pos']
        b['F'] += fw
        b['tau'] += np.cross(arm, fw)

    def applyExternalTorque(self, bid, linkIndex, torqueObj, flags, **k):
        b = self._w(k).body[1]
        t = np.array(torqueObj, dtype=np.float64)
        b['tau'] += _rot(b['quat']) @ t if flags == self.LINK_FRAME else t

    def getLinkState(self, bid, linkIndex, **k):           # cartpole.py:561: [0] = world position of the pole's centre of mass
        b = self._w(k).body[1]
        assert self._w(k).body[0] == 'cartpole' and linkIndex == 1
        x, th = b['q']
        p_ = (x + b['l'] * math.sin(th), 0.0, b['l'] * math.cos(th))
        return (p_, (0.0, 0.0, 0.0, 1.0))

    def getLinkStates(self, bid, linkIndices, **k):       # base_aviary.py:396 (ground effect): world position of each link frame
        b = self._w(k).body[1]
        R = _rot(b['quat'])
        return [(tuple(b['pos'] + R @ b['link_off'][i]),) * 2 for i in linkIndices]

    # -- cart-pole chain -----------------------------------------------------------------------------
    def resetJointState(self, bid, jointIndex, targetValue, targetVelocity=0.0, **k):
        b = self._w(k).body[1]
        b['q'][jointIndex], b['qd'][jointIndex] = float(targetValue), float(targetVelocity)

    def getJointState(self, bid, jointIndex, **k):
        b = self._w(k).body[1]
        return (float(b['q'][jointIndex]), float(b['qd'][jointIndex]), (0.0,) * 6, 0.0)

    def getNumJoints(self, *a, **k):
        return 2

    def getDynamicsInfo(self, bid, link, **k):
        kind, b = self._w(k).body
        if kind == 'quad':
            return (b['mass'], 0.0, tuple(b['J']))
        return (b['mass'].get(link, 0.0), 0.0, (0.0, 0.0, 0.0))

    # -- integration -----------------------------------------------------------------------------------
    def stepSimulation(self, **k):
        kind, b = self._w(k).body
        dt = self._w(k).dt
        if kind == 'quad':
            m, J = b['mass'], b['J']
            R = _rot(b['quat'])
            F = b['F'] + m * self._w(k).gravity
            wb = R.T @ b['omega']
            tau_b = R.T @ b['tau']
            wdot_b = (tau_b - np.cross(wb, J * wb)) / J
            b['vel'] = b['vel'] + F / m * dt
            b['omega'] = b['omega'] + (R @ wdot_b) * dt
            b['pos'] = b['pos'] + b['vel'] * dt
            w = b['omega']
            ang = float(np.linalg.norm(w))
            if ang * dt > 0.25 * math.pi:
                ang = 0.25 * math.pi / dt
            kk = (0.5 * dt - dt ** 3 * 0.020833333333 * ang * ang) if ang < 0.001 else math.sin(0.5 * ang * dt) / ang
            dq = np.array([w[0] * kk, w[1] * kk, w[2] * kk, math.cos(0.5 * ang * dt)])
            q = _quat_mul(dq, b['quat'])
            b['quat'] = q / np.linalg.norm(q)
            b['F'], b['tau'] = np.zeros(3), np.zeros(3)
        else:
            # cart + pole chain.  The pole's inertia about its centre of mass is the one Bullet derives from the collision box when the mass is
            # (re)set without URDF_USE_INERTIA_FROM_FILE (cartpole.py:301-322) -- the reading that reproduces the real-PyBullet transitions of
            # tests/golden/cartpole_pybullet_transitions.npz.  Symplectic Euler like the floating body.
            M, m, l, g = b['mass'][0], b['mass'][1], b['l'], -float(self._w(k).gravity[2])
            Ic = m / 12.0 * (b['box'][2] ** 2 + b['box'][0] ** 2)
            th, thd = b['q'][1], b['qd'][1]
            ct, st = math.cos(th), math.sin(th)
            a11, a12, a22 = M + m, m * l * ct, Ic + m * l * l
            b1, b2 = b['force'] + b['Qx'] + m * l * st * thd * thd, b['Qth'] + m * g * l * st
            det = a11 * a22 - a12 * a12
            qdd = np.array([(a22 * b1 - a12 * b2) / det, (a11 * b2 - a12 * b1) / det])
            b['qd'] = b['qd'] + qdd * dt
            b['q'] = b['q'] + b['qd'] * dt
            b['force'], b['Qx'], b['Qth'] = 0.0, 0.0, 0.0


# ------------------------------------------------------------------------------------------------
_INSTALLED = {}


def install():
    '''Puts the stand-ins and the package shells into sys.modules (idempotent); returns the FakeBullet instance.'''
    if _INSTALLED:
        return _INSTALLED['pb']
    pb = FakeBullet()
    sys.modules['pybullet'] = pb
    pd = types.ModuleType('pybullet_data')
    pd.getDataPath = lambda: '/nonexistent'
    sys.modules['pybullet_data'] = pd
    sys.modules['casadi'] = _AnyModule('casadi')
    mpl, plt = types.ModuleType('matplotlib'), types.ModuleType('matplotlib.pyplot')
    mpl.pyplot = plt
    sys.modules['matplotlib'], sys.modules['matplotlib.pyplot'] = mpl, plt

    g, sp, ut, sd = (types.ModuleType(n) for n in ('gymnasium', 'gymnasium.spaces', 'gymnasium.utils', 'gymnasium.utils.seeding'))

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            self.low = (np.asarray(low, dtype) if shape is None else np.broadcast_to(np.asarray(low, dtype), shape)).copy()
            self.high = (np.asarray(high, dtype) if shape is None else np.broadcast_to(np.asarray(high, dtype), shape)).copy()
            self.shape = self.low.shape

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)
            return [seed]

        def sample(self):                                  # gymnasium.spaces.Box.sample for bounded boxes (its own generator; not the same stream)
            if not hasattr(self, '_rng'):
                self._rng = np.random.default_rng()
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

    class Env:
        metadata = {}

        def close(self):
            pass

    class Wrapper(Env):                                   # gymnasium.Wrapper: holds `env`, forwards everything it does not define
        def __init__(self, env):
            self.env = env

        def __getattr__(self, n):
            if n.startswith('__') or n == 'env':
                raise AttributeError(n)
            return getattr(self.env, n)

        def step(self, action):
            return self.env.step(action)

        def reset(self, **kw):
            return self.env.reset(**kw)

        def close(self):
            return self.env.close()

        @property
        def unwrapped(self):
            return getattr(self.env, 'unwrapped', self.env)

    def np_random(seed=None):
        ss = np.random.SeedSequence(seed)
        return np.random.Generator(np.random.PCG64(ss)), ss.entropy
    sp.Box, g.spaces, g.Env, g.Wrapper, g.utils, ut.seeding, sd.np_random = Box, sp, Env, Wrapper, ut, sd, np_random
    for m in (g, sp, ut, sd):
        sys.modules[m.__name__] = m
    for name, sub in (('safe_control_gym', ''), ('safe_control_gym.envs', 'envs'), ('safe_control_gym.envs.gym_pybullet_drones', 'envs/gym_pybullet_drones'),
                      ('safe_control_gym.envs.gym_control', 'envs/gym_control'), ('safe_control_gym.math_and_models', 'math_and_models'),
                      ('safe_control_gym.controllers', 'controllers'), ('safe_control_gym.controllers.lqr', 'controllers/lqr'),
                      ('safe_control_gym.utils', 'utils'), ('safe_control_gym.controllers.ppo', 'controllers/ppo'), ('safe_control_gym.controllers.rarl', 'controllers/rarl'), ('safe_control_gym.experiments', 'experiments'), ('safe_control_gym.math_and_models.metrics', 'math_and_models/metrics'), ('safe_control_gym.controllers.sac', 'controllers/sac'), ('safe_control_gym.controllers.ddpg', 'controllers/ddpg'), ('safe_control_gym.controllers.safe_explorer', 'controllers/safe_explorer'),
                      ('safe_control_gym.envs.env_wrappers', 'envs/env_wrappers'),
                      ('safe_control_gym.envs.env_wrappers.vectorized_env', None)):
        if sub is None:
            continue                      # (a real package whose __init__ is wanted: make_vec_envs lives there)
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, sub)]
        sys.modules[name] = m
    if 'munch' not in sys.modules:                           # base_experiment.py:9 `munchify`: dicts with attribute access, recursively
        try:
            import munch  # noqa: F401
        except ImportError:
            mm = types.ModuleType('munch')

            class Munch(dict):
                __getattr__ = dict.__getitem__
                __setattr__ = dict.__setitem__

            def munchify(x):
                if isinstance(x, dict):
                    return Munch((k, munchify(v)) for k, v in x.items())
                if isinstance(x, (list, tuple)):
                    return type(x)(munchify(v) for v in x)
                return x
            mm.Munch, mm.munchify = Munch, munchify
            sys.modules['munch'] = mm
    if 'termcolor' not in sys.modules:
        try:
            import termcolor  # noqa: F401
        except ImportError:
            tc = types.ModuleType('termcolor')
            tc.colored = lambda s, *a, **k: s
            sys.modules['termcolor'] = tc
    for name in ('imageio', 'munch'):                    # imported at module level by utils/logging.py and utils/utils.py, used by neither path here
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = _AnyModule(name)
    try:
        import torch.utils.tensorboard  # noqa: F401
    except Exception:
        tb = types.ModuleType('torch.utils.tensorboard')
        tb.SummaryWriter = _Any
        sys.modules['torch.utils.tensorboard'] = tb
    _INSTALLED['pb'] = pb
    return pb


def load_task_config(rel_yaml):
    import yaml
    with open(os.path.join(REF_ROOT, rel_yaml)) as f:
        return yaml.safe_load(f)['task_config']


def make_ppo(env_id, task_config, output_dir, **ppo_config):
    '''The reference PPO controller (controllers/ppo/ppo.py) on a DummyVecEnv of reference envs.'''
    from functools import partial
    install()
    from safe_control_gym.controllers.ppo.ppo import PPO
    env_func = partial(make_quadrotor if env_id == 'quadrotor' else make_cartpole, **task_config)
    return PPO(env_func, training=True, output_dir=output_dir, **ppo_config)


def make_rarl(env_id, task_config, output_dir, **rarl_config):
    '''The reference RARL controller (controllers/rarl/rarl.py) on a DummyVecEnv of reference envs.'''
    from functools import partial
    install()
    from safe_control_gym.controllers.rarl.rarl import RARL
    env_func = partial(make_quadrotor if env_id == 'quadrotor' else make_cartpole, **task_config)
    return RARL(env_func, training=True, output_dir=output_dir, **rarl_config)


def make_safe_ppo(env_id, task_config, output_dir, **config):
    '''The reference SafeExplorerPPO controller (controllers/safe_explorer/safe_ppo.py) on a DummyVecEnv of reference envs.'''
    from functools import partial
    install()
    from safe_control_gym.controllers.safe_explorer.safe_ppo import SafeExplorerPPO
    env_func = partial(make_quadrotor if env_id == 'quadrotor' else make_cartpole, **task_config)
    return SafeExplorerPPO(env_func, training=True, output_dir=output_dir, **config)


def make_sac(env_id, task_config, output_dir, **config):
    '''The reference SAC controller (controllers/sac/sac.py) on a DummyVecEnv of reference envs.'''
    from functools import partial
    install()
    from safe_control_gym.controllers.sac.sac import SAC
    env_func = partial(make_quadrotor if env_id == 'quadrotor' else make_cartpole, **task_config)
    return SAC(env_func, training=True, output_dir=output_dir, **config)


def make_rap(env_id, task_config, output_dir, **config):
    '''The reference RAP controller (controllers/rarl/rap.py) on a DummyVecEnv of reference envs.'''
    from functools import partial
    install()
    from safe_control_gym.controllers.rarl.rap import RAP
    env_func = partial(make_quadrotor if env_id == 'quadrotor' else make_cartpole, **task_config)
    # rap.py does not implement BaseController's abstract `select_action` (its run() inlines the call, rap.py:218-220), so the class cannot be
    # instantiated as shipped; a subclass that only adds that method -- the body RARL has (rarl.py:195-210) -- is the least intrusive way to run it
    import torch

    class InstantiableRAP(RAP):
        def select_action(self, obs, info=None):
            with torch.inference_mode():
                return self.agent.ac.act(torch.FloatTensor(obs).to(self.device))
    return InstantiableRAP(env_func, training=True, output_dir=output_dir, **config)


def make_cartpole(**task_config):
    install()
    from safe_control_gym.envs.gym_control.cartpole import CartPole
    return CartPole(**task_config)


def make_quadrotor(**task_config):
    install()
    from safe_control_gym.envs.gym_pybullet_drones.quadrotor import Quadrotor
    return Quadrotor(**task_config)
