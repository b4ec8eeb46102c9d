'''Disturbance specifications for the CUDA engine.

Host-side mirror of safe_control_gym/envs/disturbances.py.  The reference applies each disturbance with
numpy on the host every step; here a disturbance is a POD record (`scg_disturbance`) evaluated inside the
kernels, and its randomness comes either from host tables drawn with the reference's own numpy calls
(`draw_episode` / `draw_step`, exact stream for one env) or from the in-kernel Philox streams.

Order of application inside a step: action -> dynamics -> observation
(quadrotor.py:735-736, 419-421, 806-807; cartpole.py:492-493, 546-547, 595-596).
'''

import numpy as np

from safe_control_gym_b200 import _abi

MODES = ('action', 'dynamics', 'observation')       # SCG_MODE_* order
MODE_INDEX = {m: i for i, m in enumerate(MODES)}


class Disturbance:
    '''One disturbance record (disturbances.py:6-35).'''

    KIND = None
    needs_offset = False    # draws an integer step offset per episode when step_offset is None
    step_draws = None       # None | 'uniform' | 'normal': per-step random numbers, `dim` of them

    def __init__(self, env, dim, mask=None, **kwargs):
        self.dim = int(dim)
        if self.dim > _abi.SCG_MAX_DIST_DIM:
            raise ValueError('disturbance dimension too large for the CUDA engine')
        self.mask = None
        if mask is not None:
            self.mask = np.asarray(mask, dtype=np.float32)
            assert self.dim == len(self.mask)
        self.max_step = int(env.EPISODE_LEN_SEC / env.CTRL_TIMESTEP)

    def fill(self, rec):
        rec.kind = self.KIND
        rec.dim = self.dim
        rec.has_mask = 0 if self.mask is None else 1
        rec.step_offset = 0
        for k in range(self.dim):
            rec.mask[k] = 1.0 if self.mask is None else float(self.mask[k])


class ImpulseDisturbance(Disturbance):
    '''disturbances.py:70-123.'''
    KIND = _abi.SCG_DIST_IMPULSE

    def __init__(self, env, dim, mask=None, magnitude=1, step_offset=None, duration=1, decay_rate=1, **kwargs):
        super().__init__(env, dim, mask)
        assert duration >= 1
        assert 0 < decay_rate <= 1
        self.magnitude, self.step_offset, self.duration, self.decay_rate = magnitude, step_offset, duration, decay_rate
        self.needs_offset = step_offset is None

    def fill(self, rec):
        super().fill(rec)
        rec.magnitude, rec.duration, rec.decay_rate = float(self.magnitude), float(self.duration), float(self.decay_rate)
        rec.step_offset = -1 if self.step_offset is None else int(self.step_offset)


class StepDisturbance(Disturbance):
    '''disturbances.py:126-163.'''
    KIND = _abi.SCG_DIST_STEP

    def __init__(self, env, dim, mask=None, magnitude=1, step_offset=None, **kwargs):
        super().__init__(env, dim, mask)
        self.magnitude, self.step_offset = magnitude, step_offset
        self.needs_offset = step_offset is None

    def fill(self, rec):
        super().fill(rec)
        rec.magnitude = float(self.magnitude)
        rec.step_offset = -1 if self.step_offset is None else int(self.step_offset)


class UniformNoise(Disturbance):
    '''disturbances.py:166-192 (including the `isinstance(low, list)` check used for `high`, :182).'''
    KIND = _abi.SCG_DIST_UNIFORM
    step_draws = 'uniform'

    def __init__(self, env, dim, mask=None, low=0.0, high=1.0, **kwargs):
        super().__init__(env, dim, mask)
        if isinstance(low, float):
            self.low = np.asarray([low] * self.dim)
        elif isinstance(low, list):
            self.low = np.asarray(low)
        else:
            raise ValueError('[ERROR] UniformNoise.__init__(): low must be specified as a float or list.')
        if isinstance(high, float):
            self.high = np.asarray([high] * self.dim)
        elif isinstance(low, list):
            self.high = np.asarray(high)
        else:
            raise ValueError('[ERROR] UniformNoise.__init__(): high must be specified as a float or list.')

    def fill(self, rec):
        super().fill(rec)
        for k in range(self.dim):
            rec.low[k], rec.high[k] = float(self.low[k]), float(self.high[k])


class WhiteNoise(Disturbance):
    '''disturbances.py:195-223.'''
    KIND = _abi.SCG_DIST_WHITE
    step_draws = 'normal'

    def __init__(self, env, dim, mask=None, std=1.0, **kwargs):
        super().__init__(env, dim, mask)
        if isinstance(std, float):
            self.std = np.asarray([std] * self.dim)
        elif isinstance(std, list):
            self.std = np.asarray(std)
        else:
            raise ValueError('[ERROR] WhiteNoise.__init__(): std must be specified as a float or list.')
        assert self.dim == len(self.std), 'std shape should be the same as dim.'

    def fill(self, rec):
        super().fill(rec)
        for k in range(self.dim):
            rec.std[k] = float(self.std[k])


class PeriodicNoise(Disturbance):
    '''disturbances.py:233-259 (the mask is dropped, :244; a fresh random phase every call, :253).'''
    KIND = _abi.SCG_DIST_PERIODIC
    step_draws = 'uniform'

    def __init__(self, env, dim, mask=None, scale=1.0, frequency=1.0, **kwargs):
        super().__init__(env, dim)
        self.scale, self.frequency = scale, frequency

    def fill(self, rec):
        super().fill(rec)
        rec.scale, rec.frequency = float(self.scale), float(self.frequency)


DISTURBANCE_TYPES = {
    'impulse': ImpulseDisturbance,
    'step': StepDisturbance,
    'uniform': UniformNoise,
    'white_noise': WhiteNoise,
    'periodic': PeriodicNoise,
}


class DisturbanceList:
    '''disturbances.py:38-67.'''

    def __init__(self, disturbances):
        if len(disturbances) > _abi.SCG_MAX_DIST:
            raise ValueError(f'at most {_abi.SCG_MAX_DIST} disturbances per mode are supported by the CUDA engine')
        self.disturbances = disturbances

    def seed(self, env):
        self.np_random = env.np_random

    def reset(self, env):
        return


def create_disturbance_list(disturbance_specs, shared_args, env):
    '''YAML specs -> DisturbanceList (disturbances.py:285-303).'''
    disturb_list = []
    for disturb in disturbance_specs:
        assert 'disturbance_func' in disturb.keys(), '[ERROR]: Every distrubance must specify a disturbance_func.'
        disturb_func = disturb['disturbance_func']
        assert disturb_func in DISTURBANCE_TYPES, '[ERROR] in BenchmarkEnv._setup_disturbances(), disturbance type not available.'
        cfg = {key: disturb[key] for key in disturb if key != 'disturbance_func'}
        disturb_list.append(DISTURBANCE_TYPES[disturb_func](env, **shared_args, **cfg))
    return DisturbanceList(disturb_list)


# ---------------------------------------------------------------------------------------------
# engine views
# ---------------------------------------------------------------------------------------------
def fill_config(cfg, disturbances):
    '''Writes the per-mode records into an scg_config.'''
    for mode, idx in MODE_INDEX.items():
        dl = disturbances.get(mode)
        n = 0 if dl is None else len(dl.disturbances)
        cfg.n_dist[idx] = n
        for j in range(n):
            dl.disturbances[j].fill(cfg.dist[idx][j])


def offset_slots(disturbances):
    '''Disturbances that draw a per-episode step offset, in engine (mode, list) order.'''
    out = []
    for mode in MODES:
        dl = disturbances.get(mode)
        if dl is not None:
            out += [d for d in dl.disturbances if d.needs_offset]
    return out


def noise_width(disturbances):
    w = 0
    for mode in MODES:
        dl = disturbances.get(mode)
        if dl is not None:
            w += sum(d.dim for d in dl.disturbances if d.step_draws is not None)
    return w


def draw_episode_offsets(disturbances, reset_order, np_random):
    '''Per-episode integer offsets drawn in the REFERENCE's order (the YAML order of the modes,
    benchmark_env.py:336-337, disturbances.py:102-106,148-151), returned in engine order.'''
    drawn = {}
    for mode in reset_order:
        dl = disturbances.get(mode)
        if dl is None:
            continue
        for d in dl.disturbances:
            if d.needs_offset:
                drawn[id(d)] = float(np_random.integers(d.max_step))
    return [drawn[id(d)] for d in offset_slots(disturbances)]


def draw_step_noise(disturbances, np_random):
    '''One control step's random numbers with the reference's numpy calls and call order
    (standard normals for white noise, U[0,1) for uniform / periodic phase).'''
    row = []
    for mode in MODES:
        dl = disturbances.get(mode)
        if dl is None:
            continue
        for d in dl.disturbances:
            if d.step_draws == 'normal':
                row += list(np_random.normal(0, 1.0, size=d.dim))       # std * N(0,1) == normal(0, std)
            elif d.step_draws == 'uniform':
                row += list(np_random.uniform(0.0, 1.0, size=d.dim))    # low + (high-low) * U[0,1)
    return row
