'''Abstract vectorised env (safe_control_gym/envs/env_wrappers/vectorized_env/vec_env.py:13-141).'''

from abc import ABC, abstractmethod


class VecEnv(ABC):
    closed = False
    viewer = None
    metadata = {'render.modes': []}

    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space

    @abstractmethod
    def reset(self):
        pass

    @abstractmethod
    def step_async(self, actions):
        pass

    @abstractmethod
    def step_wait(self):
        pass

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close_extras(self):
        pass

    def close(self):
        if self.closed:
            return
        self.close_extras()
        self.closed = True

    @property
    def unwrapped(self):
        return self.venv.unwrapped if isinstance(self, VecEnvWrapper) else self

    @abstractmethod
    def get_attr(self, attr_name, indices=None):
        pass

    @abstractmethod
    def set_attr(self, attr_name, values, indices=None):
        pass

    @abstractmethod
    def env_method(self, method_name, method_args=None, method_kwargs=None, indices=None):
        pass

    def _get_indices(self, indices):
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return indices


class VecEnvWrapper(VecEnv):
    def __init__(self, venv, observation_space=None, action_space=None):
        self.venv = venv
        super().__init__(num_envs=venv.num_envs, observation_space=observation_space or venv.observation_space,
                         action_space=action_space or venv.action_space)

    def step_async(self, actions):
        self.venv.step_async(actions)

    def close(self):
        return self.venv.close()

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(f'attempted to get missing private attribute \'{name}\'')
        return getattr(self.venv, name)

    def get_attr(self, attr_name, indices=None):
        return self.venv.get_attr(attr_name, indices)

    def set_attr(self, attr_name, values, indices=None):
        return self.venv.set_attr(attr_name, values, indices)

    def env_method(self, method_name, method_args=None, method_kwargs=None, indices=None):
        return self.venv.env_method(method_name, method_args=method_args, method_kwargs=method_kwargs, indices=indices)
