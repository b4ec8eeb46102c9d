'''Factory with the reference's signature (vectorized_env/__init__.py:42-66).'''

from safe_control_gym_b200.envs.env_wrappers.vectorized_env.cuda_vec_env import CudaVecEnv


def make_vec_envs(env_func, env_configs=None, batch_size=1, n_processes=1, seed=None, **engine_kwargs):
    '''`n_processes` is accepted and ignored: the batch lives on one GPU and is stepped by one kernel.'''
    if env_configs is not None and any(c for c in env_configs):
        raise NotImplementedError('per-env differing configs are not supported by the batched engine')
    return CudaVecEnv(env_func, num_envs=batch_size, seed=seed, **engine_kwargs)
