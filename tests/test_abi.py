'''The C-ABI shared library loads here (no GPU) and exports every symbol include/scg.h declares; compute entry points fail loudly.'''
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers as H
from safe_control_gym_b200 import _abi


def test_header_symbols_are_all_exported_and_bound():
    hdr = open(os.path.join(H.ROOT, 'include', 'scg.h')).read()
    declared = set(re.findall(r'\b(scg_[a-z_0-9]+)\s*\(', hdr))
    declared -= {'scg_env'}
    lib = _abi.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f'{name} declared in include/scg.h but not exported'
    assert declared == set(_abi._SYMBOLS), (declared ^ set(_abi._SYMBOLS))
    assert lib.scg_abi_version() == _abi.SCG_ABI_VERSION


def test_config_struct_layout_matches_the_header():
    '''sizeof(scg_config) is checked by scg_create (struct_size guard); here the ctypes mirror is checked against gcc.'''
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "scg.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu", sizeof(scg_config), sizeof(scg_disturbance),' \
          ' sizeof(scg_step_out), sizeof(scg_rollout_out), sizeof(scg_rollout_args), sizeof(scg_ppo_args));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, 'a.c'), 'w').write(src)
        subprocess.run(['gcc', '-I', os.path.join(H.ROOT, 'include'), os.path.join(d, 'a.c'), '-o', os.path.join(d, 'a')], check=True)
        sizes = [int(x) for x in subprocess.run([os.path.join(d, 'a')], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(_abi.scg_config), C.sizeof(_abi.scg_disturbance), C.sizeof(_abi.scg_step_out),
                     C.sizeof(_abi.scg_rollout_out), C.sizeof(_abi.scg_rollout_args), C.sizeof(_abi.scg_ppo_args)]


def test_policy_pack_layout_agrees_between_c_and_python():
    '''scg_policy_pack (C, host) and engine.pack_policy (torch) must produce the same blob.'''
    torch = pytest.importorskip('torch')
    from safe_control_gym_b200.controllers.ppo.ppo_utils import MLPActorCritic
    from safe_control_gym_b200.engine import pack_policy, policy_blob_layout
    from safe_control_gym_b200.spaces import Box
    from oracle import oracle as orc
    obs_dim, act_dim, hid = 12, 2, 64
    torch.manual_seed(3)
    ac = MLPActorCritic(Box(-np.ones(obs_dim), np.ones(obs_dim)), Box(-np.ones(act_dim), np.ones(act_dim)), [hid, hid], 'tanh')
    blob, dims = pack_policy(ac, 'cpu')
    assert dims == (obs_dim, act_dim, hid)
    assert _abi.lib().scg_policy_blob_floats(obs_dim, act_dim, hid) == policy_blob_layout(obs_dim, act_dim, hid)['total'] == blob.numel()
    pol = orc.OraclePolicy.from_state_dict(ac.state_dict(), 'tanh')
    ref, _ = H.policy_blob(pol, 'cpu')
    np.testing.assert_array_equal(blob.numpy(), ref.numpy())


def test_ppo_train_layout_is_dense_nn_linear_order():
    '''scg_ppo_param_layout: sections in the reference's parameter order, 4-float aligned, no overlap.'''
    from safe_control_gym_b200.engine import TRAIN_SECTIONS, ppo_train_layout
    obs_dim, act_dim, hid = 12, 2, 64
    L = ppo_train_layout(obs_dim, act_dim, hid)
    sizes = [hid * obs_dim, hid, hid * hid, hid, act_dim * hid, act_dim, act_dim, hid * obs_dim, hid, hid * hid, hid, hid, 1]
    o = 0
    for name, n in zip(TRAIN_SECTIONS, sizes):
        assert L[name] == o and o % 4 == 0, name
        o += (n + 3) // 4 * 4
        if name == 'actor.logstd':
            assert L['critic0'] == o
    assert L['total'] == o


def test_no_cpu_fallback_create_fails_loudly_without_a_gpu():
    torch = pytest.importorskip('torch')
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    env = H.make_env('quadrotor_2D_track')
    with pytest.raises(_abi.ScgError, match='no CUDA device'):
        env.reset()
    cfg = env.build_config(num_envs=4)
    cfg.pyb_freq = 77
    h = C.c_void_p()
    assert _abi.lib().scg_create(C.byref(cfg), C.byref(h)) == _abi.SCG_EINVAL
    assert b'not divisible' in _abi.lib().scg_last_error(None)


def test_fused_collector_variants_fit_the_sm():
    '''Resource usage of every rollout_kernel instantiation in the built library (cuobjdump, no GPU needed): registers within the
    256-thread budget of one SM (64 K registers => <= 255 per thread for one CTA, and the 8-warp default must keep one CTA resident),
    static shared memory below the 48 KB static limit, sm_100a code present.'''
    import re
    import shutil
    import subprocess
    from safe_control_gym_b200 import _abi
    if shutil.which('cuobjdump') is None:
        import pytest
        pytest.skip('cuobjdump not on PATH')
    txt = subprocess.run(['cuobjdump', '--dump-resource-usage', _abi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert 'sm_100a' in txt
    rows = re.findall(r'Function (_Z14rollout_kernel\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)', txt)
    assert len(rows) >= 8 * 7, len(rows)          # 4 systems x 2 precisions x 7 (hidden, envs-per-CTA) variants
    for name, reg, stack, shared in rows:
        m = re.search(r'ILi(\d)E([df])Li(\d+)ELi(\d+)ELi(\d+)E', name)
        nt = int(m.group(5))
        assert int(reg) <= 255 and int(reg) * nt <= 65536, (name, reg)
        assert int(shared) <= 48 * 1024, (name, shared)
        assert int(stack) <= 2048, (name, stack)    # call frames of the out-of-line cold paths (structs by value); only touched there
