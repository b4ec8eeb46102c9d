'''In-tree build of the CUDA library (sm_100a).  (The CPU oracle of the test-suite builds itself: oracle/oracle.py.)

    python -m safe_control_gym_b200.build [--force] [--prof]

nvcc cross-compiles without a GPU.  The resulting libscg_b200.so sits next to this file (git-ignored, but shipped
to the GPU box with the working tree).
'''

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
ROLLOUT_UNITS = [f'scg_rollout_{s}_{p}.cu' for s in ('quad3d', 'quad2d', 'cartpole', 'quad1d') for p in ('f64', 'f32')]   # slowest first
UNITS = ROLLOUT_UNITS + ['scg_api.cu', 'scg_ppo.cu']          # translation units of libscg_b200.so (compiled in parallel)
SOURCES = UNITS + ['scg_kernels.cuh', 'scg_device.cuh', 'scg_rollout_launch.cuh', 'scg_rollout_tc.cuh']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-shared', '-Xcompiler', '-fPIC',
              '--compress-mode=size']     # line tables make the fatbin large; the compressed image travels to the GPU box faster


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_cuda(force=False, prof=False, verbose=True, dev=False):
    '''prof: -DSCG_PROF phase counters (libscg_b200_prof.so); dev: hidden-64 rollout variants only, for quick kernel experiments
    (libscg_b200_dev.so); select either at run time with SCG_LIB_NAME.'''
    out = os.path.join(HERE, 'libscg_b200_prof.so' if prof else 'libscg_b200_dev.so' if dev else 'libscg_b200.so')
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(ROOT, 'include', 'scg.h')]
    if not force and not _stale(out, deps):
        return out
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = ([nvcc] + NVCC_FLAGS + ['--threads', str(min(len(UNITS), os.cpu_count() or 1))] + (['-DSCG_PROF'] if prof else []) + (['-DSCG_DEV_HID64_ONLY'] if dev else [])
           + ['-I', os.path.join(ROOT, 'include'), '-o', out] + [os.path.join(CSRC, u) for u in UNITS])
    if verbose:
        print(' '.join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return out


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    build_cuda(force='--force' in argv, prof='--prof' in argv, dev='--dev' in argv)


if __name__ == '__main__':
    main()
