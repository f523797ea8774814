"""Builds libssdnerf_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libssdnerf_b200.so')

NVCC_FLAGS = [
    '-O3', '-std=c++17', '-lineinfo',
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-Xcompiler', '-fPIC,-fvisibility=hidden',
    '--expt-relaxed-constexpr',
    '-shared', '-cudart', 'shared',
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(HERE, '..', 'include', '*.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB] + sources()
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError('nvcc failed building libssdnerf_b200.so')
    if verbose:
        print(res.stdout)
    return LIB


if __name__ == '__main__':
    build_lib(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print(LIB)
