"""Build libgprmax_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

    python -m gprmax_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the source snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libgprmax_b200.so')
SOURCES = ['gpb_core.cu']
DEPS = ['gpb_core.cu', 'gpb_kernels.cuh', 'gpb_kernels_v4.cuh', 'gpb_kernels_tma.cuh', os.path.join('..', '..', 'include', 'gprmax_b200.h')]


def nvcc_path():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
           '-Xcompiler', '-fPIC', '-shared', '-o', LIB]
    if verbose:
        cmd += ['-Xptxas', '-v']
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
