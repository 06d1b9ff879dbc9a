"""Build libgprmax_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

    python -m gprmax_b200.build [--force] [--verbose]

Translation units: gpb_core.cu (C ABI, host logic, register-vectorised and scalar kernels), the host-only gpb_idbuild.cpp and
gpb_vtkio.cpp (ID build, VTK re-ordering) and gpb_tma_inst.cu once per
(float type, PML variant) -- the TMA-staged kernels are specialised on the PML formulation and order, and the eight
variants compile in parallel.  Objects go to build/obj; the .so is git-ignored but travels to the GPU box with the
source snapshot.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libgprmax_b200.so')
OBJDIR = os.path.join(os.path.dirname(HERE), 'build', 'obj')
DEPS = ['gpb_core.cu', 'gpb_idbuild.cpp', 'gpb_vtkio.cpp', 'gpb_tma_inst.cu', 'gpb_tma_item.inc', 'gpb_tma.h', 'gpb_kernels.cuh', 'gpb_kernels_v4.cuh', 'gpb_kernels_coop.cuh', 'gpb_kernels_tma.cuh', 'gpb_kernels_pair.cuh',
        os.path.join('..', '..', 'include', 'gprmax_b200.h')]
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']


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


def units():
    """(object name, source, extra defines) of every translation unit."""
    out = [('gpb_core.o', 'gpb_core.cu', []), ('gpb_idbuild.o', 'gpb_idbuild.cpp', ['-x', 'cu']), ('gpb_vtkio.o', 'gpb_vtkio.cpp', ['-x', 'cu'])]
    for rname, rtype in (('f32', 'float'), ('f64', 'double')):
        for pv in range(4):   # 2 * formulation (HORIPML, MRIPML) + order - 1
            out.append(('gpb_tma_{}_pv{}.o'.format(rname, pv), 'gpb_tma_inst.cu', ['-DGPB_TMA_R=' + rtype, '-DGPB_TMA_PV={}'.format(pv)]))
    return out


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = nvcc_path()
    os.makedirs(OBJDIR, exist_ok=True)
    common = [nvcc, '-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-diag-suppress', '1886'] + ARCH
    if verbose:
        common += ['-Xptxas', '-v']
    common += os.environ.get('GPB_NVCC_FLAGS', '').split()

    def compile_one(u):
        obj, src, defs = u
        cmd = common + defs + ['-c', os.path.join(CSRC, src), '-o', os.path.join(OBJDIR, obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return u, r

    jobs = int(os.environ.get('GPB_BUILD_JOBS', '0')) or min(len(units()), os.cpu_count() or 1)
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        results = list(ex.map(compile_one, units()))
    for (obj, src, defs), r in results:
        if verbose or r.returncode:
            sys.stderr.write('--- {} {}\n{}{}'.format(src, ' '.join(defs), r.stdout, r.stderr))
        if r.returncode:
            raise RuntimeError('nvcc failed on {} {}'.format(src, ' '.join(defs)))
    objs = [os.path.join(OBJDIR, u[0]) for u in units()]
    subprocess.run([nvcc, '-shared', '-o', LIB] + ARCH + objs, check=True)
    return LIB


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
