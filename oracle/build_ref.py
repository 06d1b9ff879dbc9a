#!/usr/bin/env python
"""Build recipe: compile the REFERENCE's own Cython kernels into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product
(gprmax_b200/); only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may use it.

What it does
------------
The reference solver's CPU hot path is Cython (`gprMax/fields_updates_ext.pyx`,
`gprMax/pml_updates/*_ext.pyx`, `gprMax/snapshots_ext.pyx`), built by the
reference with `-O3 -w -fopenmp -march=native` (reference setup.py:176-180) and
directives boundscheck/wraparound/initializedcheck=False (setup.py:208-214).
This script does NOT run the reference's setup.py.  It stages the `.pyx/.pxd`
files from where they lie under /root/reference into a throw-away temp
directory (never into the repo), runs `cython` + `gcc` on them directly, and
writes ONLY the resulting `.so` files to

    oracle/_ref/f32/gprMax/...   float32 build (reference default)
    oracle/_ref/f64/gprMax/...   float64 build (constants.pxd:29-30 switched,
                                 exactly the edit the reference documents)

`oracle/_ref/` is git-ignored (binaries stay out of history) but travels to the
GPU box, where /root/reference does not exist.  -march is x86-64-v3 instead of
`native` so the binaries run on whatever host CPU the GPU box has.

Two groups are built:
  hot  - the solver kernels (needed on the GPU box for the CPU baseline)
  full - additionally the geometry/fractal extensions, only needed here to
         import the whole reference for golden-vector generation
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

import numpy as np

REF = os.environ.get('GPRMAX_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')

HOT = [
    'gprMax/fields_updates_ext.pyx',
    'gprMax/snapshots_ext.pyx',
    'gprMax/pml_updates/pml_updates_electric_HORIPML_ext.pyx',
    'gprMax/pml_updates/pml_updates_magnetic_HORIPML_ext.pyx',
    'gprMax/pml_updates/pml_updates_electric_MRIPML_ext.pyx',
    'gprMax/pml_updates/pml_updates_magnetic_MRIPML_ext.pyx',
]
REST = [
    'gprMax/yee_cell_build_ext.pyx',
    'gprMax/yee_cell_setget_rigid_ext.pyx',
    'gprMax/geometry_primitives_ext.pyx',
    'gprMax/fractals_generate_ext.pyx',
    'gprMax/geometry_outputs_ext.pyx',
]
PXD = ['gprMax/constants.pxd', 'gprMax/yee_cell_setget_rigid_ext.pxd']

CFLAGS = ['-O3', '-w', '-fopenmp', '-march=x86-64-v3', '-fPIC',
          '-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION']
DIRECTIVES = ['boundscheck=False', 'wraparound=False', 'initializedcheck=False',
              'embedsignature=True', 'language_level=3']


def _to_f64(text):
    """Apply the reference's documented precision switch (constants.py:36-49,
    constants.pxd:25-30): comment the single-precision lines, enable double."""
    out = []
    for line in text.splitlines():
        s = line.strip()
        if s in ('floattype = np.float32', 'complextype = np.complex64',
                 "cudafloattype = 'float'", "cudacomplextype = 'pycuda::complex<float>'",
                 'ctypedef np.float32_t floattype_t', 'ctypedef np.complex64_t complextype_t'):
            out.append('# ' + line)
        elif s.startswith('# floattype = np.float64') or s.startswith('# complextype = np.complex128') \
                or s.startswith("# cudafloattype = 'double'") or s.startswith("# cudacomplextype = 'pycuda::complex<double>'") \
                or s.startswith('# ctypedef np.float64_t') or s.startswith('# ctypedef np.complex128_t'):
            out.append(line.replace('# ', '', 1))
        else:
            out.append(line)
    return '\n'.join(out) + '\n'


def build(variant, files, verbose=True):
    outdir = os.path.join(OUT, variant)
    ext_suffix = sysconfig.get_config_var('EXT_SUFFIX')
    todo = []
    for f in files:
        so = os.path.join(outdir, os.path.splitext(f)[0] + ext_suffix)
        if not os.path.exists(so):
            todo.append((f, so))
    if not todo:
        return outdir
    if not os.path.isdir(REF):
        raise RuntimeError('reference tree {} not present and oracle/_ref/{} incomplete'.format(REF, variant))
    stage = tempfile.mkdtemp(prefix='gprmax_ref_stage_')
    try:
        os.makedirs(os.path.join(stage, 'gprMax', 'pml_updates'))
        for pkg in ('gprMax', 'gprMax/pml_updates'):
            open(os.path.join(stage, pkg, '__init__.py'), 'w').close()
        for f in PXD + [f for f, _ in todo]:
            with open(os.path.join(REF, f)) as fh:
                text = fh.read()
            if variant == 'f64' and f.endswith('constants.pxd'):
                text = _to_f64(text)
            with open(os.path.join(stage, f), 'w') as fh:
                fh.write(text)
        inc = ['-I' + sysconfig.get_paths()['include'], '-I' + np.get_include()]
        procs = []
        for f, so in todo:
            os.makedirs(os.path.dirname(so), exist_ok=True)
            cfile = os.path.join(stage, os.path.splitext(f)[0] + '.c')
            cmd = [sys.executable, '-m', 'cython', '-3']
            for d in DIRECTIVES:
                cmd += ['-X', d]
            cmd += ['-I', stage, '-o', cfile, os.path.join(stage, f)]
            subprocess.run(cmd, check=True, cwd=stage, stdout=subprocess.DEVNULL)
            gcc = ['/usr/bin/gcc'] + CFLAGS + inc + ['-shared', '-o', so, cfile, '-fopenmp']
            procs.append((f, subprocess.Popen(gcc)))
        for f, p in procs:
            if p.wait() != 0:
                raise RuntimeError('gcc failed for ' + f)
            if verbose:
                print('built oracle/_ref/{}/{}'.format(variant, os.path.splitext(f)[0]))
        for pkg in ('gprMax', 'gprMax/pml_updates'):
            os.makedirs(os.path.join(outdir, pkg), exist_ok=True)
        if variant == 'f64':
            # python-side precision switch used when the whole reference is imported
            with open(os.path.join(REF, 'gprMax/constants.py')) as fh:
                text = _to_f64(fh.read())
            with open(os.path.join(outdir, 'gprMax', 'constants.py'), 'w') as fh:
                fh.write(text)
    finally:
        shutil.rmtree(stage, ignore_errors=True)
    return outdir


def main(argv):
    full = '--full' in argv
    variants = [a for a in argv if a in ('f32', 'f64')] or ['f32', 'f64']
    for v in variants:
        build(v, HOT + (REST if full else []))


if __name__ == '__main__':
    main(sys.argv[1:])
