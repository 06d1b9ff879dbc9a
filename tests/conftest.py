import glob
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'needs_ref_kernels: needs oracle/_ref (the reference Cython kernels built by oracle/build_ref.py)')


def have_ref_kernels(variant='f32'):
    return bool(glob.glob(os.path.join(ROOT, 'oracle', '_ref', variant, 'gprMax', 'fields_updates_ext.*.so')))


def golden_path(name, variant='f32'):
    return os.path.join(GOLDEN, '{}_{}.npz'.format(name, variant))


# full-size fixtures of BASELINE.json configs 3 and 4: used by dedicated GPU tests only (minutes on the CPU oracle)
BIG = ('bscan_gssi_trace1', 'heterogeneous_soil_full', 'bench_300_trace')


def golden_names(variant='f32'):
    names = sorted(os.path.basename(p)[:-len('_{}.npz'.format(variant))] for p in glob.glob(os.path.join(GOLDEN, '*_{}.npz'.format(variant))))
    return [n for n in names if n not in BIG]


@pytest.fixture(scope='session')
def oracle_built():
    from oracle.solver import build_oracle
    return build_oracle()
