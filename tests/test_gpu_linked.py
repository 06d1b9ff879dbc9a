"""Linked x-slab shards (gpb_link / gpb_create_sharded): halo planes pushed into the neighbour's ghost plane by peer stores,
flag-synchronised, one CUDA graph per slab and iteration -- no host round trip.  Every run must reproduce the single-GPU
result BIT FOR BIT, including what the host-driven path cannot do: snapshots that span a cut plane and transmission lines on
a slab's first / last plane (reference semantics snapshots.py:87-130, sources.py:426-452).

  * `same`: all slabs on device 0 of one process (runs on a 1-GPU box and exercises the whole flag protocol);
  * `all`:  one slab per visible device (needs >= 2 GPUs), in-process peer access;
  * torchrun: one process per GPU, neighbours mapped through CUDA IPC.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, golden_path

pytestmark = pytest.mark.gpu

CASES = [
    # fixture, slabs, options
    ('pml_HORIPML_2_f32', 2, ['its=150']),
    ('pml_MRIPML_2_f64', 3, ['its=150']),
    ('sources_mixed_f32', 4, ['its=200']),
    ('hertzian_dipole_dispersive_f32', 3, ['its=150']),
    ('dispersive_multipole_f64', 2, ['tma', 'its=150']),
    ('bench_100_f32', 5, ['tma']),
    ('snapshots_f32', 2, []),                 # both snapshots span the cut
    ('snapshots_f64', 3, ['tma']),
    ('transmission_line_f32', 2, []),         # 41 planes -> [0,21) [21,41): the line (x = 20) sits on the LAST plane of slab 0
    ('transmission_line_f64', 2, ['tl=21']),  # ... and on the FIRST plane of slab 1 (reads H of the ghost plane)
    ('transmission_line_f32', 4, ['tl=11', 'tma']),
]


def _run(args, timeout=600):
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS='32', GPB_LINK_TIMEOUT_MS='8000')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'linked_worker.py')] + [str(a) for a in args], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, universal_newlines=True, timeout=timeout, env=env)
    assert r.returncode == 0 and 'LINKED_OK' in r.stdout, r.stdout[-3000:]
    return r.stdout


@pytest.mark.parametrize('fixture,nslabs,opts', CASES)
def test_linked_slabs_on_one_device_bit_exact(fixture, nslabs, opts):
    _run([os.path.join(ROOT, 'tests', 'golden', fixture + '.npz'), nslabs, 'same'] + opts)


def test_linked_slabs_tma_sized_grid_bit_exact():
    """The sharded benchmark's recipe scaled down, large enough for many persistent-CTA work items per slab."""
    _run(['synthetic:160,144,128,120', 3, 'same'])


@pytest.mark.parametrize('fixture,nslabs,opts', [CASES[0], CASES[6], CASES[9], ('synthetic:160,144,128,120', 0, [])])
def test_linked_slabs_across_devices_bit_exact(fixture, nslabs, opts):
    from gprmax_b200.gpu import device_count
    n = device_count()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    spec = fixture if fixture.startswith('synthetic') else os.path.join(ROOT, 'tests', 'golden', fixture + '.npz')
    _run([spec, min(n, 8) if nslabs == 0 else max(nslabs, 2), 'all'] + opts)


@pytest.mark.parametrize('spec', ['pml_HORIPML_2_f32', 'snapshots_f32', 'synthetic:160,144,128,120', 'synthetic_cut:64,48,40,90'])
def test_torchrun_p2p_shards_bit_exact(spec, tmp_path):
    """One process per GPU; the neighbours' arrays are mapped with CUDA IPC and the halo is pushed by the library."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from sharded_worker import build
    from gprmax_b200 import Solver
    from gprmax_b200.gpu import device_count
    n = device_count()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    world = min(n, 4)
    fixture = spec if spec.startswith('synthetic') else os.path.join(ROOT, 'tests', 'golden', spec + '.npz')
    G = build(fixture)
    with Solver(G, device_id=0) as sv:
        sv.run()
        ref_rx = sv.receivers()
        ref_snaps = [sv.snapshot(k) for k in range(len(G.snapshots))]
    out = str(tmp_path / 'rx.npy')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
           '--master-port', '29541', os.path.join(ROOT, 'tests', 'sharded_worker.py'), fixture, out, '1', 'p2p']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True, timeout=600,
                       env=dict(os.environ, GPB_LINK_TIMEOUT_MS='8000'))
    assert r.returncode == 0, r.stdout[-3000:]
    assert np.array_equal(np.load(out), ref_rx)
    if G.snapshots:
        z = np.load(out + '.snaps.npz')
        for k, sn in enumerate(ref_snaps):
            for c in range(6):
                assert np.array_equal(z['s{}_{}'.format(k, c)], sn[c])
