"""x-slab sharding on GPUs: every multi-shard run must reproduce the single-GPU result BIT FOR BIT
(north_star).  Two layers:
  * several shards driven by one process on one GPU (plain device copies for the halos) -- runs on
    any box, checks ownership / ghost planes / sliced PML slabs / per-shard sources and receivers;
  * one process per GPU over NCCL through torchrun (needs >= 2 GPUs), with and without the
    boundary-first overlap.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, golden_path

pytestmark = pytest.mark.gpu


def _single(G):
    from gprmax_b200 import Solver
    with Solver(G, device_id=0) as sv:
        sv.run()
        return sv.receivers(), [sv.get_field(c) for c in range(6)]


@pytest.mark.parametrize('name,variant,nshards,force_tma', [('pml_HORIPML_1', 'f32', 2, False), ('pml_MRIPML_2', 'f64', 3, False),
                                                            ('sources_mixed', 'f32', 4, False), ('hertzian_dipole_dispersive', 'f32', 3, False),
                                                            ('bench_100', 'f32', 5, False), ('pml_HORIPML_2', 'f32', 3, True),
                                                            ('bench_100', 'f32', 4, True), ('dispersive_multipole', 'f64', 2, True)])
def test_local_shards_bit_exact(name, variant, nshards, force_tma, monkeypatch):
    from gprmax_b200.model_io import load_model
    if force_tma:   # the TMA-staged kernels are normally only used above 2.5 M nodes
        monkeypatch.setenv('GPB_FORCE_TMA', '1')
    from gprmax_b200.sharded import GpuShard, run_sharded_local
    G, _ = load_model(golden_path(name, variant))
    G.iterations = min(G.iterations, 150)
    for lst in (G.hertziandipoles, G.magneticdipoles, G.voltagesources):
        for s in lst:
            s.waveformvalues_wholestep = s.waveformvalues_wholestep[:G.iterations]
            s.waveformvalues_halfstep = s.waveformvalues_halfstep[:G.iterations]
    ref_rx, ref_fields = _single(G)
    shards = [GpuShard(G, r, nshards, 0) for r in range(nshards)]
    run_sharded_local(shards, G.iterations)
    rx = sum(s.solver.receivers() for s in shards)
    # a receiver is owned by exactly one shard and zero in the others.  Iy/Iz on a shard's first plane
    # read the H ghost plane, which is exchanged before it is sampled, so all nine rows must agree.
    assert np.array_equal(rx, ref_rx)
    for c in range(6):
        full = np.concatenate([s.solver.get_field(c) for s in shards], axis=0)
        assert np.array_equal(full, ref_fields[c]), c
    for s in shards:
        s.close()


@pytest.mark.parametrize('overlap', [1, 0])
def test_nccl_shards_bit_exact(overlap, tmp_path):
    from gprmax_b200.gpu import device_count
    from gprmax_b200.model_io import load_model
    n = device_count()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    world = min(n, 4)
    fixture = golden_path('pml_HORIPML_2', 'f32')
    G, _ = load_model(fixture)
    ref_rx, _ = _single(G)
    out = str(tmp_path / 'rx.npy')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
           '--master-port', '29533', os.path.join(ROOT, 'tests', 'sharded_worker.py'), fixture, out, str(overlap)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert np.array_equal(np.load(out), ref_rx)


def test_nccl_shards_bit_exact_tma_path(tmp_path):
    """Same as above on a synthetic lossy-dielectric box large enough for the TMA-staged kernels (the sharded benchmark's
    recipe, scaled down): every rank runs boundary-plane-first launches of k_update_tma and exchanges halos over NCCL."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from sharded_worker import build
    from gprmax_b200.gpu import device_count
    n = device_count()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    world = min(n, 4)
    spec = 'synthetic:160,144,128,120'
    ref_rx, _ = _single(build(spec))
    assert np.abs(ref_rx).max() > 0
    out = str(tmp_path / 'rx.npy')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
           '--master-port', '29534', os.path.join(ROOT, 'tests', 'sharded_worker.py'), spec, out, '1']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert np.array_equal(np.load(out), ref_rx)


def test_nccl_source_on_cut_plane(tmp_path):
    """A Hertzian dipole on the first owned plane of a rank: with the boundary-first overlap that plane is sent to the left
    neighbour before the interior is updated, so the source term has to be applied to it first (gpb_half_step part 0)."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from sharded_worker import build
    from gprmax_b200.gpu import device_count
    n = device_count()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    world = 4 if n >= 4 else 2
    spec = 'synthetic_cut:64,48,40,90'
    ref_rx, _ = _single(build(spec))
    assert np.abs(ref_rx).max() > 0
    out = str(tmp_path / 'rx.npy')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
           '--master-port', '29535', os.path.join(ROOT, 'tests', 'sharded_worker.py'), spec, out, '1']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert np.array_equal(np.load(out), ref_rx)
