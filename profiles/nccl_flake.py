"""Repeat the NCCL sharded run of a fixture and compare with the single-GPU result bit for bit (flake hunt)."""
import os, subprocess, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from gprmax_b200 import Solver
from gprmax_b200.model_io import load_model
fixture = 'tests/golden/pml_HORIPML_2_f32.npz'
G, _ = load_model(fixture)
with Solver(G, device_id=0) as sv:
    sv.run(); ref = sv.receivers()
with Solver(G, device_id=0) as sv:
    sv.run(); ref2 = sv.receivers()
print('single-GPU run repeatable:', np.array_equal(ref, ref2), flush=True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for overlap in ('0', '1'):
    bad = 0
    for rep in range(n):
        out = '/tmp/flake_rx.npy'
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1', '--master-port', str(29600 + rep + 50 * int(overlap)),
               'tests/sharded_worker.py', fixture, out, overlap]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
        if r.returncode:
            print('worker failed', r.stdout[-500:]); continue
        rx = np.load(out)
        if not np.array_equal(rx, ref):
            bad += 1
            d = np.argwhere(rx != ref)
            print('overlap', overlap, 'rep', rep, 'MISMATCH', len(d), 'values; rows', sorted(set(d[:, 0].tolist())), 'first iterations', sorted(set(d[:, 1].tolist()))[:5], 'rx idx', sorted(set(d[:, 2].tolist())),
                  'max rel', float(np.abs(rx - ref).max() / np.abs(ref).max()), flush=True)
    print('overlap', overlap, ':', bad, 'of', n, 'runs differ', flush=True)
