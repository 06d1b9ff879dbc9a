"""Repeat sharded runs of a model inside ONE torchrun launch and compare every repetition with the single-GPU result bit for bit
(race hunt: round 1 found a 1-in-20 stream-ordering race this way).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 \
        profiles/shard_flake.py [reps] [fixture ...]
For every fixture and each transport (p2p = linked shards pushing halos over peer memory, nccl with and without the
boundary-first overlap) it prints `mismatches / reps`; exit code 1 if any repetition differs."""
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import torch
import torch.distributed as dist
from gprmax_b200 import Solver
from gprmax_b200.sharded import solve_gpu_sharded
from sharded_worker import build

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
fixtures = sys.argv[2:] or ['tests/golden/pml_HORIPML_2_f32.npz', 'synthetic:160,144,128,120', 'synthetic_cut:64,48,40,90']
dist.init_process_group('nccl')
rank, world = dist.get_rank(), dist.get_world_size()
local = int(os.environ.get('LOCAL_RANK', rank))
torch.cuda.set_device(local)
bad_total = 0
for fx in fixtures:
    G = build(fx)
    with Solver(G, device_id=local) as sv:
        sv.run()
        ref = sv.receivers()
    for transport, overlap in (('p2p', True), ('nccl', True), ('nccl', False)):
        bad = 0
        for rep in range(reps):
            rx, _ = solve_gpu_sharded(G, overlap=overlap, transport=transport)
            if not np.array_equal(rx, ref):
                bad += 1
                d = np.argwhere(rx != ref)
                if rank == 0:
                    print('  {} {} overlap={} rep {}: MISMATCH {} values, first iterations {}'.format(fx, transport, overlap, rep, len(d), sorted(set(d[:, 1].tolist()))[:5]), flush=True)
        bad_total += bad
        if rank == 0:
            print('{} ranks, {}: transport {} overlap {}: {} mismatches / {} repetitions'.format(world, fx, transport, overlap, bad, reps), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if bad_total else 0)
