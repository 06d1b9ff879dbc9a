"""bench.py --gpus N (N > 1, launched by torchrun, one rank per GPU): the x-slab sharded weak-scaling run.

Workload: the synthetic lossy-dielectric domain, 256 x 2048 x 1024 cells PER GPU (N = 8 is BASELINE.json
configs[4], 2048 x 2048 x 1024 = 4.3 G cells), z-directed Hertzian dipole at the centre, 10-cell HORIPML, one receiver on a
cut plane.  Per rank and run:

  1. bit-exactness gate: the 512 x 512 x 256 cut of the same recipe (heterogeneous: a random mix of materials on every
     edge) is run sharded over the N ranks AND on one GPU by every rank; every rank compares its slab of all six final field
     arrays and the receiver traces bit for bit.  `bitexact` goes into the JSON line and a mismatch fails the run;
  2. `value`: device-timed iterations of the linked shards (gpb_run: one CUDA graph per iteration and rank, halo planes
     pushed over peer memory), barrier + synchronize on both sides, max over ranks;
  3. `e2e`: the public sharded call from HOST arrays, per-cell material IDs included: every rank uploads its slab of a
     heterogeneous uint32 ID array (two dielectrics in layers), runs, and the traces are gathered.
"""
import json
import os
import sys
import time

import numpy as np

from gprmax_b200.sharded import GpuShard, HaloExchange, link_neighbours, partition_planes, run_sharded, solve_gpu_sharded


def _heterogeneous(G, seed=7):
    """Replace the homogeneous fill of a benchkit.synthetic model by a random mix of five dielectrics on every edge."""
    from benchkit.synthetic import material_rows
    real = G.updatecoeffsE.dtype
    rows = [material_rows(er, se, 1.0, 0.0, G.dx, G.dy, G.dz, G.dt, real) for er, se in ((3.0, 0.001), (9.0, 0.02), (4.5, 0.0), (12.0, 0.05))]
    G.updatecoeffsE = np.concatenate([G.updatecoeffsE, np.stack([r[0] for r in rows])])
    G.updatecoeffsH = np.concatenate([G.updatecoeffsH, np.stack([r[1] for r in rows])])
    rng = np.random.default_rng(seed)
    G.ID = rng.integers(2, G.updatecoeffsE.shape[0], size=G.ID.shape, dtype=np.uint32)
    return G


def bitexact_gate(rank, world, local, transport, dims=(512, 512, 256), iterations=40):
    """Sharded over `world` ranks vs one GPU, same heterogeneous model: (ok on every rank, detail)."""
    import torch
    import torch.distributed as dist
    from benchkit.synthetic import homogeneous_model
    from gprmax_b200 import Solver
    nx, ny, nz = dims
    cut = partition_planes(nx, world)[world // 2][0]
    G = homogeneous_model(dims, iterations=iterations, er=6.0, se=0.01, src=((cut + 3) * 1e-3, ny // 2 * 1e-3, nz // 2 * 1e-3), src_pol='z',
                          rxs=[(cut * 1e-3, (ny // 2 + 6) * 1e-3, nz // 2 * 1e-3), ((cut - 5) * 1e-3, (ny // 2 - 4) * 1e-3, (nz // 2 + 3) * 1e-3)])
    _heterogeneous(G)
    with Solver(G, device_id=local) as sv:
        sv.run()
        ref_rx = sv.receivers()
        x0, n = partition_planes(nx, world)[rank]
        ref_fields = [sv.get_field(c)[x0:x0 + n].copy() for c in range(6)]
        single_path = sv.kernel_path
    shard = GpuShard(G, rank, world, local)
    if transport == 'p2p':
        link_neighbours(shard.solver, rank, world)
        torch.cuda.synchronize()
        dist.barrier()
        shard.solver.run(iterations)
    else:
        halo = HaloExchange(rank, world)
        with torch.cuda.stream(shard.stream):
            run_sharded(shard, halo, iterations)
        torch.cuda.synchronize()
    same_fields = all(np.array_equal(shard.solver.get_field(c), ref_fields[c]) for c in range(6))
    rx = torch.from_numpy(shard.solver.receivers()).to(shard.device)
    dist.all_reduce(rx)
    same_rx = bool(np.array_equal(rx.cpu().numpy(), ref_rx)) and float(np.abs(ref_rx).max()) > 0
    ok = torch.tensor([1 if (same_fields and same_rx) else 0], device=shard.device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    torch.cuda.synchronize()
    dist.barrier()
    shard_path = shard.solver.kernel_path
    if transport == 'p2p':
        shard.solver.link()   # unmap the neighbours before anybody frees
        dist.barrier()
    shard.close()
    return bool(ok.item()), {'dims': list(dims), 'iterations': iterations, 'ranks': world, 'model': 'random mix of 5 dielectrics on every edge, z dipole 3 planes from the middle cut, '
                             'receiver on the cut plane', 'compared': 'all six final field arrays (every rank its slab) and the receiver traces, np.array_equal',
                             'kernels_single': single_path, 'kernels_shard': shard_path}


def bench_sharded(args):
    import torch
    import torch.distributed as dist
    from benchkit.synthetic import homogeneous_model, material_rows

    # stdout carries exactly one JSON line: whatever libraries print to file descriptor 1 ("NCCL version ...", NCCL_DEBUG=INFO
    # output) is sent to stderr, and the JSON line is written to the saved descriptor at the end
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    local = int(os.environ.get('LOCAL_RANK', os.environ.get('RANK', '0')))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    transport = os.environ.get('GPB_SHARD_TRANSPORT', 'p2p')
    per_gpu = int(os.environ.get('GPB_SHARD_PLANES', '256'))
    ny, nz = int(os.environ.get('GPB_SHARD_NY', '2048')), int(os.environ.get('GPB_SHARD_NZ', '1024'))
    nx = per_gpu * world
    iters = args.iters or 20
    total_its = iters * (args.warmup + args.steps)

    # ---- 1. bit-exactness gate
    gate_dims = tuple(int(v) for v in os.environ.get('GPB_GATE_DIMS', '512,512,256').split(','))
    bitexact, gate = (None, None)
    if world > 1 and not os.environ.get('GPB_NO_GATE'):
        bitexact, gate = bitexact_gate(rank, world, local, transport, dims=gate_dims)

    # ---- 2. device-timed leg.  z-directed Hertzian dipole at the centre; one receiver on a cut plane, one inside a slab
    cx, cy, cz = nx // 2, ny // 2, nz // 2
    x_start, nplanes = partition_planes(nx, world)[rank]
    cut = partition_planes(nx, world)[world // 2][0]   # first plane of the middle rank: its trace needs halo data
    G = homogeneous_model((nx, ny, nz), iterations=total_its, er=6.0, se=0.01, src=(cx * 1e-3, cy * 1e-3, cz * 1e-3), src_pol='z',
                          rxs=[(cut * 1e-3, (cy + 100) * 1e-3, cz * 1e-3), ((cx + 37) * 1e-3, (cy + 50) * 1e-3, cz * 1e-3)],
                          x_range=(x_start, nplanes), build_id=False)
    shard = GpuShard(G, rank, world, local)   # homogeneous: no host ID array, the library fills uniform_id
    plane_bytes = shard.solver.halo(0)[2]
    cells = nx * ny * nz
    times = []
    from benchkit.clocks import ClockSampler
    clocks = ClockSampler(local)
    clocks.__enter__()
    if transport == 'p2p':
        if world > 1:
            link_neighbours(shard.solver, rank, world)
        for s in range(args.warmup + args.steps):
            torch.cuda.synchronize()
            dist.barrier()
            e0 = shard.solver.elapsed
            shard.solver.run(iters)                       # CUDA events on the library's stream around the graph replays
            torch.cuda.synchronize()
            dist.barrier()
            t = torch.tensor([shard.solver.elapsed - e0], device=shard.device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if s >= args.warmup:
                times.append(float(t.item()))
    else:
        halo = HaloExchange(rank, world)
        with torch.cuda.stream(shard.stream):
            halo.wait(halo.post_h(*[shard._t[k] for k in ('send_h_a', 'send_h_b', 'recv_h_a', 'recv_h_b')]))
            for s in range(args.warmup + args.steps):
                torch.cuda.synchronize()
                dist.barrier()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                run_sharded(shard, halo, iters)
                ev1.record()
                ev1.synchronize()
                torch.cuda.synchronize()
                dist.barrier()
                t = torch.tensor([ev0.elapsed_time(ev1) * 1e-3], device=shard.device, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                if s >= args.warmup:
                    times.append(float(t.item()))
    clocks.__exit__()
    clk = clocks.summary()
    mhz = torch.tensor([clk['sm_mhz'] or 0.0], device=shard.device, dtype=torch.float64)
    dist.all_reduce(mhz, op=dist.ReduceOp.MIN)
    clk['sm_mhz_min_over_ranks'] = float(mhz.item())
    launches = torch.tensor([shard.solver.kernel_launches], device=shard.device, dtype=torch.int64)
    dist.all_reduce(launches)
    kpath = shard.solver.kernel_path
    torch.cuda.synchronize()
    dist.barrier()
    if transport == 'p2p' and world > 1:
        shard.solver.link()
        dist.barrier()
    shard.close()
    shard = None

    # ---- 3. end-to-end leg: the public sharded call from HOST arrays on every rank, per-cell material IDs included.
    # One call = one job: 500 iterations (the N = 1 benchmark's job is its model's own 1559 iterations; a run that lets a wave
    # cross this 2-metre domain would take ~3500), so that the one-off upload is weighed against a run, not against 20 steps.
    # Two dielectrics in 16-plane layers along z (IDs 2 and 3): every rank holds its slab of the uint32 ID array
    # [6][planes][ny+1][nz+1] (12.9 GB per rank at the default size) and the library uploads it plane by plane.
    e2e_iters = int(os.environ.get('GPB_E2E_ITERS', '500'))
    slab_id_bytes = 6 * nplanes * (ny + 1) * (nz + 1) * 4
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    hetero = avail > 2.5 * slab_id_bytes * min(world, 8) and not os.environ.get('GPB_E2E_HOMOGENEOUS')
    flag = torch.tensor([1 if hetero else 0], device=torch.device('cuda', local))
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    hetero = bool(flag.item())
    Ge = homogeneous_model((nx, ny, nz), iterations=e2e_iters, er=6.0, se=0.01, src=(cx * 1e-3, cy * 1e-3, cz * 1e-3), src_pol='z',
                           rxs=[(cut * 1e-3, (cy + 100) * 1e-3, cz * 1e-3), ((cx + 37) * 1e-3, (cy + 50) * 1e-3, cz * 1e-3)],
                           x_range=(x_start, nplanes), build_id=False)
    ID_local = None
    if hetero:
        rowE, rowH = material_rows(4.0, 0.005, 1.0, 0.0, Ge.dx, Ge.dy, Ge.dz, Ge.dt, Ge.updatecoeffsE.dtype)
        Ge.updatecoeffsE = np.concatenate([Ge.updatecoeffsE, rowE[None]])
        Ge.updatecoeffsH = np.concatenate([Ge.updatecoeffsH, rowH[None]])
        layer = (2 + ((np.arange(nz + 1) // 16) & 1)).astype(np.uint32)
        ID_local = np.empty((6, nplanes, ny + 1, nz + 1), dtype=np.uint32)
        ID_local[...] = layer
    e2e_t = []
    for s in range(1 + max(args.steps, 2)):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        rx_e2e, _ = solve_gpu_sharded(Ge, iterations=e2e_iters, ID_local=ID_local, transport=transport)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=torch.device('cuda', local), dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if s >= 1:
            e2e_t.append(float(dt.item()))
    e2e_value = cells * e2e_iters / (float(np.median(e2e_t)) * 1e6)
    real_bytes = np.dtype(Ge.updatecoeffsE.dtype).itemsize
    h2d = Ge.updatecoeffsE.nbytes + Ge.updatecoeffsH.nbytes + sum(8 * p.ERA.nbytes for p in Ge.pmls) \
        + sum(s_.waveformvalues_wholestep.nbytes for s_ in Ge.hertziandipoles) + 12 * len(Ge.rxs)
    h2d_total = h2d * world + (6 * (nx + 1) * (ny + 1) * (nz + 1) * 4 if hetero else 0)
    d2h = 9 * e2e_iters * len(Ge.rxs) * real_bytes * world
    t_step = float(np.mean(times))
    value = cells * iters / (t_step * 1e6)
    if rank == 0:
        S = 2 * 10 * (ny * nz + nx * nz + nx * ny)
        b_alg = 96.0 + 32.0 * S / cells
        peak = 6456.8
        try:
            with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')) as f:
                peak = float(json.load(f)['hbm_gbs'])
        except Exception:
            pass
        line = {
            'metric': 'FDTD throughput', 'value': value, 'unit': 'Mcells/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': t_step * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'synthetic {}x{}x{} lossy dielectric (er=6, sigma=0.01), x-slab sharded {} planes per GPU, z Hertzian dipole, '
                                   '10-cell HORIPML, one-plane Ey/Ez and Hy/Hz halo per half-step'.format(nx, ny, nz, per_gpu),
                       'cells': cells, 'iterations_per_step': iters, 'l2': 'working set per GPU >> 126 MB L2', 'alg_bytes_per_cell_step': b_alg,
                       'halo_bytes_per_interface_per_iteration': int(4 * plane_bytes), 'kernels': kpath,
                       'halo_transport': 'p2p: boundary planes pushed into the neighbour\'s ghost plane by k_halo_push (peer stores over NVLink, CUDA IPC mapping), '
                                         'flag-announced; one CUDA graph per iteration and rank, no host round trip, no collective' if transport == 'p2p'
                                         else 'nccl: pairwise isend/irecv from the host between the boundary and interior parts of each half-step'},
            'bitexact': bitexact, 'bitexact_gate': gate,
            'roofline': {'bound': 'hbm', 'kernel': 'whole step (all ranks)', 'achieved': value * 1e6 * b_alg / 1e9 / world, 'peak': peak, 'unit': 'GB/s',
                         'frac': value * 1e6 * b_alg / 1e9 / world / peak, 'traffic': None},
            'cpu_baseline': None,
            'e2e': {'value': e2e_value, 'unit': 'Mcells/s', 'h2d_bytes_per_step': int(h2d_total), 'd2h_bytes_per_step': int(d2h),
                    'call': 'gprmax_b200.sharded.solve_gpu_sharded(G, iterations={}, ID_local=<this rank\'s slab of the uint32 ID array>) on every rank: shard creation '
                            'from host tables, {} run, traces gathered'.format(e2e_iters, 'per-cell ID upload (two dielectrics in 16-cell layers),' if hetero else
                                                                                 'homogeneous device-side ID fill (host RAM too small for the per-cell arrays),'),
                    'iterations_per_call': e2e_iters, 'seconds_per_call': [round(t_, 4) for t_ in e2e_t], 'statistic': 'median'},
            'gpu_launches': int(launches.item()), 'clocks': clk,
        }
        sys.stdout.flush()
        os.write(out_fd, (json.dumps(line) + '\n').encode())
    dist.barrier()
    dist.destroy_process_group()
    return 0 if bitexact is not False else 1
