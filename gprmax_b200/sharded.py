"""x-slab sharding of one FDTD domain over the GPUs of a node (one process per GPU).

The reference has no domain decomposition (a model larger than one GPU is rejected,
grid.py:239-241); this is the new sharded mode SURVEY.md section 8(e) describes.

Partition: the nx+1 node planes are cut into `world` contiguous x-ranges; rank r owns planes
[x0_r, x1_r) of all six components.  x is the slowest axis, so a plane is one contiguous block.
Dependencies across a cut (same as the single-GPU kernels, just across devices):
    H half-step of plane x1-1 reads Ey,Ez of plane x1   (owned by the right neighbour)
    E half-step of plane x0   reads Hy,Hz of plane x0-1 (owned by the left neighbour)
so per half-step each interface moves TWO planes in ONE direction:
    after the H half-step:  Hy,Hz of my last  plane -> right neighbour's ghost plane x0-1
    after the E half-step:  Ey,Ez of my first plane -> left  neighbour's ghost plane x1
There is no collective on the data path, only pairwise send/recv (NCCL on GPUs, gloo in the CPU
tests).  The boundary plane is computed first (gpb_half_step part 0), its send is posted, and the
interior (part 1) runs while the planes travel over NVLink.

Every plane is advanced by the same kernels with the same operands as in a single-GPU run, so a
sharded run reproduces the single-GPU result bit for bit (tests/test_gpu_sharded.py).

The `engine` argument of `run_sharded` is anything with the small interface of
`gprmax_b200.solver.Solver` (half_step(phase, part), halo tensors); the CPU test drives the same
exchange protocol with an oracle-backed engine over gloo.
"""
import json
import os
import sys
import time

import numpy as np


def partition_planes(nx, world):
    """Contiguous, balanced ranges of the nx+1 node planes: [(x_start, nx_planes)] per rank."""
    total = nx + 1
    if world < 1 or world > total:
        raise ValueError('cannot split {} planes over {} ranks'.format(total, world))
    base, extra = divmod(total, world)
    out, x = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((x, n))
        x += n
    return out


class HaloExchange(object):
    """Pairwise plane exchange with the x-neighbours over torch.distributed (nccl or gloo)."""

    def __init__(self, rank, world, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world, self.group = rank, world, group
        self.left = rank - 1 if rank > 0 else None
        self.right = rank + 1 if rank < world - 1 else None

    def _run(self, ops):
        if not ops:
            return []
        return self.dist.batch_isend_irecv(ops)

    def post_h(self, send_a, send_b, recv_a, recv_b):
        """Hy,Hz: my last owned plane -> right neighbour; left neighbour's -> my ghost plane x0-1."""
        P = self.dist.P2POp
        ops = []
        if self.right is not None:
            ops += [P(self.dist.isend, send_a, self.right, self.group), P(self.dist.isend, send_b, self.right, self.group)]
        if self.left is not None:
            ops += [P(self.dist.irecv, recv_a, self.left, self.group), P(self.dist.irecv, recv_b, self.left, self.group)]
        return self._run(ops)

    def post_e(self, send_a, send_b, recv_a, recv_b):
        """Ey,Ez: my first owned plane -> left neighbour; right neighbour's -> my ghost plane x1."""
        P = self.dist.P2POp
        ops = []
        if self.left is not None:
            ops += [P(self.dist.isend, send_a, self.left, self.group), P(self.dist.isend, send_b, self.left, self.group)]
        if self.right is not None:
            ops += [P(self.dist.irecv, recv_a, self.right, self.group), P(self.dist.irecv, recv_b, self.right, self.group)]
        return self._run(ops)

    @staticmethod
    def wait(reqs):
        for r in reqs:
            r.wait()


def run_sharded(engine, halo, iterations, overlap=True):
    """Advance `iterations` full time steps of one shard.  `engine.halo_tensors()` returns the eight
    plane tensors {send,recv} x {E,H} x {a,b}; all engine work and all exchanges are enqueued on the
    engine's stream (GPU) or run eagerly (CPU test)."""
    t = engine.halo_tensors()
    for _ in range(iterations):
        # ---- H half-step
        engine.half_step(0, 0)                      # prologue + last owned plane
        reqs = halo.post_h(t['send_h_a'], t['send_h_b'], t['recv_h_a'], t['recv_h_b']) if overlap else []
        engine.half_step(0, 1)                      # interior while the planes travel
        if not overlap:
            reqs = halo.post_h(t['send_h_a'], t['send_h_b'], t['recv_h_a'], t['recv_h_b'])
        halo.wait(reqs)
        # ---- E half-step
        engine.half_step(1, 0)                      # first owned plane
        reqs = halo.post_e(t['send_e_a'], t['send_e_b'], t['recv_e_a'], t['recv_e_b']) if overlap else []
        engine.half_step(1, 1)
        if not overlap:
            reqs = halo.post_e(t['send_e_a'], t['send_e_b'], t['recv_e_a'], t['recv_e_b'])
        halo.wait(reqs)


class _DevPlane(object):
    """Device memory of the library exposed through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr, nelem, typestr):
        self.__cuda_array_interface__ = {'shape': (nelem,), 'typestr': typestr, 'data': (ptr, False), 'version': 3, 'strides': None}


class GpuShard(object):
    """One x-slab on one GPU: a `Solver` restricted to its planes + torch views of its halo planes."""

    def __init__(self, G, rank, world, device_id, ID_local=None):
        import torch
        from .solver import Solver
        self.torch = torch
        self.rank, self.world = rank, world
        self.x_start, self.nx_planes = partition_planes(G.nx, world)[rank]
        if ID_local is None and getattr(G, 'ID', None) is not None:
            ID_local = np.ascontiguousarray(G.ID[:, self.x_start:self.x_start + self.nx_planes])
        self.solver = Solver(G, device_id=device_id, x_start=self.x_start, nx_planes=self.nx_planes, ID=ID_local)
        self.device = torch.device('cuda', device_id)
        self.stream = torch.cuda.ExternalStream(self.solver.stream, device=self.device)
        real = np.dtype(self.solver.real)
        typestr = '<f4' if real == np.float32 else '<f8'
        self._t = {}
        for which, name in ((0, 'send_e'), (1, 'recv_e'), (2, 'send_h'), (3, 'recv_h')):
            a, b, nbytes = self.solver.halo(which)
            n = nbytes // real.itemsize
            self._t[name + '_a'] = torch.as_tensor(_DevPlane(a, n, typestr), device=self.device)
            self._t[name + '_b'] = torch.as_tensor(_DevPlane(b, n, typestr), device=self.device)

    def halo_tensors(self):
        return self._t

    def half_step(self, phase, part):
        self.solver.half_step(phase, part)

    def close(self):
        self.solver.close()


def run_sharded_local(shards, iterations):
    """Single process driving several shards (on one or several visible devices): the halo planes
    are moved with plain device copies.  Same per-plane kernels as the multi-process path; used to
    check bit-exactness of the sharding on a box with a single GPU."""
    import torch
    def sync():
        for s in shards:
            s.solver.synchronize()
    for _ in range(iterations):
        for s in shards:
            s.half_step(0, 0)
        for s in shards:
            s.half_step(0, 1)
        sync()
        for left, right in zip(shards, shards[1:]):
            right._t['recv_h_a'].copy_(left._t['send_h_a'])
            right._t['recv_h_b'].copy_(left._t['send_h_b'])
        torch.cuda.synchronize()
        for s in shards:
            s.half_step(1, 0)
        for s in shards:
            s.half_step(1, 1)
        sync()
        for left, right in zip(shards, shards[1:]):
            left._t['recv_e_a'].copy_(right._t['send_e_a'])
            left._t['recv_e_b'].copy_(right._t['send_e_b'])
        torch.cuda.synchronize()


def solve_gpu_sharded(G, iterations=None, overlap=True, ID_local=None, timing=None):
    """Run `G` sharded over all ranks of the default process group (call under torchrun).
    Returns (rxs, seconds): the R[9][iterations][nrx] receiver array (identical on every rank) and
    the loop time (max over ranks, device-timed)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    shard = GpuShard(G, rank, world, local, ID_local=ID_local)
    halo = HaloExchange(rank, world)
    nit = int(G.iterations if iterations is None else iterations)
    with torch.cuda.stream(shard.stream):
        # one untimed exchange sets up the NCCL channels
        halo.wait(halo.post_h(*[shard._t[k] for k in ('send_h_a', 'send_h_b', 'recv_h_a', 'recv_h_b')]))
        torch.cuda.synchronize()
        dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        run_sharded(shard, halo, nit, overlap=overlap)
        ev1.record()
        ev1.synchronize()
        seconds = torch.tensor([ev0.elapsed_time(ev1) * 1e-3], device=shard.device)
        dist.all_reduce(seconds, op=dist.ReduceOp.MAX)
        rxs = torch.from_numpy(shard.solver.receivers()).to(shard.device)
        # a receiver is owned by exactly one rank and zero elsewhere: the sum is exact
        dist.all_reduce(rxs, op=dist.ReduceOp.SUM)
        # read back on the SAME stream: the library's stream is non-blocking, so a copy on torch's default stream outside this
        # block is not ordered after the all-reduce (seen as a 1-in-20 flake: rank 0 stored its own receivers only)
        out = rxs.cpu().numpy()
        seconds_host = float(seconds.item())
    torch.cuda.synchronize()
    if timing is not None:
        timing['launches'] = shard.solver.kernel_launches
        timing['mem'] = shard.solver.mem_used
    shard.close()
    return out, seconds_host


# ------------------------------------------------------------------------------------------ bench
def bench_sharded(args):
    """bench.py --gpus N (N > 1), launched by torchrun: weak-scaling x-slab sharded run of the synthetic
    homogeneous lossy-dielectric domain (BASELINE.json configs[4] at N = 8)."""
    import torch
    import torch.distributed as dist
    from benchkit.synthetic import homogeneous_model

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
    per_gpu = int(os.environ.get('GPB_SHARD_PLANES', '256'))
    ny, nz = int(os.environ.get('GPB_SHARD_NY', '2048')), int(os.environ.get('GPB_SHARD_NZ', '1024'))
    nx = per_gpu * world
    iters = args.iters or 20
    total_its = iters * (args.warmup + args.steps)
    # z-directed Hertzian dipole at the centre; one receiver on a cut plane, one inside a slab
    cx, cy, cz = nx // 2, ny // 2, nz // 2
    x_start, nplanes = partition_planes(nx, world)[rank]
    cut = partition_planes(nx, world)[world // 2][0]   # first plane of the middle rank: its trace needs halo data
    G = homogeneous_model((nx, ny, nz), iterations=total_its, er=6.0, se=0.01, src=(cx * 1e-3, cy * 1e-3, cz * 1e-3), src_pol='z',
                          rxs=[(cut * 1e-3, (cy + 100) * 1e-3, cz * 1e-3), ((cx + 37) * 1e-3, (cy + 50) * 1e-3, cz * 1e-3)],
                          x_range=(x_start, nplanes), build_id=False)
    shard = GpuShard(G, rank, world, local)   # homogeneous: no host ID array, the library fills uniform_id
    halo = HaloExchange(rank, world)
    plane_bytes = shard.solver.halo(0)[2]
    cells = nx * ny * nz
    times = []
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
            t = torch.tensor([ev0.elapsed_time(ev1) * 1e-3], device=shard.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if s >= args.warmup:
                times.append(float(t.item()))
    launches = torch.tensor([shard.solver.kernel_launches], device=shard.device, dtype=torch.int64)
    dist.all_reduce(launches)
    # ---- end-to-end leg: the public sharded call from HOST tables -- every call creates the shard (coefficient / PML /
    # waveform tables host -> device, homogeneous ID fill on the device), runs `iters` iterations with halo exchange and
    # copies the receiver traces back; host wall clock between barriers, max over ranks.  (The domain is homogeneous, so
    # there is no per-cell host array to upload; the N = 1 benchmark is the one that moves a 654 MB ID array.)
    shard.close()
    shard = None
    e2e_t = []
    for s in range(1 + max(args.steps, 3)):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        rx_e2e, _ = solve_gpu_sharded(G, iterations=iters)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=torch.device('cuda', local), dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if s >= 1:
            e2e_t.append(float(dt.item()))
    e2e_value = cells * iters / (float(np.median(e2e_t)) * 1e6)
    real_bytes = np.dtype(G.updatecoeffsE.dtype).itemsize
    h2d = G.updatecoeffsE.nbytes + G.updatecoeffsH.nbytes + sum(8 * p.ERA.nbytes for p in G.pmls) \
        + sum(s_.waveformvalues_wholestep.nbytes for s_ in G.hertziandipoles) + 12 * len(G.rxs)
    d2h = 9 * total_its * len(G.rxs) * real_bytes
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
                                   '10-cell HORIPML, one-plane Ey/Ez and Hy/Hz halo per half-step over NCCL'.format(nx, ny, nz, per_gpu),
                       'cells': cells, 'iterations_per_step': iters, 'l2': 'working set per GPU >> 126 MB L2', 'alg_bytes_per_cell_step': b_alg,
                       'halo_bytes_per_interface_per_iteration': int(4 * plane_bytes)},
            'roofline': {'bound': 'hbm', 'kernel': 'whole step (all ranks)', 'achieved': value * 1e6 * b_alg / 1e9 / world, 'peak': peak, 'unit': 'GB/s',
                         'frac': value * 1e6 * b_alg / 1e9 / world / peak, 'traffic': None},
            'cpu_baseline': None,
            'e2e': {'value': e2e_value, 'unit': 'Mcells/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'call': 'gprmax_b200.sharded.solve_gpu_sharded(G, iterations) from host tables on every rank (homogeneous domain: '
                            'IDs are filled on the device)', 'seconds_per_call': [round(t_, 4) for t_ in e2e_t], 'statistic': 'median'},
            'gpu_launches': int(launches.item()),
        }
        sys.stdout.flush()
        os.write(out_fd, (json.dumps(line) + '\n').encode())
    dist.barrier()
    dist.destroy_process_group()
    return 0
