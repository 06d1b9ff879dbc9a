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
There is no collective on the data path.  Two transports move the planes:

  'p2p'   (default on GPUs) the shards are LINKED (gpb_link): every rank maps its neighbours' field arrays (CUDA IPC) and
          the library pushes the boundary planes into the neighbour's ghost plane with peer stores over NVLink, announced by
          flags in peer memory; a rank's whole iteration is one CUDA graph and `Solver.run(n)` advances n iterations with
          no host round trip at all;
  'nccl'  host-driven: the boundary plane is computed first (gpb_half_step part 0), its pairwise isend/irecv is posted over
          torch.distributed (NCCL on GPUs, gloo in the CPU tests), and the interior (part 1) runs while the planes travel.

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
        # (ID_local None and G.ID global: the library reads the slab straight out of G.ID, solver.PackedModel)
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


def link_neighbours(solver, rank, world, group=None):
    """Exchange gpb_link_t records between the ranks and link `solver` to its x-neighbours (transport 'p2p')."""
    import torch.distributed as dist
    infos = [None] * world
    dist.all_gather_object(infos, solver.link_info(), group=group)
    solver.link(left=infos[rank - 1] if rank > 0 else None, right=infos[rank + 1] if rank < world - 1 else None)


def build_id_slab(G, solid, rigidE, rigidH, ID_local, group=None, create_electric_average=None, create_magnetic_average=None):
    """Sharded host build of the per-edge material IDs (yee_cell_build_ext.pyx:110-257): this rank fills
    `ID_local` = uint32[6][its node planes][ny+1][nz+1] (as the geometry commands left it: rigid edges set) from ITS slab of the
    geometry only -- `solid` uint32 / `rigidE` int8[12] / `rigidH` int8[6] holding the cell planes [max(x0 - 1, 0), min(x1, nx))
    for its node planes [x0, x1) = partition_planes(G.nx, world)[rank].  The distinct material combinations of all ranks are
    exchanged with all_gather_object and resolved in the reference's order on every rank, so `G.materials` ends identical on
    all ranks and equal to the single-process build (yee_build.build_slab).  The result goes to `solve_gpu_sharded(G,
    ID_local=ID_local)` after the material coefficients have been computed from `G.materials`."""
    import types
    import torch.distributed as dist
    from . import yee_build
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    x0, n = partition_planes(G.nx, world)[rank]

    def gather(mine):
        out = [None] * world
        dist.all_gather_object(out, mine, group=group)
        return out
    slab = types.SimpleNamespace(nx=G.nx, ny=G.ny, nz=G.nz, solid=solid, rigidE=rigidE, rigidH=rigidH, ID=ID_local, materials=G.materials)
    return yee_build.build_slab(slab, (x0, x0 + n), max(x0 - 1, 0), x0, gather if world > 1 else None, create_electric_average, create_magnetic_average)


def solve_gpu_sharded(G, iterations=None, overlap=True, ID_local=None, timing=None, transport=None, results=None):
    """Run `G` sharded over all ranks of the default process group (call under torchrun).
    Returns (rxs, seconds): the R[9][iterations][nrx] receiver array (identical on every rank) and
    the loop time (max over ranks, device-timed).  `results` (a dict) additionally receives the gathered snapshots
    ('snapshots': [6 arrays each]) and transmission-line totals ('tlines': [(V, I)])."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    transport = transport or os.environ.get('GPB_SHARD_TRANSPORT', 'p2p')
    shard = GpuShard(G, rank, world, local, ID_local=ID_local)
    nit = int(G.iterations if iterations is None else iterations)
    dev = shard.device
    if transport == 'p2p':
        link_neighbours(shard.solver, rank, world)
        torch.cuda.synchronize()
        dist.barrier()
        shard.solver.run(nit)                      # graph-replayed iterations, halo pushed over peer memory
        seconds = torch.tensor([shard.solver.elapsed], device=dev, dtype=torch.float64)
    else:
        halo = HaloExchange(rank, world)
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
            seconds = torch.tensor([ev0.elapsed_time(ev1) * 1e-3], device=dev, dtype=torch.float64)
    torch.cuda.synchronize()
    dist.all_reduce(seconds, op=dist.ReduceOp.MAX)

    def gather(a):
        # owned by exactly one rank and zero elsewhere: the sum is exact
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()

    out = gather(shard.solver.receivers())
    if results is not None:
        results['snapshots'] = [[gather(c) for c in shard.solver.snapshot(n)] for n in range(len(G.snapshots))]
        results['tlines'] = [tuple(gather(v) for v in shard.solver.tline(n)) for n in range(len(G.transmissionlines))]
    seconds_host = float(seconds.item())
    torch.cuda.synchronize()
    dist.barrier()                                 # nobody unlinks while a neighbour may still push into its planes
    if transport == 'p2p':
        shard.solver.link()                        # drop my mappings of the neighbours' arrays ...
        dist.barrier()                             # ... before anybody frees them
    if timing is not None:
        timing['launches'] = shard.solver.kernel_launches
        timing['mem'] = shard.solver.mem_used
        timing['transport'] = transport
    shard.close()
    return out, seconds_host
