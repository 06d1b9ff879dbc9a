"""Host-side logic of the x-slab sharded mode, on CPU with world_size 2 and 3 over gloo.

The halo protocol of gprmax_b200/sharded.py (which planes move, in which direction, at which point
of the step, with the boundary plane computed first) is exercised with an oracle-backed engine:
every rank holds full-size arrays but after each half-step everything it does not own is poisoned
with NaN, and only the planes the protocol delivers are written into its ghost planes.  If the
protocol were wrong (a missing plane, a wrong direction, a stale halo) NaNs or wrong values would
reach the owned planes.  The receiver traces gathered from the owning ranks must equal a plain
single-domain oracle run BIT FOR BIT.
"""
import os
import socket

import numpy as np
import pytest

from conftest import golden_path
from gprmax_b200.sharded import HaloExchange, partition_planes, run_sharded


def test_partition_planes():
    for nx, world in ((300, 1), (300, 2), (300, 7), (2048, 8), (5, 6)):
        parts = partition_planes(nx, world)
        assert len(parts) == world
        assert parts[0][0] == 0 and parts[-1][0] + parts[-1][1] == nx + 1
        for (a, n), (b, _) in zip(parts, parts[1:]):
            assert a + n == b
        sizes = [n for _, n in parts]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        partition_planes(3, 5)


class OracleShard(object):
    """Engine with the Solver's sharded interface, backed by the CPU oracle on full-size arrays."""

    def __init__(self, G, rank, world):
        import torch
        from oracle.solver import OracleKernels, State, _hertzian
        self.torch = torch
        self.G = G
        self.K = OracleKernels(G.updatecoeffsE.dtype)
        self.S = State(G, self.K)
        self.hertz = _hertzian
        self.x0, n = partition_planes(G.nx, world)[rank]
        self.x1 = self.x0 + n
        self.rank, self.world = rank, world
        self.it = 0
        plane = (G.ny + 1) * (G.nz + 1)
        z = lambda: torch.zeros(plane, dtype=torch.float32 if self.S.real == np.float32 else torch.float64)
        self.t = {k: z() for k in ('send_e_a', 'send_e_b', 'recv_e_a', 'recv_e_b', 'send_h_a', 'send_h_b', 'recv_h_a', 'recv_h_b')}
        self.rx = [dict((k, np.zeros(G.iterations, dtype=self.S.real)) for k in ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz')) for _ in G.rxs]
        self.pending = None   # 'h' / 'e': ghost planes received during the previous half-step

    def halo_tensors(self):
        return self.t

    def _owned(self, i):
        return self.x0 <= i < self.x1

    def _apply_ghosts(self):
        S = self.S
        if self.pending == 'h' and self.x0 > 0:
            S.Hy[self.x0 - 1] = self.t['recv_h_a'].numpy().reshape(S.Hy.shape[1:])
            S.Hz[self.x0 - 1] = self.t['recv_h_b'].numpy().reshape(S.Hz.shape[1:])
        if self.pending == 'e' and self.x1 <= self.G.nx:
            S.Ey[self.x1] = self.t['recv_e_a'].numpy().reshape(S.Ey.shape[1:])
            S.Ez[self.x1] = self.t['recv_e_b'].numpy().reshape(S.Ez.shape[1:])
        self.pending = None

    def _poison(self):
        for n in ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz'):
            F = getattr(self.S, n)
            F[:self.x0] = np.nan
            F[self.x1:] = np.nan

    def _phase(self, phase, S):
        G = self.G
        if phase == 0:
            self.K.update_magnetic(S)
            for n, pml in enumerate(G.pmls):
                self.K.pml_magnetic(S, n, pml)
        else:
            self.K.update_electric(S)
            for n, pml in enumerate(G.pmls):
                self.K.pml_electric(S, n, pml)
            for src in G.hertziandipoles:
                if self._owned(src.xcoord):
                    self.hertz(src, self.it, S, G)

    def half_step(self, phase, part):
        import copy
        S = self.S
        if part == 0:
            self._apply_ghosts()
            if phase == 0:
                for rx, out in zip(self.G.rxs, self.rx):
                    if self._owned(rx.xcoord):
                        for k in out:
                            out[k][self.it] = getattr(S, k)[rx.xcoord, rx.ycoord, rx.zcoord]
            # boundary plane first: computed on a scratch copy, published to the send tensors
            T = copy.deepcopy(S)
            self._phase(phase, T)
            if phase == 0:
                self.t['send_h_a'].copy_(self.torch.from_numpy(T.Hy[self.x1 - 1].reshape(-1).copy()))
                self.t['send_h_b'].copy_(self.torch.from_numpy(T.Hz[self.x1 - 1].reshape(-1).copy()))
            else:
                self.t['send_e_a'].copy_(self.torch.from_numpy(T.Ey[self.x0].reshape(-1).copy()))
                self.t['send_e_b'].copy_(self.torch.from_numpy(T.Ez[self.x0].reshape(-1).copy()))
        else:
            self._phase(phase, S)
            self._poison()
            self.pending = 'h' if phase == 0 else 'e'
            if phase == 1:
                self.it += 1


def _worker(rank, world, port, name, nit, q):
    import torch.distributed as dist
    from gprmax_b200.model_io import load_model
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ['OMP_NUM_THREADS'] = '1'
    dist.init_process_group('gloo', rank=rank, world_size=world)
    G, _ = load_model(golden_path(name, 'f32'))
    G.iterations = nit
    eng = OracleShard(G, rank, world)
    run_sharded(eng, HaloExchange(rank, world), nit, overlap=True)
    q.put((rank, eng.rx, [eng._owned(rx.xcoord) for rx in G.rxs]))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize('world', [2, 3])
def test_halo_protocol_bit_exact_over_gloo(world, oracle_built):
    import torch.multiprocessing as mp
    from gprmax_b200.model_io import load_model
    from oracle.solver import solve_cpu
    name, nit = 'pml_HORIPML_1', 60
    G, _ = load_model(golden_path(name, 'f32'))
    ref = solve_cpu(G, kernels='oracle', iterations=nit)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, nit, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned_by = {}
    for rank, rx, owned in results:
        for n, o in enumerate(owned):
            if o:
                assert n not in owned_by
                owned_by[n] = rx[n]
    assert sorted(owned_by) == list(range(len(G.rxs)))
    for n, out in owned_by.items():
        for k, v in out.items():
            r = ref['rx{}_{}'.format(n, k)]
            assert not np.isnan(v).any(), (n, k)
            assert np.array_equal(v[:nit], r[:nit]), (n, k)


# ---------------------------------------------------------------------------------------- sharded host build of the ID array

class _Mat(object):
    def __init__(self, numID, ID):
        self.numID, self.ID = numID, ID


def _avg(i, j, k, *args):
    """Stand-in for the reference's create_electric_average / create_magnetic_average (yee_cell_build_ext.pyx:31-107) with the
    property this test is about: the number a combination gets depends on the ORDER in which the combinations are resolved."""
    G, comp, ids = args[-1], args[-2], args[:-2]
    name = '+'.join(sorted(G.materials[n].ID for n in ids))
    for m in G.materials:
        if m.ID == name:
            G.ID[comp, i, j, k] = m.numID
            return
    G.materials.append(_Mat(len(G.materials), name))
    G.ID[comp, i, j, k] = len(G.materials) - 1


def _geometry(nx=37, ny=18, nz=22, seed=4):
    import types
    rng = np.random.default_rng(seed)
    G = types.SimpleNamespace(nx=nx, ny=ny, nz=nz)
    blocks = rng.integers(0, 5, size=(nx // 3 + 1, ny // 4 + 1, nz // 5 + 1)).astype(np.uint32)     # blocky geometry: runs of equal cells and interfaces
    G.solid = np.ascontiguousarray(np.repeat(np.repeat(np.repeat(blocks, 3, 0), 4, 1), 5, 2)[:nx, :ny, :nz])
    G.rigidE = (rng.random((12, nx, ny, nz)) < 0.02).astype(np.int8)
    G.rigidH = (rng.random((6, nx, ny, nz)) < 0.02).astype(np.int8)
    G.ID = rng.integers(0, 5, size=(6, nx + 1, ny + 1, nz + 1)).astype(np.uint32)     # what the geometry commands left on rigid edges
    G.materials = [_Mat(n, 'm{}'.format(n)) for n in range(5)]
    return G


def _build_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    G = _geometry()
    x0, n = partition_planes(G.nx, world)[rank]
    x1 = x0 + n
    s0, s1 = max(x0 - 1, 0), min(x1, G.nx)
    # this rank holds ONLY its slab (+ the cell plane to the left of its first node plane)
    from gprmax_b200.sharded import build_id_slab
    ID_local = np.ascontiguousarray(G.ID[:, x0:x1])
    ncombos = build_id_slab(G, np.ascontiguousarray(G.solid[s0:s1]), np.ascontiguousarray(G.rigidE[:, s0:s1]), np.ascontiguousarray(G.rigidH[:, s0:s1]),
                            ID_local, create_electric_average=_avg, create_magnetic_average=_avg)
    q.put((rank, x0, x1, ID_local, [(m.numID, m.ID) for m in G.materials], ncombos))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_id_build_over_gloo(world):
    """Every rank builds the ID planes of its slab from its slab of solid / rigid only; the distinct material combinations are
    exchanged (all_gather_object) and resolved in the global scan order on every rank: the slabs put together equal the
    single-process build and every rank ends with the same material list."""
    import torch.multiprocessing as mp
    from gprmax_b200 import yee_build
    G = _geometry()
    yee_build.build_components(G, _avg, _avg)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_build_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    mats = [(m.numID, m.ID) for m in G.materials]
    assert len(mats) > 5                                        # the geometry does produce averaged materials
    for rank, x0, x1, ID, mats_r, ncombos in results:
        assert np.array_equal(ID, G.ID[:, x0:x1]), rank
        assert mats_r == mats, rank
        assert ncombos > 0
