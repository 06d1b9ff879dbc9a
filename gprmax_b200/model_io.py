"""Solver-ready grid container and its on-disk form.

`solve_gpu` (gprmax_b200/solver.py) consumes the reference's `FDTDGrid` object `G`
exactly as the reference's own `solve_gpu` does (model_build_run.py:477-716); it
only ever *reads attributes*.  This module provides

  * `SolverGrid` & friends - plain attribute bags with the same attribute names
    as the reference classes (`FDTDGrid` grid.py:80-155, `PML` pml.py:149-199,
    `HertzianDipole`/`MagneticDipole`/`VoltageSource`/`TransmissionLine`
    sources.py:31-452, `Rx` receivers.py:26-42, `Snapshot` snapshots.py:28-84), so
    that a grid built by anything (the reference's model build, the synthetic
    builders in synthetic.py, a worker of the B-scan farm) can be handed to the core;
  * `save_model` / `load_model` - a compressed .npz of everything the time loop
    reads (the ID array, coefficient tables, PML tables, pre-sampled waveforms,
    receiver/snapshot descriptions).  Used for the parity fixtures in tests/golden/
    (written from the real reference by tests/golden/make_golden.py) and for
    shipping a built model to another process/GPU.
"""
from collections import OrderedDict
import json

import numpy as np

DIRECTIONS = ['xminus', 'yminus', 'zminus', 'xplus', 'yplus', 'zplus']  # pml.py:163
RX_FIELD_OUTPUTS = ['Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz']
RX_ALL_OUTPUTS = RX_FIELD_OUTPUTS + ['Ix', 'Iy', 'Iz']  # receivers.py:29


class _Bag(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __repr__(self):
        return '{}({})'.format(type(self).__name__, ', '.join(sorted(self.__dict__)))


class SolverGrid(_Bag):
    """Attribute-compatible stand-in for the reference FDTDGrid (grid.py:80-155)."""

    def __init__(self, **kw):
        self.title = ''
        self.messages = False
        self.progressbars = False
        self.gpu = None
        self.nthreads = 0
        self.mode = '3D'
        self.pmlformulation = 'HORIPML'
        self.cfs = [None]
        self.pmls = []
        self.maxpoles = 0
        self.voltagesources = []
        self.hertziandipoles = []
        self.magneticdipoles = []
        self.transmissionlines = []
        self.rxs = []
        self.snapshots = []
        self.snapsgpu2cpu = False
        self.srcsteps = [0, 0, 0]
        self.rxsteps = [0, 0, 0]
        self.IDlookup = {'Ex': 0, 'Ey': 1, 'Ez': 2, 'Hx': 3, 'Hy': 4, 'Hz': 5}
        _Bag.__init__(self, **kw)


class PMLSlab(_Bag):
    pass


class Source(_Bag):
    pass


class Receiver(_Bag):
    pass


class SnapshotSpec(_Bag):
    pass


def grid_maxpoles(G):
    """Number of dispersive poles the update must carry (reference: Material.maxpoles,
    a class attribute, materials.py:28; it equals updatecoeffsdispersive.shape[1] / 3,
    grid.py:190)."""
    mp = getattr(G, 'maxpoles', None)
    if mp is not None:
        return int(mp)
    ucd = getattr(G, 'updatecoeffsdispersive', None)
    if ucd is not None:
        return int(ucd.shape[1] // 3)
    return 0


def _src_dict(s, kind):
    d = dict(kind=kind, xcoord=int(s.xcoord), ycoord=int(s.ycoord), zcoord=int(s.zcoord),
             polarisation=str(s.polarisation), start=float(s.start), stop=float(s.stop),
             ID=str(getattr(s, 'ID', '') or ''))
    if kind == 'hertzian':
        d['dl'] = float(s.dl)
    if kind in ('voltage', 'tl'):
        d['resistance'] = float(s.resistance)
    if kind == 'tl':
        d.update(dl=float(s.dl), nl=int(s.nl), srcpos=int(s.srcpos), antpos=int(s.antpos),
                 abcv0=float(s.abcv0), abcv1=float(s.abcv1))
    return d


def save_model(G, path, golden=None, meta=None):
    """Write the solver-relevant state of `G` (+ optional golden outputs) to `path` (.npz)."""
    arrays = {}
    head = dict(
        nx=int(G.nx), ny=int(G.ny), nz=int(G.nz),
        dx=float(G.dx), dy=float(G.dy), dz=float(G.dz), dt=float(G.dt),
        iterations=int(G.iterations), mode=str(G.mode), title=str(getattr(G, 'title', '')),
        pmlformulation=str(G.pmlformulation), pmlorder=len(G.cfs) if G.pmls else 0,
        maxpoles=grid_maxpoles(G), dtype=np.dtype(G.updatecoeffsE.dtype).name,
        srcsteps=[int(v) for v in G.srcsteps], rxsteps=[int(v) for v in G.rxsteps],
        meta=meta or {})
    arrays['ID'] = np.ascontiguousarray(G.ID, dtype=np.uint32)
    arrays['updatecoeffsE'] = np.ascontiguousarray(G.updatecoeffsE)
    arrays['updatecoeffsH'] = np.ascontiguousarray(G.updatecoeffsH)
    if head['maxpoles'] > 0:
        arrays['updatecoeffsdispersive'] = np.ascontiguousarray(G.updatecoeffsdispersive)
    pmls = []
    for n, p in enumerate(G.pmls):
        pmls.append(dict(direction=str(p.direction), xs=int(p.xs), xf=int(p.xf), ys=int(p.ys), yf=int(p.yf),
                         zs=int(p.zs), zf=int(p.zf), thickness=int(p.thickness), d=float(p.d)))
        for t in ('ERA', 'ERB', 'ERE', 'ERF', 'HRA', 'HRB', 'HRE', 'HRF'):
            arrays['pml{}_{}'.format(n, t)] = np.ascontiguousarray(getattr(p, t))
    head['pmls'] = pmls
    srcs = []
    for kind, lst in (('voltage', G.voltagesources), ('hertzian', G.hertziandipoles),
                      ('magnetic', G.magneticdipoles), ('tl', G.transmissionlines)):
        for s in lst:
            n = len(srcs)
            srcs.append(_src_dict(s, kind))
            arrays['src{}_whole'.format(n)] = np.ascontiguousarray(s.waveformvalues_wholestep)
            arrays['src{}_half'.format(n)] = np.ascontiguousarray(s.waveformvalues_halfstep)
            if kind == 'tl':
                arrays['src{}_voltage'.format(n)] = np.ascontiguousarray(s.voltage[:s.nl])
                arrays['src{}_current'.format(n)] = np.ascontiguousarray(s.current[:s.nl])
    head['sources'] = srcs
    head['rxs'] = [dict(xcoord=int(r.xcoord), ycoord=int(r.ycoord), zcoord=int(r.zcoord),
                        ID=str(getattr(r, 'ID', '') or ''), outputs=list(r.outputs.keys())) for r in G.rxs]
    head['snapshots'] = [dict(xs=int(s.xs), xf=int(s.xf), ys=int(s.ys), yf=int(s.yf), zs=int(s.zs), zf=int(s.zf),
                              dx=int(s.dx), dy=int(s.dy), dz=int(s.dz), nx=int(s.nx), ny=int(s.ny), nz=int(s.nz),
                              time=int(s.time)) for s in G.snapshots]
    if golden:
        for k, v in golden.items():
            arrays['golden_' + k] = np.asarray(v)
    arrays['head'] = np.frombuffer(json.dumps(head).encode('utf-8'), dtype=np.uint8)
    np.savez_compressed(path, **arrays)


def load_model(path):
    """Read a model written by `save_model`.  Returns (G, golden) where `golden` maps the
    stored reference outputs (e.g. 'rx0_Ez') to arrays (empty when none were stored)."""
    z = np.load(path)
    head = json.loads(bytes(z['head']).decode('utf-8'))
    real = np.dtype(head['dtype'])
    G = SolverGrid(nx=head['nx'], ny=head['ny'], nz=head['nz'], dx=head['dx'], dy=head['dy'], dz=head['dz'],
                   dt=head['dt'], iterations=head['iterations'], mode=head['mode'], title=head['title'],
                   pmlformulation=head['pmlformulation'], maxpoles=head['maxpoles'],
                   srcsteps=head['srcsteps'], rxsteps=head['rxsteps'])
    G.meta = head.get('meta', {})
    G.cfs = [None] * max(1, head['pmlorder'])
    G.ID = z['ID']
    G.updatecoeffsE = z['updatecoeffsE']
    G.updatecoeffsH = z['updatecoeffsH']
    if head['maxpoles'] > 0:
        G.updatecoeffsdispersive = z['updatecoeffsdispersive']
    for n, p in enumerate(head['pmls']):
        slab = PMLSlab(**p)
        slab.ID = {'xminus': 'x0', 'yminus': 'y0', 'zminus': 'z0', 'xplus': 'xmax', 'yplus': 'ymax', 'zplus': 'zmax'}[p['direction']]
        slab.nx, slab.ny, slab.nz = p['xf'] - p['xs'], p['yf'] - p['ys'], p['zf'] - p['zs']
        for t in ('ERA', 'ERB', 'ERE', 'ERF', 'HRA', 'HRB', 'HRE', 'HRF'):
            setattr(slab, t, z['pml{}_{}'.format(n, t)])
        G.pmls.append(slab)
    for n, s in enumerate(head['sources']):
        kind = s.pop('kind')
        src = Source(**s)
        src.waveformvalues_wholestep = z['src{}_whole'.format(n)]
        src.waveformvalues_halfstep = z['src{}_half'.format(n)]
        if kind == 'tl':
            src.voltage = z['src{}_voltage'.format(n)].copy()
            src.current = z['src{}_current'.format(n)].copy()
            src.Vtotal = np.zeros(G.iterations, dtype=real)
            src.Itotal = np.zeros(G.iterations, dtype=real)
        {'voltage': G.voltagesources, 'hertzian': G.hertziandipoles,
         'magnetic': G.magneticdipoles, 'tl': G.transmissionlines}[kind].append(src)
    for r in head['rxs']:
        rx = Receiver(xcoord=r['xcoord'], ycoord=r['ycoord'], zcoord=r['zcoord'], ID=r['ID'])
        rx.outputs = OrderedDict((k, np.zeros(G.iterations, dtype=real)) for k in r['outputs'])
        G.rxs.append(rx)
    for s in head['snapshots']:
        G.snapshots.append(SnapshotSpec(**s))
    golden = {k[len('golden_'):]: z[k] for k in z.files if k.startswith('golden_')}
    return G, golden
