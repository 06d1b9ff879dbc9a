"""CPU oracle for the Yee time-stepping loop: restatement of the reference's `solve_cpu`.

TEST INFRASTRUCTURE ONLY -- the product (gprmax_b200/) never imports this module.
Allowed users: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference.

Parity status: PINNED (see the header of fdtd_oracle.c).

The loop below follows model_build_run.py:408-474 step for step:
    store rx -> snapshots -> H update -> H-PML (G.pmls order) -> [TL, magnetic dipoles]
    -> E update (plain | dispersive A) -> E-PML -> [voltage, TL, Hertzian] -> dispersive B
Field kernels come from one of two back-ends with identical signatures:
    kernels='oracle' : oracle/fdtd_oracle.c (this repo's plain-C restatement)
    kernels='ref'    : the reference's own Cython kernels compiled into oracle/_ref by
                       oracle/build_ref.py (used to pin the restatement and as the CPU baseline)
Point sources / receivers / transmission lines are Python scalar code in the reference
(sources.py, fields_outputs.py:40-64); they are restated here in Python with the same
NumPy scalar expressions so that rounding follows the same promotion rules.
"""
import ctypes
import glob
import importlib.util
import os
import subprocess
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, '_build')
DIRECTIONS = ['xminus', 'yminus', 'zminus', 'xplus', 'yplus', 'zplus']
c0 = 299792458.0  # scipy.constants.c, used by TransmissionLine (sources.py:26)


# --------------------------------------------------------------------------- build / load
def build_oracle(force=False):
    """gcc the plain-C restatement twice (float / double) into oracle/_build/."""
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(HERE, 'fdtd_oracle.c')
    out = {}
    for sfx, real in (('f32', 'float'), ('f64', 'double')):
        so = os.path.join(BUILD, 'liboracle_{}.so'.format(sfx))
        if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.run(['/usr/bin/gcc', '-O2', '-fopenmp', '-fPIC', '-shared', '-march=x86-64-v3',
                            '-DREAL=' + real, '-DSFX=' + sfx, '-o', so, src], check=True)
        out[sfx] = so
    return out


_libs = {}


def _lib(sfx):
    if sfx not in _libs:
        _libs[sfx] = ctypes.CDLL(build_oracle()[sfx])
    return _libs[sfx]


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class OracleKernels(object):
    """fdtd_oracle.c behind the call shapes the loop needs."""
    name = 'oracle'

    def __init__(self, real):
        self.real = np.dtype(real)
        self.sfx = 'f32' if self.real == np.float32 else 'f64'
        self.lib = _lib(self.sfx)

    def _f(self, name):
        return getattr(self.lib, '{}_{}'.format(name, self.sfx))

    def update_magnetic(self, S):
        self._f('oracle_update_magnetic')(S.nx, S.ny, S.nz, _p(S.cH), _p(S.ID), _p(S.Ex), _p(S.Ey), _p(S.Ez), _p(S.Hx), _p(S.Hy), _p(S.Hz))

    def update_electric(self, S):
        self._f('oracle_update_electric')(S.nx, S.ny, S.nz, _p(S.cE), _p(S.ID), _p(S.Ex), _p(S.Ey), _p(S.Ez), _p(S.Hx), _p(S.Hy), _p(S.Hz))

    def update_electric_dispersive_A(self, S):
        self._f('oracle_update_electric_dispersive_A')(S.nx, S.ny, S.nz, S.maxpoles, _p(S.cE), _p(S.cd), _p(S.ID), _p(S.Tx), _p(S.Ty), _p(S.Tz),
                                                       _p(S.Ex), _p(S.Ey), _p(S.Ez), _p(S.Hx), _p(S.Hy), _p(S.Hz))

    def update_electric_dispersive_B(self, S):
        self._f('oracle_update_electric_dispersive_B')(S.nx, S.ny, S.nz, S.maxpoles, _p(S.cd), _p(S.ID), _p(S.Tx), _p(S.Ty), _p(S.Tz),
                                                       _p(S.Ex), _p(S.Ey), _p(S.Ez))

    def _pml(self, fn, S, pml, coeffs, phi1, phi2, R):
        fn(S.formulation, S.order, DIRECTIONS.index(pml.direction), pml.xs, pml.xf, pml.ys, pml.yf, pml.zs, pml.zf,
           S.nx, S.ny, S.nz, _p(coeffs), _p(S.ID), _p(S.Ex), _p(S.Ey), _p(S.Ez), _p(S.Hx), _p(S.Hy), _p(S.Hz),
           _p(phi1), _p(phi2), _p(R[0]), _p(R[1]), _p(R[2]), _p(R[3]), ctypes.c_float(pml.d))

    def pml_electric(self, S, n, pml):
        self._pml(self._f('oracle_pml_electric'), S, pml, S.cE, S.EPhi1[n], S.EPhi2[n], S.ER[n])

    def pml_magnetic(self, S, n, pml):
        self._pml(self._f('oracle_pml_magnetic'), S, pml, S.cH, S.HPhi1[n], S.HPhi2[n], S.HR[n])

    def alloc_phi(self, S, pml):
        shp = (S.order, pml.xf - pml.xs, pml.yf - pml.ys, pml.zf - pml.zs)
        return [np.zeros(shp, dtype=self.real) for _ in range(4)]  # EPhi1, EPhi2, HPhi1, HPhi2

    def snapshot(self, S, snap):
        outs = [np.zeros((snap.nx, snap.ny, snap.nz), dtype=self.real) for _ in range(6)]
        self._f('oracle_snapshot')(S.nx, S.ny, S.nz, snap.xs, snap.ys, snap.zs, snap.dx, snap.dy, snap.dz, snap.nx, snap.ny, snap.nz,
                                   _p(S.Ex), _p(S.Ey), _p(S.Ez), _p(S.Hx), _p(S.Hy), _p(S.Hz), *[_p(o) for o in outs])
        return outs


def _load_ref_module(variant, relpath):
    pat = os.path.join(HERE, '_ref', variant, 'gprMax', relpath + '.*.so')
    hits = glob.glob(pat)
    if not hits:
        raise RuntimeError('reference kernels not built: {} (run oracle/build_ref.py where /root/reference exists)'.format(pat))
    name = 'gprMax.' + relpath.replace('/', '.')
    spec = importlib.util.spec_from_file_location(name, hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def have_ref(variant='f32'):
    """True when the reference's compiled hot-path kernels are present under oracle/_ref."""
    need = ['fields_updates_ext', 'pml_updates/pml_updates_electric_HORIPML_ext', 'pml_updates/pml_updates_magnetic_HORIPML_ext']
    return all(glob.glob(os.path.join(HERE, '_ref', variant, 'gprMax', n + '.*.so')) for n in need)


_refmods = {}


def ref_module(variant, relpath):
    key = (variant, relpath)
    if key not in _refmods:
        _refmods[key] = _load_ref_module(variant, relpath)
    return _refmods[key]


class ReferenceKernels(object):
    """The reference's own compiled Cython kernels (oracle/_ref), called the way
    model_build_run.py:436-470 and pml.py:276-296 call them."""
    name = 'ref'

    def __init__(self, real, nthreads=None):
        self.real = np.dtype(real)
        self.variant = 'f32' if self.real == np.float32 else 'f64'
        self.nthreads = int(nthreads or os.environ.get('OMP_NUM_THREADS') or os.cpu_count() or 1)
        self.fu = ref_module(self.variant, 'fields_updates_ext')
        self.snap = ref_module(self.variant, 'snapshots_ext')
        self.pml = {}

    def _pmlmod(self, S, which):
        key = (which, S.formulation)
        if key not in self.pml:
            self.pml[key] = ref_module(self.variant, 'pml_updates/pml_updates_{}_{}_ext'.format(which, ['HORIPML', 'MRIPML'][S.formulation]))
        return self.pml[key]

    def update_magnetic(self, S):
        self.fu.update_magnetic(S.nx, S.ny, S.nz, self.nthreads, S.cH, S.ID, S.Ex, S.Ey, S.Ez, S.Hx, S.Hy, S.Hz)

    def update_electric(self, S):
        self.fu.update_electric(S.nx, S.ny, S.nz, self.nthreads, S.cE, S.ID, S.Ex, S.Ey, S.Ez, S.Hx, S.Hy, S.Hz)

    def update_electric_dispersive_A(self, S):
        if S.maxpoles == 1:
            self.fu.update_electric_dispersive_1pole_A(S.nx, S.ny, S.nz, self.nthreads, S.cE, S.cdc, S.ID, S.Txc, S.Tyc, S.Tzc, S.Ex, S.Ey, S.Ez, S.Hx, S.Hy, S.Hz)
        else:
            self.fu.update_electric_dispersive_multipole_A(S.nx, S.ny, S.nz, self.nthreads, S.maxpoles, S.cE, S.cdc, S.ID, S.Txc, S.Tyc, S.Tzc, S.Ex, S.Ey, S.Ez, S.Hx, S.Hy, S.Hz)

    def update_electric_dispersive_B(self, S):
        if S.maxpoles == 1:
            self.fu.update_electric_dispersive_1pole_B(S.nx, S.ny, S.nz, self.nthreads, S.cdc, S.ID, S.Txc, S.Tyc, S.Tzc, S.Ex, S.Ey, S.Ez)
        else:
            self.fu.update_electric_dispersive_multipole_B(S.nx, S.ny, S.nz, self.nthreads, S.maxpoles, S.cdc, S.ID, S.Txc, S.Tyc, S.Tzc, S.Ex, S.Ey, S.Ez)

    def pml_electric(self, S, n, pml):
        f = getattr(self._pmlmod(S, 'electric'), 'order{}_{}'.format(S.order, pml.direction))
        R = S.ER[n]
        f(pml.xs, pml.xf, pml.ys, pml.yf, pml.zs, pml.zf, self.nthreads, S.cE, S.ID, S.Ex, S.Ey, S.Ez, S.Hx, S.Hy, S.Hz,
          S.EPhi1[n], S.EPhi2[n], R[0], R[1], R[2], R[3], pml.d)

    def pml_magnetic(self, S, n, pml):
        f = getattr(self._pmlmod(S, 'magnetic'), 'order{}_{}'.format(S.order, pml.direction))
        R = S.HR[n]
        f(pml.xs, pml.xf, pml.ys, pml.yf, pml.zs, pml.zf, self.nthreads, S.cH, S.ID, S.Ex, S.Ey, S.Ez, S.Hx, S.Hy, S.Hz,
          S.HPhi1[n], S.HPhi2[n], R[0], R[1], R[2], R[3], pml.d)

    def alloc_phi(self, S, pml):
        # shapes as the reference allocates them, pml.py:202-219
        o, nx, ny, nz = S.order, pml.xf - pml.xs, pml.yf - pml.ys, pml.zf - pml.zs
        a = pml.direction[0]
        if a == 'x':
            shp = [(o, nx + 1, ny, nz + 1), (o, nx + 1, ny + 1, nz), (o, nx, ny + 1, nz), (o, nx, ny, nz + 1)]
        elif a == 'y':
            shp = [(o, nx, ny + 1, nz + 1), (o, nx + 1, ny + 1, nz), (o, nx + 1, ny, nz), (o, nx, ny, nz + 1)]
        else:
            shp = [(o, nx, ny + 1, nz + 1), (o, nx + 1, ny, nz + 1), (o, nx + 1, ny, nz), (o, nx, ny + 1, nz)]
        return [np.zeros(s, dtype=self.real) for s in shp]

    def snapshot(self, S, snap):
        # snapshots.py:87-130
        sx = slice(snap.xs, snap.xf + snap.dx, snap.dx)
        sy = slice(snap.ys, snap.yf + snap.dy, snap.dy)
        sz = slice(snap.zs, snap.zf + snap.dz, snap.dz)
        sl = [np.ascontiguousarray(F[sx, sy, sz]) for F in (S.Ex, S.Ey, S.Ez, S.Hx, S.Hy, S.Hz)]
        outs = [np.zeros((snap.nx, snap.ny, snap.nz), dtype=self.real) for _ in range(6)]
        self.snap.calculate_snapshot_fields(snap.nx, snap.ny, snap.nz, *sl, *outs)
        return outs


# --------------------------------------------------------------------------- state
class State(object):
    """Host arrays of one run, laid out exactly as the reference holds them."""

    def __init__(self, G, kernels):
        real = np.dtype(G.updatecoeffsE.dtype)
        cplx = np.dtype(np.complex64 if real == np.float32 else np.complex128)
        self.real, self.cplx = real, cplx
        self.nx, self.ny, self.nz = int(G.nx), int(G.ny), int(G.nz)
        shp = (self.nx + 1, self.ny + 1, self.nz + 1)
        self.ID = np.ascontiguousarray(G.ID, dtype=np.uint32)
        self.cE = np.ascontiguousarray(G.updatecoeffsE, dtype=real)
        self.cH = np.ascontiguousarray(G.updatecoeffsH, dtype=real)
        for n in ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz'):
            setattr(self, n, np.zeros(shp, dtype=real))
        # Material.maxpoles (materials.py:28) = updatecoeffsdispersive.shape[1] / 3 (grid.py:190)
        mp = getattr(G, 'maxpoles', None)
        ucd = getattr(G, 'updatecoeffsdispersive', None)
        self.maxpoles = int(mp) if mp is not None else (int(ucd.shape[1] // 3) if ucd is not None else 0)
        if self.maxpoles:
            self.cdc = np.ascontiguousarray(G.updatecoeffsdispersive, dtype=cplx)
            self.cd = self.cdc.view(real)
            for n in ('Tx', 'Ty', 'Tz'):
                tc = np.zeros((self.maxpoles,) + shp, dtype=cplx)
                setattr(self, n + 'c', tc)
                setattr(self, n, tc.view(real))
        self.formulation = ['HORIPML', 'MRIPML'].index(G.pmlformulation)
        self.order = len(G.cfs) if G.pmls else 1
        self.EPhi1, self.EPhi2, self.HPhi1, self.HPhi2, self.ER, self.HR = [], [], [], [], [], []
        for pml in G.pmls:
            a, b, c, d = kernels.alloc_phi(self, pml)
            self.EPhi1.append(a)
            self.EPhi2.append(b)
            self.HPhi1.append(c)
            self.HPhi2.append(d)
            self.ER.append([np.ascontiguousarray(getattr(pml, t), dtype=real) for t in ('ERA', 'ERB', 'ERE', 'ERF')])
            self.HR.append([np.ascontiguousarray(getattr(pml, t), dtype=real) for t in ('HRA', 'HRB', 'HRE', 'HRF')])


# --------------------------------------------------------------------------- point sources
def _Ix(x, y, z, S, G):  # grid.py:413-427
    if y == 0 or z == 0:
        return 0
    return G.dy * (S.Hy[x, y, z - 1] - S.Hy[x, y, z]) + G.dz * (S.Hz[x, y, z] - S.Hz[x, y - 1, z])


def _Iy(x, y, z, S, G):  # grid.py:430-444
    if x == 0 or z == 0:
        return 0
    return G.dx * (S.Hx[x, y, z] - S.Hx[x, y, z - 1]) + G.dz * (S.Hz[x - 1, y, z] - S.Hz[x, y, z])


def _Iz(x, y, z, S, G):  # grid.py:447-461
    if x == 0 or y == 0:
        return 0
    return G.dx * (S.Hx[x, y - 1, z] - S.Hx[x, y, z]) + G.dy * (S.Hy[x, y, z] - S.Hy[x - 1, y, z])


_CUR = {'x': _Ix, 'y': _Iy, 'z': _Iz}
_POL = {'x': 0, 'y': 1, 'z': 2}


def _active(src, iteration, G):
    return iteration * G.dt >= src.start and iteration * G.dt <= src.stop


def _hertzian(src, it, S, G):  # sources.py:163-193
    if _active(src, it, G):
        i, j, k = src.xcoord, src.ycoord, src.zcoord
        p = _POL[src.polarisation]
        E = (S.Ex, S.Ey, S.Ez)[p]
        E[i, j, k] -= (S.cE[S.ID[p, i, j, k], 4] * src.waveformvalues_wholestep[it] * src.dl * (1 / (G.dx * G.dy * G.dz)))


def _magnetic_dipole(src, it, S, G):  # sources.py:202-232
    if _active(src, it, G):
        i, j, k = src.xcoord, src.ycoord, src.zcoord
        p = _POL[src.polarisation]
        H = (S.Hx, S.Hy, S.Hz)[p]
        H[i, j, k] -= (S.cH[S.ID[3 + p, i, j, k], 4] * src.waveformvalues_halfstep[it] * (1 / (G.dx * G.dy * G.dz)))


def _voltage(src, it, S, G):  # sources.py:81-120
    if _active(src, it, G):
        i, j, k = src.xcoord, src.ycoord, src.zcoord
        p = _POL[src.polarisation]
        E = (S.Ex, S.Ey, S.Ez)[p]
        d = (G.dx, G.dy, G.dz)
        if src.resistance != 0:
            d1, d2 = ((G.dy, G.dz), (G.dx, G.dz), (G.dx, G.dy))[p]
            E[i, j, k] -= (S.cE[S.ID[p, i, j, k], 4] * src.waveformvalues_wholestep[it] * (1 / (src.resistance * d1 * d2)))
        else:
            E[i, j, k] = - src.waveformvalues_halfstep[it] / d[p]


class _TL(object):
    """Transmission line state machine, sources.py:286-452 (after calculate_incident_V_I,
    i.e. with nl = antpos + 1 and whatever line state that pre-run left behind)."""

    def __init__(self, src, G, real):
        self.src = src
        self.nl = int(src.nl)
        self.voltage = np.array(src.voltage[:self.nl], dtype=real)
        self.current = np.array(src.current[:self.nl], dtype=real)
        self.abcv0 = src.abcv0
        self.abcv1 = src.abcv1
        self.dl = np.float64(src.dl)
        self.Vtotal = np.zeros(G.iterations, dtype=real)
        self.Itotal = np.zeros(G.iterations, dtype=real)

    def update_abc(self, G):  # :348-358
        h = (c0 * G.dt - self.dl) / (c0 * G.dt + self.dl)
        self.voltage[0] = h * (self.voltage[1] - self.abcv0) + self.abcv1
        self.abcv0 = self.voltage[0]
        self.abcv1 = self.voltage[1]

    def update_voltage(self, it, G):  # :360-377
        s = self.src
        self.voltage[1:self.nl] -= (s.resistance * (c0 * G.dt / self.dl) * (self.current[1:self.nl] - self.current[0:self.nl - 1]))
        self.voltage[s.srcpos] += ((c0 * G.dt / self.dl) * s.waveformvalues_wholestep[it])
        self.update_abc(G)

    def update_current(self, it, G):  # :379-393
        s = self.src
        self.current[0:self.nl - 1] -= ((1 / s.resistance) * (c0 * G.dt / self.dl) * (self.voltage[1:self.nl] - self.voltage[0:self.nl - 1]))
        self.current[s.srcpos - 1] += ((1 / s.resistance) * (c0 * G.dt / self.dl) * s.waveformvalues_halfstep[it])

    def update_electric(self, it, S, G):  # :395-424
        s = self.src
        if _active(s, it, G):
            self.update_voltage(it, G)
            p = _POL[s.polarisation]
            E = (S.Ex, S.Ey, S.Ez)[p]
            E[s.xcoord, s.ycoord, s.zcoord] = - self.voltage[s.antpos] / (G.dx, G.dy, G.dz)[p]

    def update_magnetic(self, it, S, G):  # :426-452
        s = self.src
        if _active(s, it, G):
            self.update_current(it, G)
            self.current[s.antpos] = _CUR[s.polarisation](s.xcoord, s.ycoord, s.zcoord, S, G)


# --------------------------------------------------------------------------- the loop
def solve_cpu(G, kernels='oracle', nthreads=None, iterations=None, keep_state=False):
    """Run the time loop on the CPU.  Returns a dict:
        rx{n}_{comp}      receiver traces (only the outputs each #rx asked for)
        tl{n}_Vtotal/Itotal, snap{n}_E{x,y,z}/H{x,y,z}
        tsolve            loop wall time (the reference's own definition, model_build_run.py:422,472)
        state             the State (final fields) when keep_state
    """
    real = np.dtype(G.updatecoeffsE.dtype)
    K = OracleKernels(real) if kernels == 'oracle' else ReferenceKernels(real, nthreads)
    if nthreads and kernels == 'oracle':
        os.environ['OMP_NUM_THREADS'] = str(nthreads)
    S = State(G, K)
    nit = int(G.iterations if iterations is None else iterations)
    fields = {'Ex': S.Ex, 'Ey': S.Ey, 'Ez': S.Ez, 'Hx': S.Hx, 'Hy': S.Hy, 'Hz': S.Hz}
    rxout = [dict((k, np.zeros(nit, dtype=real)) for k in rx.outputs) for rx in G.rxs]
    tls = [_TL(t, G, real) for t in G.transmissionlines]
    snaps = {}

    t0 = time.perf_counter()
    for it in range(nit):
        # fields_outputs.py:40-64
        for rx, out in zip(G.rxs, rxout):
            for name, arr in out.items():
                if name in fields:
                    arr[it] = fields[name][rx.xcoord, rx.ycoord, rx.zcoord]
                else:
                    arr[it] = _CUR[name[1]](rx.xcoord, rx.ycoord, rx.zcoord, S, G)
        for tl in tls:
            tl.Vtotal[it] = tl.voltage[tl.src.antpos]
            tl.Itotal[it] = tl.current[tl.src.antpos]
        for n, snap in enumerate(G.snapshots):
            if snap.time == it + 1:
                snaps[n] = K.snapshot(S, snap)
        K.update_magnetic(S)
        for n, pml in enumerate(G.pmls):
            K.pml_magnetic(S, n, pml)
        for tl in tls:
            tl.update_magnetic(it, S, G)
        for src in G.magneticdipoles:
            _magnetic_dipole(src, it, S, G)
        if S.maxpoles == 0:
            K.update_electric(S)
        else:
            K.update_electric_dispersive_A(S)
        for n, pml in enumerate(G.pmls):
            K.pml_electric(S, n, pml)
        for src in G.voltagesources:
            _voltage(src, it, S, G)
        for tl in tls:
            tl.update_electric(it, S, G)
        for src in G.hertziandipoles:
            _hertzian(src, it, S, G)
        if S.maxpoles:
            K.update_electric_dispersive_B(S)
    tsolve = time.perf_counter() - t0

    out = {'tsolve': tsolve, 'iterations': nit}
    for n, d in enumerate(rxout):
        for k, v in d.items():
            out['rx{}_{}'.format(n, k)] = v
    for n, tl in enumerate(tls):
        out['tl{}_Vtotal'.format(n)] = tl.Vtotal
        out['tl{}_Itotal'.format(n)] = tl.Itotal
    for n, arrs in snaps.items():
        for name, a in zip(('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz'), arrs):
            out['snap{}_{}'.format(n, name)] = a
    if keep_state:
        out['state'] = S
    return out
