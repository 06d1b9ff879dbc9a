"""`solve_gpu` -- drop-in for the reference's PyCUDA solver driver
(gprMax/model_build_run.py:477-716), running the time loop on libgprmax_b200.so.

Same call, same effects:

    tsolve, memsolve = solve_gpu(currentmodelrun, modelend, G)

reads the fully built grid `G` (FDTDGrid, grid.py:80-155) exactly as the reference does, and on
return has filled `rx.outputs[...]` for every receiver (receivers.py:71-88), `snap.electric /
snap.magnetic` for every snapshot (snapshots.py:214-228), and -- a feature the reference only has
on the CPU -- `tl.Vtotal / tl.Itotal` for transmission lines.  Unlike the reference's GPU path it
fills only the receiver outputs the `#rx` command asked for (as the CPU solver does), including
Ix/Iy/Iz.

There is no CPU fallback: without the compiled library or without a CUDA device this raises.
"""
import ctypes as C
import sys

import numpy as np

from . import _lib, vtk_writers
from .exceptions import GeneralError
from .model_io import DIRECTIONS, grid_maxpoles

_POL = {'x': 0, 'y': 1, 'z': 2}


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _active_range(src, G):
    """Inclusive iteration range in which the reference's CPU test
    `iteration * G.dt >= start and iteration * G.dt <= stop` (sources.py:92) holds."""
    it = np.arange(G.iterations)
    on = np.nonzero((it * G.dt >= src.start) & (it * G.dt <= src.stop))[0]
    if on.size == 0:
        return 1, 0
    return int(on[0]), int(on[-1])


class PackedModel(object):
    """The gpb_model_t for a grid plus the NumPy arrays that back its pointers."""

    def __init__(self, G, x_start=0, nx_planes=None, ID=None, with_id=True):
        self.keep = []
        real = np.dtype(G.updatecoeffsE.dtype)
        if real not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise GeneralError('unsupported floating-point type {}'.format(real))
        self.real = real
        cplx = np.dtype(np.complex64 if real == np.float32 else np.complex128)
        m = _lib.Model()
        m.abi_version = _lib.GPB_ABI_VERSION
        m.dtype = _lib.GPB_F32 if real == np.float32 else _lib.GPB_F64
        m.nx, m.ny, m.nz = int(G.nx), int(G.ny), int(G.nz)
        m.x_start = int(x_start)
        m.nx_planes = int(G.nx + 1 if nx_planes is None else nx_planes)
        m.dx, m.dy, m.dz, m.dt = float(G.dx), float(G.dy), float(G.dz), float(G.dt)
        m.iterations = int(G.iterations)
        if ID is None:
            ID = getattr(G, 'ID', None)
        m.id_comp_stride = 0
        if not with_id:
            m.ID = None
            m.uniform_id = 0
        elif ID is None:
            # homogeneous synthetic domain (synthetic.homogeneous_model(build_id=False)): no ID array at all
            m.ID = None
            m.uniform_id = int(G.fill_id)
        else:
            want = (6, m.nx_planes, m.ny + 1, m.nz + 1)
            whole = (6, m.nx + 1, m.ny + 1, m.nz + 1)
            if tuple(ID.shape) == whole and want != whole:
                # an x-slab of the global array: hand the library a pointer INTO G.ID with the global component stride -- the
                # slab goes to the device plane by plane without a second host copy (gpb_model_t::id_comp_stride)
                ID = self._hold(ID, np.uint32)
                m.ID = C.c_void_p(ID.ctypes.data + m.x_start * ID.strides[1])
                m.id_comp_stride = ID.strides[0] // 4
            elif tuple(ID.shape) != want:
                raise GeneralError('ID array has shape {}, expected {}'.format(tuple(ID.shape), want))
            else:
                m.ID = _ptr(self._hold(ID, np.uint32))
        cE = self._hold(G.updatecoeffsE, real)
        cH = self._hold(G.updatecoeffsH, real)
        m.nmaterials = int(cE.shape[0])
        m.updatecoeffsE, m.updatecoeffsH = _ptr(cE), _ptr(cH)
        m.maxpoles = grid_maxpoles(G)
        if m.maxpoles:
            m.updatecoeffsdispersive = _ptr(self._hold(G.updatecoeffsdispersive, cplx))
        # PML (pml.py:149-274)
        m.pml_formulation = {'HORIPML': _lib.GPB_HORIPML, 'MRIPML': _lib.GPB_MRIPML}[G.pmlformulation]
        m.pml_order = len(G.cfs) if G.pmls else 1
        m.npml = len(G.pmls)
        pmls = (_lib.Pml * max(1, m.npml))()
        for n, p in enumerate(G.pmls):
            q = pmls[n]
            q.direction = DIRECTIONS.index(p.direction)
            q.xs, q.xf, q.ys, q.yf, q.zs, q.zf = int(p.xs), int(p.xf), int(p.ys), int(p.yf), int(p.zs), int(p.zf)
            q.thickness = int(p.thickness)
            q.d = float(p.d)
            for t in ('ERA', 'ERB', 'ERE', 'ERF', 'HRA', 'HRB', 'HRE', 'HRF'):
                setattr(q, t, _ptr(self._hold(getattr(p, t), real)))
        self.keep.append(pmls)
        m.pmls = pmls
        # point sources, packed like sources.py:235-283; order inside each kind is list order
        srcs = []
        for s in G.voltagesources:
            hard = not s.resistance
            srcs.append((_lib.GPB_SRC_VOLTAGE, s, float(s.resistance or 0.0),
                         s.waveformvalues_halfstep if hard else s.waveformvalues_wholestep))
        for s in G.hertziandipoles:
            srcs.append((_lib.GPB_SRC_HERTZIAN, s, float(s.dl), s.waveformvalues_wholestep))
        for s in G.magneticdipoles:
            srcs.append((_lib.GPB_SRC_MAGNETIC, s, 0.0, s.waveformvalues_halfstep))
        m.nsources = len(srcs)
        arr = (_lib.Source * max(1, len(srcs)))()
        for n, (kind, s, param, wave) in enumerate(srcs):
            q = arr[n]
            q.kind = kind
            q.i, q.j, q.k = int(s.xcoord), int(s.ycoord), int(s.zcoord)
            q.polarisation = _POL[s.polarisation]
            q.it_first, q.it_last = _active_range(s, G)
            q.param = param
            q.waveform = _ptr(self._hold(wave, real, n=m.iterations))
        self.keep.append(arr)
        m.sources = arr
        # transmission lines (sources.py:286-452)
        tls = G.transmissionlines
        m.ntlines = len(tls)
        tarr = (_lib.TLine * max(1, len(tls)))()
        for n, t in enumerate(tls):
            q = tarr[n]
            q.i, q.j, q.k = int(t.xcoord), int(t.ycoord), int(t.zcoord)
            q.polarisation = _POL[t.polarisation]
            q.it_first, q.it_last = _active_range(t, G)
            q.nl, q.srcpos, q.antpos = int(t.nl), int(t.srcpos), int(t.antpos)
            q.resistance, q.dl = float(t.resistance), float(t.dl)
            q.abcv0, q.abcv1 = float(t.abcv0), float(t.abcv1)
            q.voltage0 = _ptr(self._hold(np.asarray(t.voltage)[:q.nl], real))
            q.current0 = _ptr(self._hold(np.asarray(t.current)[:q.nl], real))
            q.wave_whole = _ptr(self._hold(t.waveformvalues_wholestep, real, n=m.iterations))
            q.wave_half = _ptr(self._hold(t.waveformvalues_halfstep, real, n=m.iterations))
        self.keep.append(tarr)
        m.tlines = tarr
        # receivers (receivers.py:45-66)
        m.nrx = len(G.rxs)
        rxc = np.zeros((max(1, m.nrx), 3), dtype=np.int32)
        for n, rx in enumerate(G.rxs):
            rxc[n] = (rx.xcoord, rx.ycoord, rx.zcoord)
        m.rxcoords = _ptr(self._hold(rxc, np.int32))
        # snapshots (snapshots.py:28-84)
        m.nsnapshots = len(G.snapshots)
        sarr = (_lib.Snapshot * max(1, m.nsnapshots))()
        for n, s in enumerate(G.snapshots):
            q = sarr[n]
            for f in ('xs', 'ys', 'zs', 'xf', 'yf', 'zf', 'dx', 'dy', 'dz', 'nx', 'ny', 'nz', 'time'):
                setattr(q, f, int(getattr(s, f)))
        self.keep.append(sarr)
        m.snapshots = sarr
        self.model = m

    def _hold(self, a, dtype, n=None):
        b = np.ascontiguousarray(a, dtype=dtype)
        if n is not None and b.size != n:
            raise GeneralError('waveform has {} samples, expected {}'.format(b.size, n))
        self.keep.append(b)
        return b


class Solver(object):
    """Handle-owning wrapper around one gpb_handle."""

    def __init__(self, G, device_id=None, x_start=0, nx_planes=None, ID=None, devices=None):
        """device_id: CUDA runtime ordinal (default: G.gpu).  devices: a list of ordinals = ONE domain cut into x-slabs over
        those devices inside this process (gpb_create_sharded); the handle then stands for the whole domain."""
        self.L = _lib.lib()
        self.G = G
        gpu = getattr(G, 'gpu', None)
        if device_id is None and devices is None:
            ords = list(getattr(gpu, 'shard_ordinals', None) or [])
            if len(ords) > 1:
                devices = ords
            else:
                device_id = int(getattr(gpu, 'ordinal', getattr(gpu, 'deviceID', 0)) or 0)
        if devices is not None and len(devices) == 1:
            device_id, devices = int(devices[0]), None
        self.device_id = device_id
        self.devices = None if devices is None else [int(d) for d in devices]
        packed = PackedModel(G, x_start=x_start, nx_planes=nx_planes, ID=ID)
        self.real = packed.real
        self.iterations = int(G.iterations)
        self.nrx = len(G.rxs)
        self.x_start = int(packed.model.x_start)
        self.nx_planes = int(packed.model.nx_planes)
        self.h = C.c_void_p()
        if self.devices is not None:
            arr = (C.c_int * len(self.devices))(*self.devices)
            rc = self.L.gpb_create_sharded(C.byref(packed.model), arr, len(self.devices), C.byref(self.h))
        else:
            rc = self.L.gpb_create(C.byref(packed.model), int(device_id), C.byref(self.h))
        del packed  # the library has copied everything it needs
        if rc:
            self.h = None
            raise GeneralError(_lib.last_error())

    def _ck(self, rc):
        if rc:
            raise GeneralError(_lib.last_error())

    def close(self):
        if getattr(self, 'h', None):
            self.L.gpb_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def run(self, n=None):
        self._ck(self.L.gpb_run(self.h, int(self.iterations - self.iteration if n is None else n)))

    def half_step(self, phase, part=-1):
        self._ck(self.L.gpb_half_step(self.h, int(phase), int(part)))

    def profile(self, n):
        """Device milliseconds {prologue, H update, E update, sources} over n plain-launch iterations."""
        ms = (C.c_double * 4)()
        self._ck(self.L.gpb_profile(self.h, int(n), ms))
        return dict(begin=ms[0], update_h=ms[1], update_e=ms[2], sources=ms[3])

    @property
    def kernel_path(self):
        buf = C.create_string_buffer(256)
        self._ck(self.L.gpb_kernel_path(self.h, buf, 256))
        return buf.value.decode()

    def reset(self):
        self._ck(self.L.gpb_reset(self.h))

    def set_points(self, G):
        """Next trace on the resident geometry: new sources / receivers / transmission lines / snapshots of `G`, field state
        cleared, ID and coefficient arrays stay on the device (gpb_set_points)."""
        packed = PackedModel(G, x_start=self.x_start if self.devices is None else 0,
                             nx_planes=self.nx_planes if self.devices is None else None, with_id=False)
        self._ck(self.L.gpb_set_points(self.h, C.byref(packed.model)))
        self.G = G
        self.nrx = len(G.rxs)

    def synchronize(self):
        self._ck(self.L.gpb_synchronize(self.h))

    @property
    def iteration(self):
        v = C.c_int(0)
        self._ck(self.L.gpb_iteration(self.h, C.byref(v)))
        return v.value

    @property
    def elapsed(self):
        v = C.c_double(0)
        self._ck(self.L.gpb_elapsed_seconds(self.h, C.byref(v)))
        return v.value

    @property
    def mem_used(self):
        v = C.c_uint64(0)
        self._ck(self.L.gpb_mem_used(self.h, C.byref(v)))
        return int(v.value)

    @property
    def kernel_launches(self):
        v = C.c_uint64(0)
        self._ck(self.L.gpb_kernel_launches(self.h, C.byref(v)))
        return int(v.value)

    @property
    def stream(self):
        v = C.c_void_p()
        self._ck(self.L.gpb_stream(self.h, C.byref(v)))
        return v.value

    def link_info(self):
        """This slab's gpb_link_t (plain bytes: send it to the neighbouring ranks)."""
        info = _lib.Link()
        self._ck(self.L.gpb_link_info(self.h, C.byref(info)))
        return bytes(info)

    def link(self, left=None, right=None):
        """Link this slab to its x-neighbours (their link_info() bytes, or None at a domain face); link() unlinks."""
        l = _lib.Link.from_buffer_copy(left) if left is not None else None
        r = _lib.Link.from_buffer_copy(right) if right is not None else None
        self._ck(self.L.gpb_link(self.h, C.byref(l) if l is not None else None, C.byref(r) if r is not None else None))

    def halo(self, which):
        a, b, n = C.c_void_p(), C.c_void_p(), C.c_size_t(0)
        self._ck(self.L.gpb_halo(self.h, int(which), C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def receivers(self):
        """R[9][iterations][nrx]: Ex,Ey,Ez,Hx,Hy,Hz,Ix,Iy,Iz (fields_outputs.py:81-105 + grid.py:413-461)."""
        out = np.zeros((_lib.GPB_NRXOUT, self.iterations, self.nrx), dtype=self.real)
        self._ck(self.L.gpb_get_receivers(self.h, _ptr(out), out.nbytes))
        return out

    def snapshot(self, index):
        s = self.G.snapshots[index]
        outs = [np.zeros((s.nx, s.ny, s.nz), dtype=self.real) for _ in range(6)]
        ptrs = (C.c_void_p * 6)(*[o.ctypes.data for o in outs])
        self._ck(self.L.gpb_get_snapshot(self.h, int(index), ptrs, outs[0].nbytes))
        return outs

    def tline(self, index):
        v = np.zeros(self.iterations, dtype=self.real)
        i = np.zeros(self.iterations, dtype=self.real)
        self._ck(self.L.gpb_get_tline(self.h, int(index), _ptr(v), _ptr(i), v.nbytes))
        return v, i

    def get_field(self, comp):
        out = np.zeros((self.nx_planes, self.G.ny + 1, self.G.nz + 1), dtype=self.real)
        self._ck(self.L.gpb_get_field(self.h, int(comp), _ptr(out), out.nbytes))
        return out

    def set_field(self, comp, a):
        a = np.ascontiguousarray(a, dtype=self.real)
        self._ck(self.L.gpb_set_field(self.h, int(comp), _ptr(a), a.nbytes))


def store_results(G, solver):
    """Copy device results into the objects `write_hdf5_outputfile` (fields_outputs.py:111-161) and
    the snapshot writer (snapshots.py:132-167) read."""
    if G.rxs:
        rxs = solver.receivers()
        for n, rx in enumerate(G.rxs):
            for name in list(rx.outputs.keys()):
                rx.outputs[name] = np.ascontiguousarray(rxs[_lib.RX_ROWS.index(name), :, n])
    for n, snap in enumerate(G.snapshots):
        fields = solver.snapshot(n)
        if vtk_writers.installed():
            # the streaming writer re-orders the six component arrays block by block while it writes the file
            # (vtk_writers.write_vtk_imagedata): the two interleaved copies of the snapshot are never built
            snap.fields = fields
            snap.electric = snap.magnetic = None
        else:
            # snapshots.py:128-130: Paraview ordering, np.stack((ex, ey, ez)).reshape(-1, order='F') on all host cores
            snap.fields = None
            snap.electric = vtk_writers.paraview_vectors(*fields[:3])
            snap.magnetic = vtk_writers.paraview_vectors(*fields[3:])
    for n, tl in enumerate(G.transmissionlines):
        tl.Vtotal, tl.Itotal = solver.tline(n)


def solve_gpu(currentmodelrun, modelend, G):
    """Solving using FDTD method on GPU (drop-in for model_build_run.py:477-716).

    Args:
        currentmodelrun (int): Current model run number.
        modelend (int): Number of last model to run.
        G (class): Grid class instance - holds essential parameters describing the model.

    Returns:
        tsolve (float): Time taken to execute solving (time loop incl. the final receiver copy)
        memsolve (int): device memory used by the solver in bytes
    """
    # The same FDTDGrid object again = the reference's --geometry-fixed B-scan (model_build_run.py:109, 288-330: `G` is kept
    # between model runs and only sources / receivers are stepped): the solver of the previous trace is still resident on the
    # device and only takes the new points.  A new grid object gets a new solver.
    solver = _resident_solver(G)
    keep = solver is not None or bool(getattr(G, 'srcsteps', None) and any(G.srcsteps)) or bool(getattr(G, 'rxsteps', None) and any(G.rxsteps))
    if solver is None:
        solver = Solver(G)   # G.gpu decides: one device, or x-slabs over G.gpu.shard_ordinals (gprmax_b200/gpu.py)
    try:
        progress = bool(getattr(G, 'progressbars', False))
        total = int(G.iterations)
        if progress:
            from tqdm import tqdm
            bar = tqdm(total=total, desc='Running simulation, model ' + str(currentmodelrun) + '/' + str(modelend), file=sys.stdout)
            chunk = max(1, total // 50)
            done = 0
            while done < total:
                n = min(chunk, total - done)
                solver.run(n)
                done += n
                bar.update(n)
            bar.close()
        else:
            solver.run(total)
        import time
        t0 = time.perf_counter()
        store_results(G, solver)
        tcopy = time.perf_counter() - t0
        tsolve = solver.elapsed + tcopy
        memsolve = solver.mem_used
    except Exception:
        keep = False
        raise
    finally:
        if keep and currentmodelrun < modelend:
            _keep_resident(G, solver)
        else:
            _RESIDENT.clear()
            solver.close()
    return tsolve, memsolve


# (id(G) -> (weak reference to G, Solver)): at most one entry
_RESIDENT = {}


def _resident_solver(G):
    ent = _RESIDENT.get(id(G))
    if ent is None:
        for _, sv in _RESIDENT.values():   # another grid: the resident one is stale
            sv.close()
        _RESIDENT.clear()
        return None
    ref, sv = ent
    if ref() is not G or sv.h is None:
        _RESIDENT.clear()
        sv.close()
        return None
    sv.set_points(G)
    return sv


def _keep_resident(G, solver):
    import weakref
    try:
        ref = weakref.ref(G)
    except TypeError:       # an object that cannot be weakly referenced: no residency
        solver.close()
        return
    _RESIDENT.clear()
    _RESIDENT[id(G)] = (ref, solver)
