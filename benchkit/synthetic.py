"""Solver-ready grids for the benchmark workloads, built without the reference.

bench.py and the multi-GPU runs need the models BASELINE.json names -- the
`tests/benchmarking/bench_NxNxN.in` free-space cubes and the synthetic homogeneous
lossy-dielectric domain -- on a box where /root/reference does not exist, and (for the
4.3 Gcell case) shard by shard without ever holding the global ID array.  These models are a
homogeneous box + default PML + one Hertzian dipole + receivers, so the host build reduces to a
few closed-form tables.  The formulas restate the reference's host code:

    time step, iterations      input_cmds_singleuse.py:145-166, :195-203
    update coefficients        materials.py:61-143, :200-201
    CFS-PML profiles/tables    pml.py:56-146, :221-274 (default CFS: alpha 0, kappa 1, sigma quartic)
    waveforms                  waveforms.py:46-122
    Hertzian dipole / rx       input_cmds_multiuse.py:170-260 (dl = cell size along the polarisation)

NumPy expressions keep the operand dtypes of the reference (it mixes float32 profile arrays with
float64 scalars), so the tables are bit-identical to the reference's for the same NumPy; that is
pinned by tests/test_synthetic.py against the golden `bench_100` fixture written by the reference.
"""
from collections import OrderedDict
import decimal as d

import numpy as np
from scipy.constants import c, epsilon_0 as e0, mu_0 as m0

from gprmax_b200.model_io import PMLSlab, Receiver, SolverGrid, Source

z0 = np.sqrt(m0 / e0)
_SCALING = {'constant': 0, 'linear': 1, 'quadratic': 2, 'cubic': 3, 'quartic': 4, 'quintic': 5, 'sextic': 6, 'septic': 7, 'octic': 8}


def _round_half_down(value):
    return int(d.Decimal(value).quantize(d.Decimal('1'), rounding=d.ROUND_HALF_DOWN))


def _round_floor(value, places):
    return float(d.Decimal(value).quantize(d.Decimal('1.' + '0' * places), rounding=d.ROUND_FLOOR))


def time_step(dx, dy, dz, nx, ny, nz):
    """CFL time step, rounded down (input_cmds_singleuse.py:145-166)."""
    if nx == 1:
        dt = 1 / (c * np.sqrt((1 / dy) * (1 / dy) + (1 / dz) * (1 / dz)))
    elif ny == 1:
        dt = 1 / (c * np.sqrt((1 / dx) * (1 / dx) + (1 / dz) * (1 / dz)))
    elif nz == 1:
        dt = 1 / (c * np.sqrt((1 / dx) * (1 / dx) + (1 / dy) * (1 / dy)))
    else:
        dt = 1 / (c * np.sqrt((1 / dx) * (1 / dx) + (1 / dy) * (1 / dy) + (1 / dz) * (1 / dz)))
    return _round_floor(dt, d.getcontext().prec - 1)


def material_rows(er, se, mr, sm, dx, dy, dz, dt, real):
    """[CA,CBx,CBy,CBz,srce], [DA,DBx,DBy,DBz,srcm] of a non-dispersive material (materials.py:61-143)."""
    HA = (m0 * mr / dt) + 0.5 * sm
    HB = (m0 * mr / dt) - 0.5 * sm
    rowH = [HB / HA, (1 / dx) * 1 / HA, (1 / dy) * 1 / HA, (1 / dz) * 1 / HA, 1 / HA]
    if se == float('inf'):
        rowE = [0, 0, 0, 0, 0]
    else:
        EA = (e0 * er / dt) + 0.5 * se
        EB = (e0 * er / dt) - 0.5 * se
        rowE = [EB / EA, (1 / dx) * 1 / EA, (1 / dy) * 1 / EA, (1 / dz) * 1 / EA, 1 / EA]
    return np.array(rowE, dtype=real), np.array(rowH, dtype=real)


def waveform_value(wtype, amp, freq, time):
    """waveforms.py:46-122 for the analytic pulse shapes."""
    if wtype in ('gaussian', 'gaussiandot', 'gaussiandotnorm'):
        chi = 1 / freq
        zeta = 2 * np.pi**2 * freq**2
    elif wtype in ('gaussiandotdot', 'gaussiandotdotnorm', 'ricker'):
        chi = np.sqrt(2) / freq
        zeta = np.pi**2 * freq**2
    else:
        raise ValueError('waveform type {} not available in the synthetic builder'.format(wtype))
    delay = time - chi
    if wtype == 'gaussian':
        v = np.exp(-zeta * delay**2)
    elif wtype == 'gaussiandot':
        v = -2 * zeta * delay * np.exp(-zeta * delay**2)
    elif wtype == 'gaussiandotnorm':
        v = -2 * zeta * delay * np.exp(-zeta * delay**2) * np.sqrt(np.exp(1) / (2 * zeta))
    elif wtype == 'gaussiandotdot':
        v = 2 * zeta * (2 * zeta * delay**2 - 1) * np.exp(-zeta * delay**2)
    elif wtype == 'gaussiandotdotnorm':
        v = 2 * zeta * (2 * zeta * delay**2 - 1) * np.exp(-zeta * delay**2) * (1 / (2 * zeta))
    else:  # ricker
        v = - (2 * zeta * (2 * zeta * delay**2 - 1) * np.exp(-zeta * delay**2)) * (1 / (2 * zeta))
    return v * amp


def sample_waveform(wtype, amp, freq, dt, iterations, real, start=0.0, stop=None):
    """Whole- and half-step samples, zero outside [start, stop] (sources.py:47-68)."""
    whole = np.zeros(iterations, dtype=real)
    half = np.zeros(iterations, dtype=real)
    stop = np.inf if stop is None else stop
    for it in range(iterations):
        time = dt * it
        if time >= start and time <= stop:
            time -= start
            whole[it] = waveform_value(wtype, amp, freq, time)
            half[it] = waveform_value(wtype, amp, freq, time + 0.5 * dt)
    return whole, half


def _profile(thickness, profile, direction, vmin, vmax, real):
    """E and H sample points of one CFS parameter across a slab (pml.py:104-146)."""
    E = np.zeros(thickness + 1, dtype=real)
    H = np.zeros(thickness + 1, dtype=real)
    if profile == 'constant':
        E += vmax
        H += vmax
    else:
        order = _SCALING[profile]
        tmp = (np.linspace(0, (len(E) - 1) + 0.5, num=2 * len(E)) / (len(E) - 1)) ** order
        E = tmp[0:-1:2] * (vmax - vmin) + vmin
        H = tmp[1::2] * (vmax - vmin) + vmin
    if direction == 'reverse':
        E = E[::-1]
        H = np.roll(H[::-1], -1)
    return E[:-1], H[:-1]


DEFAULT_CFS = [dict(alpha=('constant', 'forward', 0, 0), kappa=('constant', 'forward', 1, 1), sigma=('quartic', 'forward', 0, None))]


def pml_tables(thickness, dspace, dt, er, mr, real, formulation='HORIPML', cfs=None):
    """ERA..HRF [order][thickness] of one slab (pml.py:221-274); sigma max from pml.py:71-83."""
    cfs = cfs or DEFAULT_CFS
    out = OrderedDict((k, np.zeros((len(cfs), thickness), dtype=real)) for k in ('ERA', 'ERB', 'ERE', 'ERF', 'HRA', 'HRB', 'HRE', 'HRF'))
    for x, term in enumerate(cfs):
        sprof, sdir, smin, smax = term['sigma']
        if not smax:
            smax = (0.8 * (_SCALING[sprof] + 1)) / (z0 * dspace * np.sqrt(er * mr))
        Ea, Ha = _profile(thickness, *term['alpha'], real=real)
        Ek, Hk = _profile(thickness, *term['kappa'], real=real)
        Es, Hs = _profile(thickness, sprof, sdir, smin, smax, real=real)
        for pre, a, k, s in (('E', Ea, Ek, Es), ('H', Ha, Hk, Hs)):
            if formulation == 'HORIPML':
                tmp = (2 * e0 * k) + dt * (a * k + s)
                out[pre + 'RA'][x, :] = (2 * e0 + dt * a) / tmp
                out[pre + 'RB'][x, :] = (2 * e0 * k) / tmp
                out[pre + 'RE'][x, :] = ((2 * e0 * k) - dt * (a * k + s)) / tmp
                out[pre + 'RF'][x, :] = (2 * s * dt) / (k * tmp)
            else:
                tmp = 2 * e0 + dt * a
                out[pre + 'RA'][x, :] = k + (dt * s) / tmp
                out[pre + 'RB'][x, :] = (2 * e0) / tmp
                out[pre + 'RE'][x, :] = ((2 * e0) - dt * a) / tmp
                out[pre + 'RF'][x, :] = (2 * s * dt) / tmp
    return out


def homogeneous_model(n, dcell=0.001, time_window=3e-9, iterations=None, real=np.float32,
                      er=1.0, se=0.0, mr=1.0, sm=0.0, pml_cells=10, formulation='HORIPML', cfs=None,
                      src=None, src_pol='x', waveform=('gaussiandotnorm', 1.0, 900e6), rxs=None,
                      x_range=None, build_id=True):
    """A homogeneous box with PML on all faces, one Hertzian dipole and receivers.

    With the defaults and n = (N, N, N) this is tests/benchmarking/bench_NxNxN.in: free space,
    1 mm cells, 3 ns, x-directed Hertzian dipole (gaussiandotnorm, 900 MHz) and one receiver in the
    cell at 0.05 m, default 10-cell first-order HORIPML.

    er/se/mr/sm != free space: the box is filled with that material (numID 2), as `#box` over the
    whole domain does.
    x_range = (x_start, nx_planes): build only that x-slab of the ID array (sharded runs).
    """
    nx, ny, nz = (n, n, n) if np.isscalar(n) else n
    real = np.dtype(real)
    dx = dy = dz = float(dcell)
    dt = time_step(dx, dy, dz, nx, ny, nz)
    window_given = iterations is None
    if iterations is None:
        iterations = int(np.ceil(time_window / dt)) + 1
    G = SolverGrid(nx=nx, ny=ny, nz=nz, dx=dx, dy=dy, dz=dz, dt=dt, iterations=int(iterations),
                   mode='3D', title='synthetic homogeneous box', pmlformulation=formulation)
    free = (er == 1.0 and se == 0.0 and mr == 1.0 and sm == 0.0)
    rows = [material_rows(1.0, float('inf'), 1.0, 0.0, dx, dy, dz, dt, real),   # 0 pec   (model_build_run.py:152-156)
            material_rows(1.0, 0.0, 1.0, 0.0, dx, dy, dz, dt, real)]            # 1 free_space (:157-159)
    if not free:
        rows.append(material_rows(er, se, mr, sm, dx, dy, dz, dt, real))      # 2 user material
    G.updatecoeffsE = np.stack([r[0] for r in rows])
    G.updatecoeffsH = np.stack([r[1] for r in rows])
    G.fill_id = 1 if free else 2
    x_start, nx_planes = (0, nx + 1) if x_range is None else x_range
    G.x_start, G.nx_planes = int(x_start), int(nx_planes)
    if build_id:
        G.ID = np.full((6, G.nx_planes, ny + 1, nz + 1), G.fill_id, dtype=np.uint32)
    # PML slabs in G.pmls order x0, y0, z0, xmax, ymax, zmax (grid.py:132-136, pml.py:383-417)
    G.cfs = list(cfs or DEFAULT_CFS)
    t = int(pml_cells)
    if t > 0:
        ext = {'xminus': (0, t, 0, ny, 0, nz), 'yminus': (0, nx, 0, t, 0, nz), 'zminus': (0, nx, 0, ny, 0, t),
               'xplus': (nx - t, nx, 0, ny, 0, nz), 'yplus': (0, nx, ny - t, ny, 0, nz), 'zplus': (0, nx, 0, ny, nz - t, nz)}
        for direction in ('xminus', 'yminus', 'zminus', 'xplus', 'yplus', 'zplus'):
            xs, xf, ys, yf, zs, zf = ext[direction]
            slab = PMLSlab(direction=direction, xs=xs, xf=xf, ys=ys, yf=yf, zs=zs, zf=zf, thickness=t,
                           d={'x': dx, 'y': dy, 'z': dz}[direction[0]])
            slab.nx, slab.ny, slab.nz = xf - xs, yf - ys, zf - zs
            for k, v in pml_tables(t, slab.d, dt, er, mr, real, formulation, G.cfs).items():
                setattr(slab, k, v)
            G.pmls.append(slab)
    # source (input_cmds_multiuse.py:170-230): coordinates round half down, dl = cell size
    if src is None:
        src = (0.05, 0.05, 0.05)
    sc = [_round_half_down(float(v) / dcell) for v in src]
    wt, amp, freq = waveform
    whole, half = sample_waveform(wt, amp, freq, dt, G.iterations, real)
    # the source stops at the time window (input_cmds_multiuse.py:213-216); with a float window iterations = ceil(tw / dt) + 1,
    # so the last iteration lies after it and its waveform sample is zero (sources.py:62-63)
    tw = time_window if window_given else (G.iterations - 1) * dt
    whole, half = sample_waveform(wt, amp, freq, dt, G.iterations, real, stop=float(tw))
    G.hertziandipoles = [Source(xcoord=sc[0], ycoord=sc[1], zcoord=sc[2], polarisation=src_pol, dl=dcell,
                                start=0.0, stop=float(tw), ID='HertzianDipole',
                                waveformvalues_wholestep=whole, waveformvalues_halfstep=half)]
    for r in (rxs if rxs is not None else [src]):
        rc = [_round_half_down(float(v) / dcell) for v in r]
        rx = Receiver(xcoord=rc[0], ycoord=rc[1], zcoord=rc[2], ID='')
        rx.outputs = OrderedDict((k, np.zeros(G.iterations, dtype=real)) for k in ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz'))
        G.rxs.append(rx)
    return G


def bench_model(N, real=np.float32, iterations=None):
    """tests/benchmarking/bench_NxNxN.in (BASELINE.json configs[1] for N = 300)."""
    side = N * 0.001
    return homogeneous_model(N, real=real, iterations=iterations, src=(0.05, 0.05, 0.05), rxs=[(0.05, 0.05, 0.05)]) if side else None
