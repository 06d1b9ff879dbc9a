"""Pin the plain-C oracle (oracle/fdtd_oracle.c) against the reference's OWN compiled
Cython kernels (oracle/_ref, built from /root/reference by oracle/build_ref.py) at kernel
level: same random arrays in, element-wise comparison out, for every function on the path:
update_magnetic / update_electric (3-D and the three 2-D modes), dispersive A/B (1 and 3
poles), all 2 x 2 x 2 x 6 PML slab kernels, and the snapshot averaging.

The two builds differ only in FMA contraction (the reference is compiled with gcc's default
-ffp-contract=fast, the oracle with contraction off), hence a few-ulp tolerance.
"""
import numpy as np
import pytest

from conftest import have_ref_kernels
from gprmax_b200.model_io import SolverGrid, PMLSlab, SnapshotSpec
from oracle.solver import OracleKernels, ReferenceKernels, State, DIRECTIONS

pytestmark = pytest.mark.skipif(not have_ref_kernels('f32') or not have_ref_kernels('f64'),
                                reason='oracle/_ref not built (needs /root/reference once)')

TOL = {np.dtype(np.float32): 2e-5, np.dtype(np.float64): 1e-13}


def make_grid(real, dims, nmat=7, maxpoles=0, seed=0, pml=None, order=1, formulation='HORIPML'):
    rng = np.random.default_rng(seed)
    nx, ny, nz = dims
    real = np.dtype(real)
    cplx = np.complex64 if real == np.float32 else np.complex128
    G = SolverGrid(nx=nx, ny=ny, nz=nz, dx=1e-3, dy=1.5e-3, dz=2e-3, dt=1e-12, iterations=4)
    G.ID = rng.integers(0, nmat, size=(6, nx + 1, ny + 1, nz + 1), dtype=np.uint32)
    G.updatecoeffsE = rng.uniform(0.1, 1.0, size=(nmat, 5)).astype(real)
    G.updatecoeffsH = rng.uniform(0.1, 1.0, size=(nmat, 5)).astype(real)
    G.updatecoeffsE[0] = 0
    G.maxpoles = maxpoles
    if maxpoles:
        G.updatecoeffsdispersive = (rng.uniform(-1, 1, size=(nmat, 3 * maxpoles)) + 1j * rng.uniform(-1, 1, size=(nmat, 3 * maxpoles))).astype(cplx)
    G.pmlformulation = formulation
    G.cfs = [None] * order
    if pml is not None:
        direction, ext = pml
        xs, xf, ys, yf, zs, zf = ext
        t = (xf - xs, yf - ys, zf - zs)['xyz'.index(direction[0])]
        slab = PMLSlab(direction=direction, xs=xs, xf=xf, ys=ys, yf=yf, zs=zs, zf=zf, thickness=t,
                       d=float({'x': G.dx, 'y': G.dy, 'z': G.dz}[direction[0]]))
        for name in ('ERA', 'ERB', 'ERE', 'ERF', 'HRA', 'HRB', 'HRE', 'HRF'):
            setattr(slab, name, rng.uniform(0.5, 1.5, size=(order, t)).astype(real))
        G.pmls = [slab]
    return G


def fill_random(S, seed):
    rng = np.random.default_rng(seed)
    for n in ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz'):
        getattr(S, n)[...] = rng.uniform(-1, 1, size=getattr(S, n).shape)
    if S.maxpoles:
        for n in ('Tx', 'Ty', 'Tz'):
            getattr(S, n)[...] = rng.uniform(-1, 1, size=getattr(S, n).shape)


def two_states(G, seed=1):
    real = np.dtype(G.updatecoeffsE.dtype)
    Ko, Kr = OracleKernels(real), ReferenceKernels(real, nthreads=2)
    So, Sr = State(G, Ko), State(G, Kr)
    fill_random(So, seed)
    fill_random(Sr, seed)
    return Ko, Kr, So, Sr


def assert_fields_close(So, Sr, names, tol):
    for n in names:
        a, b = getattr(So, n), getattr(Sr, n)
        scale = max(1.0, float(np.abs(b).max()))
        assert np.abs(a - b).max() <= tol * scale, n


DIMS = [(9, 7, 8), (1, 9, 8), (9, 1, 8), (9, 8, 1)]


@pytest.mark.parametrize('real', [np.float32, np.float64])
@pytest.mark.parametrize('dims', DIMS)
def test_update_magnetic_electric(real, dims, oracle_built):
    G = make_grid(real, dims)
    Ko, Kr, So, Sr = two_states(G)
    Ko.update_magnetic(So)
    Kr.update_magnetic(Sr)
    assert_fields_close(So, Sr, ('Hx', 'Hy', 'Hz', 'Ex', 'Ey', 'Ez'), TOL[np.dtype(real)])
    Ko.update_electric(So)
    Kr.update_electric(Sr)
    assert_fields_close(So, Sr, ('Hx', 'Hy', 'Hz', 'Ex', 'Ey', 'Ez'), TOL[np.dtype(real)])


@pytest.mark.parametrize('real', [np.float32, np.float64])
@pytest.mark.parametrize('dims', DIMS)
@pytest.mark.parametrize('maxpoles', [1, 3])
def test_dispersive(real, dims, maxpoles, oracle_built):
    G = make_grid(real, dims, maxpoles=maxpoles)
    Ko, Kr, So, Sr = two_states(G)
    # the float `phi` of the reference (fields_updates_ext.pyx:143) costs float32 accuracy even in f64
    tol = 2e-5 if real == np.float32 else 1e-13
    for _ in range(2):
        Ko.update_electric_dispersive_A(So)
        Kr.update_electric_dispersive_A(Sr)
        assert_fields_close(So, Sr, ('Ex', 'Ey', 'Ez', 'Tx', 'Ty', 'Tz'), tol)
        Ko.update_electric_dispersive_B(So)
        Kr.update_electric_dispersive_B(Sr)
        assert_fields_close(So, Sr, ('Ex', 'Ey', 'Ez', 'Tx', 'Ty', 'Tz'), tol)


def slab_extent(direction, dims, t=3):
    nx, ny, nz = dims
    return {'xminus': (0, t, 0, ny, 0, nz), 'xplus': (nx - t, nx, 0, ny, 0, nz),
            'yminus': (0, nx, 0, t, 0, nz), 'yplus': (0, nx, ny - t, ny, 0, nz),
            'zminus': (0, nx, 0, ny, 0, t), 'zplus': (0, nx, 0, ny, nz - t, nz)}[direction]


@pytest.mark.parametrize('real', [np.float32, np.float64])
@pytest.mark.parametrize('formulation', ['HORIPML', 'MRIPML'])
@pytest.mark.parametrize('order', [1, 2])
@pytest.mark.parametrize('direction', DIRECTIONS)
def test_pml_slabs(real, formulation, order, direction, oracle_built):
    dims = (9, 8, 7)
    G = make_grid(real, dims, pml=(direction, slab_extent(direction, dims)), order=order, formulation=formulation)
    Ko, Kr, So, Sr = two_states(G)
    rng = np.random.default_rng(5)
    nxs, nys, nzs = So.EPhi1[0].shape[1:]
    for name in ('EPhi1', 'EPhi2', 'HPhi1', 'HPhi2'):
        v = rng.uniform(-1, 1, size=(order, nxs, nys, nzs))
        getattr(So, name)[0][...] = v
        getattr(Sr, name)[0][:, :nxs, :nys, :nzs] = v
    tol = TOL[np.dtype(real)] * 20
    for _ in range(2):
        Ko.pml_magnetic(So, 0, G.pmls[0])
        Kr.pml_magnetic(Sr, 0, G.pmls[0])
        Ko.pml_electric(So, 0, G.pmls[0])
        Kr.pml_electric(Sr, 0, G.pmls[0])
        assert_fields_close(So, Sr, ('Hx', 'Hy', 'Hz', 'Ex', 'Ey', 'Ez'), tol)
        for name in ('EPhi1', 'EPhi2', 'HPhi1', 'HPhi2'):
            a, b = getattr(So, name)[0], getattr(Sr, name)[0][:, :nxs, :nys, :nzs]
            assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max()), name


@pytest.mark.parametrize('real', [np.float32, np.float64])
def test_snapshot(real, oracle_built):
    G = make_grid(real, (12, 10, 9))
    Ko, Kr, So, Sr = two_states(G)
    for spec in (dict(xs=0, xf=12, ys=0, yf=10, zs=0, zf=9, dx=1, dy=1, dz=1, nx=12, ny=10, nz=9),
                 dict(xs=2, xf=10, ys=1, yf=10, zs=3, zf=9, dx=2, dy=3, dz=2, nx=4, ny=3, nz=3)):
        snap = SnapshotSpec(time=1, **spec)
        a = Ko.snapshot(So, snap)
        b = Kr.snapshot(Sr, snap)
        for x, y in zip(a, b):
            assert np.abs(x - y).max() <= TOL[np.dtype(real)]
