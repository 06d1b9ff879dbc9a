"""Streaming VTK writers (gprmax_b200/vtk_writers.py + csrc/gpb_vtkio.cpp): the re-ordering routine
against NumPy, and the written files byte for byte against the reference's own `GeometryView.write_vtk`
(geometry_outputs.py:119-290) and `Snapshot.write_vtk_imagedata` (snapshots.py:132-167).  CPU only; the file comparisons need
the vendored reference (baseline/_ref)."""
import os
import sys
import types

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, ROOT)
from gprmax_b200 import vtk_writers  # noqa: E402

HAVE_REF = os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'gprMax'))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason='baseline/_ref not installed')


class Bar(object):
    def __init__(self):
        self.n = 0

    def update(self, n=1):
        self.n += n


@pytest.mark.parametrize('dtype', [np.int8, np.uint16, np.uint32, np.float32, np.float64])
@pytest.mark.parametrize('ncomp', [1, 3])
def test_transpose_matches_numpy(dtype, ncomp):
    rng = np.random.default_rng(5)
    shape = (37, 45, 71)
    arrays = [(rng.random(shape) * 100).astype(dtype) for _ in range(ncomp)]
    for start, step in (((0, 0, 0), (1, 1, 1)), ((3, 1, 2), (2, 3, 4)), ((5, 7, 40), (1, 2, 1))):
        count = tuple(len(range(start[a], shape[a], step[a])) for a in range(3))
        got = vtk_writers.transpose(arrays, start, count, step)
        sl = tuple(slice(start[a], None, step[a]) for a in range(3))
        want = np.stack([a[sl] for a in arrays]).reshape(-1, order='F')
        assert got.dtype == want.dtype and np.array_equal(got, want)


def test_transpose_large_multithreaded_and_views():
    rng = np.random.default_rng(6)
    big = rng.random((130, 96, 150)).astype(np.float32)
    ex, ey, ez = big[:, :, 0:50], big[:, :, 50:100], big[:, :, 100:150]     # views: same shape and strides, not contiguous
    got = vtk_writers.paraview_vectors(ex, ey, ez)
    assert np.array_equal(got, np.stack((ex, ey, ez)).reshape(-1, order='F'))
    # a z range of the output = the same rows of the whole output
    part = vtk_writers.transpose([ex, ey, ez], (0, 0, 17), (130, 96, 9), (1, 1, 1))
    assert np.array_equal(part, got.reshape(50, -1)[17:26].reshape(-1))


def test_transpose_rejects_bad_ranges():
    a = np.zeros((4, 5, 6), dtype=np.float32)
    with pytest.raises(ValueError):
        vtk_writers.transpose([a], (0, 0, 0), (4, 5, 7), (1, 1, 1))
    with pytest.raises(ValueError):
        vtk_writers.transpose([a, np.zeros((4, 5, 7), dtype=np.float32)], (0, 0, 0), (4, 5, 6), (1, 1, 1))


def _grid(nx, ny, nz, seed=1):
    """A stand-in for FDTDGrid with just what the writers read (grid.py:157-169; geometry_outputs.py:134-141, 292-310)."""
    rng = np.random.default_rng(seed)
    G = types.SimpleNamespace(nx=nx, ny=ny, nz=nz, dx=0.002, dy=0.0025, dz=0.001)
    G.solid = rng.integers(0, 7, size=(nx, ny, nz)).astype(np.uint32)
    G.ID = rng.integers(0, 9, size=(6, nx + 1, ny + 1, nz + 1)).astype(np.uint32)
    t = 3
    boxes = [(0, t, 0, ny, 0, nz), (0, nx, 0, t, 0, nz), (0, nx, 0, ny, 0, t), (nx - t, nx, 0, ny, 0, nz), (0, nx, ny - t, ny, 0, nz), (0, nx, 0, ny, nz - t, nz)]
    G.pmls = [types.SimpleNamespace(xs=b[0], xf=b[1], ys=b[2], yf=b[3], zs=b[4], zf=b[5]) for b in boxes]

    def pt(x, y, z, name):
        return types.SimpleNamespace(xcoord=x, ycoord=y, zcoord=z, ID=name)
    G.hertziandipoles = [pt(6, 6, 6, 'hd1'), pt(1, 1, 1, 'hd_in_pml'), pt(7, 6, 6, 'hd_off_sample')]
    G.magneticdipoles = [pt(6, 6, 6, 'md_same_cell')]
    G.voltagesources = [pt(nx - 2, 8, 10, 'vs')]
    G.transmissionlines = [pt(8, ny - 1, 4, 'tl')]
    G.rxs = [pt(6, 8, 10, 'rx1'), pt(nx - 1, ny - 1, nz - 1, 'rx_last_cell'), pt(nx, ny, nz, 'rx_on_far_node')]
    G.materials = [types.SimpleNamespace(ID='mat{}'.format(n), numID=n) for n in range(9)]
    return G


def _both(tmp_path, make_view, G, block_bytes):
    import baseline
    baseline.use_reference()
    from gprMax.geometry_outputs import GeometryView
    ref_write = GeometryView.__dict__.get('_b200_ref_write_vtk') or GeometryView.write_vtk
    if ref_write is vtk_writers.write_vtk:
        pytest.skip('the reference writer is already patched in this process')
    files, bars = [], []
    for name, writer in (('ref', ref_write), ('new', vtk_writers.write_vtk)):
        v = make_view(GeometryView)
        v.filename = str(tmp_path / (name + v.fileext))
        bar = Bar()
        old = vtk_writers.BLOCK_BYTES
        vtk_writers.BLOCK_BYTES = block_bytes
        try:
            writer(v, G, bar)
        finally:
            vtk_writers.BLOCK_BYTES = old
        with open(v.filename, 'rb') as f:
            files.append(f.read())
        bars.append((bar.n, v.datawritesize))
    return files, bars


@needs_ref
@pytest.mark.parametrize('block_bytes', [1 << 10, 64 << 20])
@pytest.mark.parametrize('box', [((0, 0, 0), (24, 20, 28), (1, 1, 1)), ((0, 0, 0), (24, 20, 28), (2, 2, 2)), ((3, 2, 4), (21, 17, 25), (3, 5, 7)),
                                 ((6, 6, 6), (7, 7, 7), (1, 1, 1))])
def test_geometry_view_vti_identical(tmp_path, box, block_bytes):
    G = _grid(24, 20, 28)
    (xs, ys, zs), (xf, yf, zf), (dx, dy, dz) = box
    files, bars = _both(tmp_path, lambda GV: GV(xs, ys, zs, xf, yf, zf, dx, dy, dz, 'view', '.vti'), G, block_bytes)
    assert files[0] == files[1]
    assert bars[0][0] == bars[1][0]       # the progress bar receives the same total


@needs_ref
@pytest.mark.parametrize('block_bytes', [1 << 10, 64 << 20])
@pytest.mark.parametrize('box', [((0, 0, 0), (12, 10, 14)), ((3, 2, 4), (9, 10, 5)), ((0, 0, 0), (1, 10, 14))])
def test_geometry_view_vtp_identical(tmp_path, box, block_bytes):
    G = _grid(12, 10, 14)
    (xs, ys, zs), (xf, yf, zf) = box
    files, bars = _both(tmp_path, lambda GV: GV(xs, ys, zs, xf, yf, zf, 1, 1, 1, 'view', '.vtp'), G, block_bytes)
    assert files[0] == files[1]
    assert bars[0][0] == bars[1][0]


@needs_ref
def test_geometry_view_rejects_ragged_extent(tmp_path):
    import baseline
    baseline.use_reference()
    from gprMax.geometry_outputs import GeometryView
    G = _grid(24, 20, 28)
    v = GeometryView(0, 0, 0, 23, 20, 28, 2, 2, 2, 'view', '.vti')     # 23 / 2: the reference's loop overruns its arrays
    v.filename = str(tmp_path / 'ragged.vti')
    with pytest.raises(ValueError):
        vtk_writers.write_vtk(v, G, Bar())


@needs_ref
@pytest.mark.parametrize('block_bytes', [1 << 10, 64 << 20])
@pytest.mark.parametrize('box', [((0, 0, 0), (24, 20, 28), (1, 1, 1)), ((2, 4, 6), (20, 16, 26), (2, 3, 4))])
def test_snapshot_vti_identical(tmp_path, box, block_bytes):
    import baseline
    baseline.use_reference()
    from gprMax.constants import floattype
    from gprMax.snapshots import Snapshot
    ref_write = Snapshot.write_vtk_imagedata
    if ref_write is vtk_writers.write_vtk_imagedata:
        pytest.skip('the reference writer is already patched in this process')
    G = _grid(24, 20, 28)
    (xs, ys, zs), (xf, yf, zf), (dx, dy, dz) = box
    rng = np.random.default_rng(3)
    out = []
    for mode in ('ref', 'flat', 'components'):
        s = Snapshot(xs, ys, zs, xf, yf, zf, dx, dy, dz, 10, 'snap')
        s.filename = str(tmp_path / (mode + '.vti'))
        fields = [rng.standard_normal((s.nx, s.ny, s.nz)).astype(floattype) for _ in range(6)] if mode == 'ref' else fields  # noqa: F821
        if mode == 'components':
            s.fields, s.electric, s.magnetic = fields, None, None
        else:
            s.electric = np.stack(fields[:3]).reshape(-1, order='F')
            s.magnetic = np.stack(fields[3:]).reshape(-1, order='F')
        bar = Bar()
        old = vtk_writers.BLOCK_BYTES
        vtk_writers.BLOCK_BYTES = block_bytes
        try:
            (ref_write if mode == 'ref' else vtk_writers.write_vtk_imagedata)(s, bar, G)
        finally:
            vtk_writers.BLOCK_BYTES = old
        with open(s.filename, 'rb') as f:
            out.append((f.read(), bar.n, s.vtkdatawritesize))
    assert out[0][0] == out[1][0] == out[2][0]
    assert out[0][1] == out[1][1] == out[2][1] == out[0][2]


MODEL = '''#title: writers end to end
#domain: 0.100 0.080 0.060
#dx_dy_dz: 0.002 0.002 0.002
#time_window: 30
#material: 6 0.01 1 0 half_space
#waveform: ricker 1 1.5e9 my_ricker
#hertzian_dipole: z 0.050 0.044 0.030 my_ricker
#rx: 0.060 0.044 0.030
#box: 0 0 0 0.100 0.040 0.060 half_space
#cylinder: 0.050 0.020 0 0.050 0.020 0.060 0.006 pec
#geometry_view: 0 0 0 0.100 0.080 0.060 0.002 0.002 0.002 view_n n
#geometry_view: 0.010 0.010 0.004 0.090 0.070 0.052 0.004 0.002 0.008 view_sub n
#geometry_view: 0.010 0.010 0.004 0.090 0.070 0.052 0.002 0.002 0.002 view_f f
#snapshot: 0 0 0 0.100 0.080 0.060 0.002 0.002 0.002 20 snap_full
#snapshot: 0.010 0.010 0.004 0.090 0.070 0.052 0.004 0.002 0.008 25 snap_sub
'''


@needs_ref
def test_command_line_writes_identical_files(tmp_path):
    """`python -m gprmax_b200 model.in` (CPU solve: no device here) with the streaming writers installed by the drop-in against
    the same command with GPRMAX_B200_REF_WRITERS=1: every .vti / .vtp byte-identical."""
    import subprocess
    outs = {}
    for mode in ('new', 'ref'):
        d = tmp_path / mode
        d.mkdir()
        (d / 'model.in').write_text(MODEL)
        env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS='2')
        if mode == 'ref':
            env['GPRMAX_B200_REF_WRITERS'] = '1'
        r = subprocess.run([sys.executable, '-m', 'gprmax_b200', str(d / 'model.in')], env=env, cwd=str(d), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        files = {}
        for base, _, names in os.walk(str(d)):
            for n in names:
                if n.endswith(('.vti', '.vtp')):
                    with open(os.path.join(base, n), 'rb') as f:
                        files[os.path.relpath(os.path.join(base, n), str(d))] = f.read()
        outs[mode] = files
    assert sorted(outs['new']) == sorted(outs['ref']) and len(outs['new']) == 5, sorted(outs['new'])
    for name in outs['ref']:
        assert outs['new'][name] == outs['ref'][name], name


@needs_ref
def test_wide_block_headers_beyond_4_gib(tmp_path, monkeypatch):
    """An appended array of 4 GiB or more cannot be described by the reference's UInt32 byte count (its writer ends in a
    struct.error); the streaming writers switch to `header_type="UInt64"` there.  Forced on a small view: same payload, 8-byte
    counts, offsets that account for them."""
    import baseline
    baseline.use_reference()
    from gprMax.geometry_outputs import GeometryView
    from gprMax.snapshots import Snapshot
    from gprMax.constants import floattype
    G = _grid(24, 20, 28)
    narrow, wide = {}, {}
    for store, limit in ((narrow, 1 << 32), (wide, 0)):
        monkeypatch.setattr(vtk_writers, 'WIDE_HEADER_FROM', limit)
        v = GeometryView(0, 0, 0, 24, 20, 28, 2, 2, 2, 'view', '.vti')
        v.filename = str(tmp_path / 'v.vti')
        vtk_writers.write_vtk(v, G, Bar())
        store['view'] = open(v.filename, 'rb').read()
        s = Snapshot(0, 0, 0, 24, 20, 28, 1, 1, 1, 5, 'snap')
        s.filename = str(tmp_path / 's.vti')
        rng = np.random.default_rng(8)
        s.fields = [rng.standard_normal((s.nx, s.ny, s.nz)).astype(floattype) for _ in range(6)]
        s.electric = s.magnetic = None
        vtk_writers.write_vtk_imagedata(s, Bar(), G)
        store['snap'] = open(s.filename, 'rb').read()
    mark = b'<AppendedData encoding="raw">\n_'
    for kind, nblocks in (('view', 3), ('snap', 2)):
        a, b = narrow[kind], wide[kind]
        assert b'header_type' not in a and b'header_type="UInt64"' in b
        pa, pb = a.index(mark) + len(mark), b.index(mark) + len(mark)
        import re
        offs_a = [int(x) for x in re.findall(rb'offset="(\d+)"', a[:pa])]
        offs_b = [int(x) for x in re.findall(rb'offset="(\d+)"', b[:pb])]
        assert len(offs_a) == len(offs_b) == nblocks
        for n in range(nblocks):
            size_a = int(np.frombuffer(a[pa + offs_a[n]:pa + offs_a[n] + 4], dtype='<u4')[0])
            size_b = int(np.frombuffer(b[pb + offs_b[n]:pb + offs_b[n] + 8], dtype='<u8')[0])
            assert size_a == size_b
            assert a[pa + offs_a[n] + 4:pa + offs_a[n] + 4 + size_a] == b[pb + offs_b[n] + 8:pb + offs_b[n] + 8 + size_b]
            assert offs_b[n] == offs_a[n] + 4 * n
