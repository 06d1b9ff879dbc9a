"""Streaming VTK writers -- drop-ins for `Snapshot.write_vtk_imagedata` (gprMax/snapshots.py:132-167) and
`GeometryView.write_vtk` (gprMax/geometry_outputs.py:119-290), SURVEY.md 8(f) rank 4.

The files are byte-identical to the reference's (tests/test_vtk_writers.py compares them with the reference's own writers);
what changes is how they are produced:

* ParaView wants x fastest / z slowest, the model arrays are z fastest.  The reference re-orders the whole volume in one go
  -- a serial Cython loop that walks the slowest memory axis in its inner loop for geometry views
  (geometry_outputs_ext.pyx:81-110), `np.stack(...).reshape(-1, order='F')` for snapshots (snapshots.py:128-130) -- and holds
  every output array in memory before the first byte is written.  Here the re-ordering is the blocked, multi-threaded
  `gpb_vtk_transpose` of libgprmax_b200.so, asked for a few z planes at a time, so a file is streamed through a bounded
  buffer (BLOCK_BYTES) whatever the size of the model;
* the `Sources_PML` / `Receivers` arrays of a `.vti` geometry view are painted per block from the PML boxes and the point
  lists; the two int8 copies of the whole grid the reference allocates (geometry_outputs.py:134-141) do not exist;
* a `.vtp` (per-edge) geometry view is generated section by section in x-plane chunks from closed forms of the reference's
  running counters (geometry_outputs_ext.pyx:48-78) instead of being held in memory at 60 bytes per cell, and the cell
  offsets are one `arange` instead of a Python loop with one `struct.pack` and one progress-bar update per line
  (geometry_outputs.py:266-268).

* an appended array of 4 GiB or more (a `.vti` view of more than 1.07 G cells, a snapshot of more than 358 M cells) gets 64-bit
  block headers (`header_type="UInt64"`); the reference's `pack('I', ...)` fails there.

`install()` patches the two methods on the reference's classes (done by `gprmax_b200.dropin.install()`).
"""
import ctypes as C
from struct import pack

import numpy as np

from . import _lib
from .exceptions import GeneralError

BLOCK_BYTES = 64 << 20   # size of the streaming buffer
WIDE_HEADER_FROM = 1 << 32   # an appended array of this many bytes or more needs 64-bit block headers (see _block_header)


def _block_header(sizes):
    """Every appended array is preceded by its byte count.  The reference writes it as UInt32 (`pack('I', ...)`: an array of 4 GiB
    or more -- a `.vti` view of more than 1.07 G cells, a snapshot of more than 358 M cells -- ends in a struct.error).  Below that
    limit the files are the reference's byte for byte; above it the count is written as UInt64 and the file says so
    (`header_type="UInt64"`, VTK XML format 1.0).  Returns (struct format, header bytes, attribute text for <VTKFile>)."""
    if max(sizes) >= WIDE_HEADER_FROM:
        return 'Q', 8, ' header_type="UInt64"'
    return 'I', 4, ''


def transpose(arrays, start, count, step):
    """ParaView order out of z-fastest arrays: flat array `out[((k*ny + j)*nx + i)*ncomp + c] = arrays[c][start + (i,j,k)*step]`
    with (nx, ny, nz) = count.  `arrays`: same dtype, same shape, same strides (views are fine)."""
    a0 = arrays[0]
    item = a0.dtype.itemsize
    for a in arrays:
        if a.dtype != a0.dtype or a.shape != a0.shape or a.strides != a0.strides or a.ndim != 3:
            raise GeneralError('transpose: the component arrays must be 3-D and share dtype, shape and strides')
    if any(s % item or s < 0 for s in a0.strides):
        raise GeneralError('transpose: strides must be non-negative multiples of the element size')
    for ax in range(3):
        if count[ax] and start[ax] + (count[ax] - 1) * step[ax] >= a0.shape[ax]:
            raise GeneralError('transpose: range outside the array')
    out = np.empty(int(count[0]) * int(count[1]) * int(count[2]) * len(arrays), dtype=a0.dtype)
    src = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
    rc = _lib.lib().gpb_vtk_transpose(src, len(arrays), item, (C.c_int64 * 3)(*[s // item for s in a0.strides]),
                                      (C.c_int32 * 3)(*[int(v) for v in start]), (C.c_int32 * 3)(*[int(v) for v in count]),
                                      (C.c_int32 * 3)(*[int(v) for v in step]), out.ctypes.data)
    if rc:
        raise GeneralError('gpb_vtk_transpose failed ({})'.format(rc))
    return out


def paraview_vectors(fx, fy, fz):
    """`np.stack((fx, fy, fz)).reshape(-1, order='F')` (snapshots.py:128-130, 223-228) on all host cores."""
    return transpose([fx, fy, fz], (0, 0, 0), fx.shape, (1, 1, 1))


def _kblocks(nz, bytes_per_plane):
    kb = max(1, BLOCK_BYTES // max(1, bytes_per_plane))
    for k0 in range(0, nz, kb):
        yield k0, min(kb, nz - k0)


def _write_chunks(f, flat, pbar):
    """`flat.tofile(f)` in pieces, so that the progress bar moves and no second copy is made."""
    step = max(1, BLOCK_BYTES // flat.dtype.itemsize)
    for a in range(0, flat.size, step):
        part = flat[a:a + step]
        f.write(memoryview(part))
        pbar.update(n=part.nbytes)


# ------------------------------------------------------------------------------------------------------------ snapshots

def write_vtk_imagedata(self, pbar, G):
    """Drop-in for `Snapshot.write_vtk_imagedata` (snapshots.py:132-167): same file, written in blocks.

    Data source: `self.electric` / `self.magnetic` in ParaView order as `solve_gpu` / `Snapshot.store` leave them, or --
    when `self.fields` holds the six cell-centred component arrays `[nx][ny][nz]` (Ex..Hz) -- those, re-ordered block by
    block while the file is written (that is how `solve_gpu` leaves a snapshot once this writer is installed: the two
    interleaved copies are never built)."""
    from gprMax.snapshots import Snapshot
    from gprMax.utilities import round_value
    itemsize = np.dtype(_floattype()).itemsize
    hfmt, hbytes, hattr = _block_header([self.datasizefield])
    hfield_offset = 3 * itemsize * self.ncells + hbytes
    ext = (self.xs, round_value(self.xf / self.dx), self.ys, round_value(self.yf / self.dy), self.zs, round_value(self.zf / self.dz))
    with open(self.filename, 'wb') as f:
        f.write('<?xml version="1.0"?>\n'.encode('utf-8'))
        f.write('<VTKFile type="ImageData" version="1.0" byte_order="{}"{}>\n'.format(Snapshot.byteorder, hattr).encode('utf-8'))
        f.write('<ImageData WholeExtent="{} {} {} {} {} {}" Origin="0 0 0" Spacing="{:.3} {:.3} {:.3}">\n'.format(*ext, self.dx * G.dx, self.dy * G.dy, self.dz * G.dz).encode('utf-8'))
        f.write('<Piece Extent="{} {} {} {} {} {}">\n'.format(*ext).encode('utf-8'))
        f.write('<CellData Vectors="E-field H-field">\n'.encode('utf-8'))
        f.write('<DataArray type="{}" Name="E-field" NumberOfComponents="3" format="appended" offset="0" />\n'.format(Snapshot.floatname).encode('utf-8'))
        f.write('<DataArray type="{}" Name="H-field" NumberOfComponents="3" format="appended" offset="{}" />\n'.format(Snapshot.floatname, hfield_offset).encode('utf-8'))
        f.write('</CellData>\n</Piece>\n</ImageData>\n<AppendedData encoding="raw">\n_'.encode('utf-8'))
        fields = getattr(self, 'fields', None)
        for n, name in enumerate(('electric', 'magnetic')):
            f.write(pack(hfmt, self.datasizefield))
            pbar.update(n=4)
            if fields is not None:
                comps = fields[3 * n:3 * n + 3]
                for k0, kn in _kblocks(self.nz, 3 * itemsize * self.nx * self.ny):
                    block = transpose(comps, (0, 0, k0), (self.nx, self.ny, kn), (1, 1, 1))
                    f.write(memoryview(block))
                    pbar.update(n=block.nbytes)
            else:
                _write_chunks(f, getattr(self, name), pbar)
        f.write('\n</AppendedData>\n</VTKFile>'.encode('utf-8'))


def _floattype():
    from gprMax.constants import floattype
    return floattype


# ------------------------------------------------------------------------------------------------------- geometry views

def _samples(s, f, d):
    return len(range(s, f, d))


def _paint_boxes(block, boxes, view, k0):
    """block[kk][j][i] = 1 where the sampled cell lies in one of `boxes` = (xs, xf, ys, yf, zs, zf) (geometry_outputs.py:136-137)."""
    kn, nys, nxs = block.shape

    def rng(lo, hi, s, d, n):
        a = max(0, -((s - lo) // d))        # first sample index a with s + a*d >= lo
        b = min(n, -((s - hi) // d))        # first sample index with s + a*d >= hi
        return a, max(a, b)
    for (xs, xf, ys, yf, zs, zf) in boxes:
        ia, ib = rng(xs, xf, view.xs, view.dx, nxs)
        ja, jb = rng(ys, yf, view.ys, view.dy, nys)
        ka, kb = rng(zs, zf, view.zs + k0 * view.dz, view.dz, kn)
        block[ka:kb, ja:jb, ia:ib] = 1


def _paint_points(block, points, view, k0):
    """block[kk][j][i] = value for every (x, y, z, value) that falls on a sampled cell, in list order (later entries win,
    geometry_outputs.py:138-141)."""
    kn, nys, nxs = block.shape
    for (x, y, z, value) in points:
        qx, rx = divmod(x - view.xs, view.dx)
        qy, ry = divmod(y - view.ys, view.dy)
        qz, rz = divmod(z - view.zs - k0 * view.dz, view.dz)
        if rx or ry or rz or not (0 <= qx < nxs and 0 <= qy < nys and 0 <= qz < kn):
            continue
        block[qz, qy, qx] = value    # int8 element: numpy refuses a number that does not fit, like the reference's assignment


def _write_vti(self, G, pbar):
    from gprMax.geometry_outputs import GeometryView
    nxs, nys, nzs = _samples(self.xs, self.xf, self.dx), _samples(self.ys, self.yf, self.dy), _samples(self.zs, self.zf, self.dz)
    if (nxs, nys, nzs) != (self.vtk_nxcells, self.vtk_nycells, self.vtk_nzcells):
        # the reference's loop would run past the end of its output arrays here (geometry_outputs_ext.pyx:102-110)
        raise GeneralError('geometry view {}: the extent is not a whole number of steps in every direction'.format(self.basefilename))
    ncells = nxs * nys * nzs
    srcs = G.hertziandipoles + G.magneticdipoles + G.voltagesources + G.transmissionlines
    src_points = [(s.xcoord, s.ycoord, s.zcoord, index + 2) for index, s in enumerate(srcs)]
    rx_points = [(r.xcoord, r.ycoord, r.zcoord, index + 1) for index, r in enumerate(G.rxs)]
    pml_boxes = [(p.xs, p.xf, p.ys, p.yf, p.zs, p.zf) for p in G.pmls]
    u32 = np.dtype(np.uint32).itemsize
    hfmt, hbytes, hattr = _block_header([u32 * ncells])
    vtk_srcs_pml_offset = u32 * ncells + hbytes
    vtk_rxs_offset = u32 * ncells + hbytes + ncells + hbytes
    spacing = (self.dx * G.dx, self.dy * G.dy, self.dz * G.dz)
    ext = (self.vtk_xscells, self.vtk_xfcells, self.vtk_yscells, self.vtk_yfcells, self.vtk_zscells, self.vtk_zfcells)
    with open(self.filename, 'wb') as f:
        f.write('<?xml version="1.0"?>\n'.encode('utf-8'))
        f.write('<VTKFile type="ImageData" version="1.0" byte_order="{}"{}>\n'.format(GeometryView.byteorder, hattr).encode('utf-8'))
        f.write('<ImageData WholeExtent="{} {} {} {} {} {}" Origin="0 0 0" Spacing="{:.3} {:.3} {:.3}">\n'.format(*ext, *spacing).encode('utf-8'))
        f.write('<Piece Extent="{} {} {} {} {} {}">\n'.format(*ext).encode('utf-8'))
        f.write('<CellData Scalars="Material">\n'.encode('utf-8'))
        f.write('<DataArray type="UInt32" Name="Material" format="appended" offset="0" />\n'.encode('utf-8'))
        f.write('<DataArray type="Int8" Name="Sources_PML" format="appended" offset="{}" />\n'.format(vtk_srcs_pml_offset).encode('utf-8'))
        f.write('<DataArray type="Int8" Name="Receivers" format="appended" offset="{}" />\n'.format(vtk_rxs_offset).encode('utf-8'))
        f.write('</CellData>\n'.encode('utf-8'))
        f.write('</Piece>\n</ImageData>\n<AppendedData encoding="raw">\n_'.encode('utf-8'))

        # Material: G.solid sampled and re-ordered, a few z planes at a time
        f.write(pack(hfmt, u32 * ncells))
        pbar.update(n=4)
        for k0, kn in _kblocks(nzs, u32 * nxs * nys):
            block = transpose([G.solid], (self.xs, self.ys, self.zs + k0 * self.dz), (nxs, nys, kn), (self.dx, self.dy, self.dz))
            f.write(memoryview(block))
            pbar.update(n=block.nbytes)
        # Sources_PML (0 not set, 1 PML, sources from 2) and Receivers (from 1)
        for boxes, points in ((pml_boxes, src_points), ([], rx_points)):
            f.write(pack(hfmt, ncells))
            pbar.update(n=4)
            for k0, kn in _kblocks(nzs, nxs * nys):
                block = np.zeros((kn, nys, nxs), dtype=np.int8)
                _paint_boxes(block, boxes, self, k0)
                _paint_points(block, points, self, k0)
                f.write(memoryview(block))
                pbar.update(n=block.nbytes)
        f.write('\n</AppendedData>\n</VTKFile>'.encode('utf-8'))
        self.write_gprmax_info(f, G)


def _xchunks(n, bytes_per_plane):
    cb = max(1, BLOCK_BYTES // max(1, bytes_per_plane))
    for a in range(0, n, cb):
        yield a, min(n, a + cb)


def _write_vtp(self, G, pbar):
    from gprMax.geometry_outputs import GeometryView
    nx, ny, nz = self.nx, self.ny, self.nz
    u32 = np.dtype(np.uint32).itemsize
    P = (ny + 1) * (nz + 1)
    # define_fine_geometry takes dx, dy, dz as C floats and multiplies in float (geometry_outputs_ext.pyx:30-32, 54-56)
    dxf, dyf, dzf = np.float32(G.dx), np.float32(G.dy), np.float32(G.dz)
    with open(self.filename, 'wb') as f:
        f.write('<?xml version="1.0"?>\n'.encode('utf-8'))
        f.write('<VTKFile type="PolyData" version="1.0" byte_order="{}">\n'.format(GeometryView.byteorder).encode('utf-8'))
        f.write('<PolyData>\n<Piece NumberOfPoints="{}" NumberOfVerts="0" NumberOfLines="{}" NumberOfStrips="0" NumberOfPolys="0">\n'.format(self.vtk_numpoints, self.vtk_numlines).encode('utf-8'))
        f.write('<Points>\n<DataArray type="Float32" NumberOfComponents="3" format="appended" offset="0" />\n</Points>\n'.encode('utf-8'))
        f.write('<Lines>\n<DataArray type="UInt32" Name="connectivity" format="appended" offset="{}" />\n'.format(self.vtk_connectivity_offset).encode('utf-8'))
        f.write('<DataArray type="UInt32" Name="offsets" format="appended" offset="{}" />\n</Lines>\n'.format(self.vtk_offsets_offset).encode('utf-8'))
        f.write('<CellData Scalars="Material">\n'.encode('utf-8'))
        f.write('<DataArray type="UInt32" Name="Material" format="appended" offset="{}" />\n'.format(self.vtk_materials_offset).encode('utf-8'))
        f.write('</CellData>\n'.encode('utf-8'))
        f.write('</Piece>\n</PolyData>\n<AppendedData encoding="raw">\n_'.encode('utf-8'))

        # points: label = ((i-xs)*(ny+1) + (j-ys))*(nz+1) + (k-zs), coordinates (i*dx, j*dy, k*dz) in float32
        f.write(pack('I', 4 * 3 * self.vtk_numpoints))
        yc = np.arange(self.ys, self.yf + 1).astype(np.float32) * dyf
        zc = np.arange(self.zs, self.zf + 1).astype(np.float32) * dzf
        for a, b in _xchunks(nx + 1, 12 * P):
            pts = np.empty((b - a, ny + 1, nz + 1, 3), dtype=np.float32)
            pts[..., 0] = (np.arange(self.xs + a, self.xs + b).astype(np.float32) * dxf)[:, None, None]
            pts[..., 1] = yc[None, :, None]
            pts[..., 2] = zc[None, None, :]
            f.write(memoryview(pts))
            pbar.update(n=pts.nbytes)

        # connectivity: x lines (i < xf) join label and label + P, y lines (j < yf) label + nz + 1, z lines (k < zf) label + 1
        f.write(pack('I', u32 * self.vtk_numlines * self.vtk_numline_components))
        pbar.update(n=4)
        jk = (np.arange(ny + 1, dtype=np.int64) * (nz + 1))[:, None] + np.arange(nz + 1, dtype=np.int64)[None, :]
        for (ni, sel, other) in ((nx, jk, P), (nx + 1, jk[:ny, :], nz + 1), (nx + 1, jk[:, :nz], 1)):
            sel = sel.reshape(-1)
            for a, b in _xchunks(ni, 2 * u32 * max(1, sel.size)):
                first = (np.arange(a, b, dtype=np.int64) * P)[:, None] + sel[None, :]
                lines = np.empty(first.shape + (2,), dtype=np.uint32)
                lines[..., 0] = first       # the reference stores its 64-bit counters into uint32 elements the same way
                lines[..., 1] = first + other
                f.write(memoryview(lines))
                pbar.update(n=lines.nbytes)

        # cell (line) offsets: 2, 4, ..., 2 * numlines
        f.write(pack('I', u32 * self.vtk_numlines))
        pbar.update(n=4)
        step = max(1, BLOCK_BYTES // u32)
        for a in range(0, self.vtk_numlines, step):
            b = min(self.vtk_numlines, a + step)
            offs = (np.arange(a + 1, b + 1, dtype=np.int64) * self.vtk_numline_components).astype(np.uint32)
            f.write(memoryview(offs))
            pbar.update(n=offs.nbytes)

        # materials per edge: slices of G.ID in their own (C) order
        f.write(pack('I', u32 * self.vtk_numlines))
        pbar.update(n=4)
        xs, ys, zs, xf, yf, zf = self.xs, self.ys, self.zs, self.xf, self.yf, self.zf
        for comp, (i1, j1, k1) in enumerate(((xf, yf + 1, zf + 1), (xf + 1, yf, zf + 1), (xf + 1, yf + 1, zf))):
            for a, b in _xchunks(i1 - xs, u32 * max(1, (j1 - ys) * (k1 - zs))):
                part = np.ascontiguousarray(G.ID[comp, xs + a:xs + b, ys:j1, zs:k1])
                f.write(memoryview(part))
                pbar.update(n=part.nbytes)

        f.write('\n</AppendedData>\n</VTKFile>'.encode('utf-8'))
        self.write_gprmax_info(f, G, materialsonly=True)


def write_vtk(self, G, pbar):
    """Drop-in for `GeometryView.write_vtk` (geometry_outputs.py:119-290)."""
    if self.fileext == '.vti':
        _write_vti(self, G, pbar)
    elif self.fileext == '.vtp':
        _write_vtp(self, G, pbar)


_installed = False


def installed():
    return _installed


def install():
    """Patch the two writer methods on the reference's classes (idempotent)."""
    global _installed
    if _installed:
        return
    from gprMax.geometry_outputs import GeometryView
    from gprMax.snapshots import Snapshot
    GeometryView.write_vtk = write_vtk
    Snapshot.write_vtk_imagedata = write_vtk_imagedata
    _installed = True
