"""Multi-threaded build of the per-edge material ID array -- drop-in for the reference's
`build_electric_components` + `build_magnetic_components` (gprMax/yee_cell_build_ext.pyx:110-257, called at
model_build_run.py:216-218), SURVEY.md 8(f) rank 1.

The 6 N-edge scan runs on all host cores inside libgprmax_b200.so (csrc/gpb_idbuild.cpp); the order-dependent part -- which
dielectric-smoothed material a combination of neighbouring cell materials becomes, and which number it gets -- is NOT
re-implemented: for every distinct combination, in the reference's scan order, the reference's own `create_electric_average`
/ `create_magnetic_average` is called once on the first edge that shows it.  G.ID and G.materials come out identical to the
reference's build (tests/test_yee_build.py).
"""
import ctypes as C

import numpy as np

from . import _lib
from .exceptions import GeneralError


def build_components(G, create_electric_average=None, create_magnetic_average=None, x_range=None):
    """Fill G.ID from G.solid / G.rigidE / G.rigidH (both the electric and the magnetic components).

    create_*_average: the reference's functions (default: imported from gprMax.yee_cell_build_ext).
    x_range = (x0, x1): only the node planes [x0, x1) -- a shard builds just its slab of G.ID (the distinct combinations of
    all shards have to be resolved in plane order, see `resolve`)."""
    if create_electric_average is None:
        from gprMax.yee_cell_build_ext import create_electric_average, create_magnetic_average
    combos = scan(G, x_range)
    numid = resolve(G, combos, create_electric_average, create_magnetic_average)
    apply(G, combos, numid, x_range)


def _arrays(G):
    solid, rE, rH, ID = G.solid, G.rigidE, G.rigidH, G.ID
    for a, dt in ((solid, np.uint32), (rE, np.int8), (rH, np.int8), (ID, np.uint32)):
        if a.dtype != dt or not a.flags['C_CONTIGUOUS']:
            raise GeneralError('geometry arrays must be C-contiguous uint32 / int8 as the reference allocates them (grid.py:157-169)')
    return solid, rE, rH, ID


def scan(G, x_range=None):
    """Pass 1: agreeing edges written, distinct disagreeing combinations returned in scan order (list of _lib.IdCombo)."""
    L = _lib.lib()
    solid, rE, rH, ID = _arrays(G)
    x0, x1 = (0, G.nx + 1) if x_range is None else x_range
    cap = 4096
    while True:
        buf = (_lib.IdCombo * cap)()
        n = C.c_int(0)
        rc = L.gpb_ids_scan(solid.ctypes.data, rE.ctypes.data, rH.ctypes.data, ID.ctypes.data, int(G.nx), int(G.ny), int(G.nz), int(x0), int(x1), buf, cap, C.byref(n))
        if rc == 2:
            cap = n.value + 16
            continue
        if rc:
            raise GeneralError('gpb_ids_scan failed ({})'.format(rc))
        return [buf[q] for q in range(n.value)]


def resolve(G, combos, create_electric_average, create_magnetic_average):
    """The order-dependent part, by the reference's own functions: one call per distinct combination, in scan order, on the
    edge recorded for it; returns the material numbers (and G.materials has grown exactly as in the reference)."""
    numid = np.zeros(len(combos), dtype=np.uint32)
    for q, c in enumerate(combos):
        if c.comp < 3:
            create_electric_average(c.i, c.j, c.k, int(c.id[0]), int(c.id[1]), int(c.id[2]), int(c.id[3]), c.comp, G)
        else:
            create_magnetic_average(c.i, c.j, c.k, int(c.id[0]), int(c.id[1]), c.comp, G)
        numid[q] = G.ID[c.comp, c.i, c.j, c.k]
    return numid


def apply(G, combos, numid, x_range=None):
    """Pass 2: the resolved material on every disagreeing edge."""
    L = _lib.lib()
    solid, rE, rH, ID = _arrays(G)
    x0, x1 = (0, G.nx + 1) if x_range is None else x_range
    arr = (_lib.IdCombo * max(1, len(combos)))(*combos)
    rc = L.gpb_ids_apply(solid.ctypes.data, rE.ctypes.data, rH.ctypes.data, ID.ctypes.data, int(G.nx), int(G.ny), int(G.nz), int(x0), int(x1),
                         arr, numid.ctypes.data, len(combos))
    if rc:
        raise GeneralError('gpb_ids_apply failed ({})'.format(rc))


# ------------------------------------------------------------------------------------------------- slab-local (sharded) build

class _MaterialsOnly(object):
    """What `create_electric_average` / `create_magnetic_average` touch of a grid (yee_cell_build_ext.pyx:31-107): the material
    list they read and extend, and one element of `ID` they write the resulting number into."""

    def __init__(self, materials):
        self.materials = materials
        self.ID = np.zeros((6, 1, 1, 1), dtype=np.uint32)


def resolve_materials(materials, combos, create_electric_average, create_magnetic_average):
    """`resolve` without a grid: extends `materials` exactly as the reference's build would and returns the material number of
    every combination.  `combos`: (comp, i, j, k, id0, id1, id2, id3) tuples or _lib.IdCombo, already in scan order."""
    proxy = _MaterialsOnly(materials)
    numid = np.zeros(len(combos), dtype=np.uint32)
    for q, c in enumerate(combos):
        comp, ids = (c.comp, [int(v) for v in c.id]) if isinstance(c, _lib.IdCombo) else (int(c[0]), [int(v) for v in c[4:8]])
        if comp < 3:
            create_electric_average(0, 0, 0, ids[0], ids[1], ids[2], ids[3], comp, proxy)
        else:
            create_magnetic_average(0, 0, 0, ids[0], ids[1], comp, proxy)
        numid[q] = proxy.ID[comp, 0, 0, 0]
    return numid


def merge_combos(per_slab):
    """The distinct combinations of all slabs in the reference's scan order (component, i, j, k of the first edge that shows
    each).  `per_slab`: one list of (comp, i, j, k, id0..id3) tuples per slab."""
    seen, merged = set(), []
    for c in sorted((tuple(int(v) for v in c) for slab in per_slab for c in slab), key=lambda c: c[:4]):
        key = (c[0],) + c[4:8]
        if key not in seen:
            seen.add(key)
            merged.append(c)
    return merged


def build_slab(G, x_range, solid_x0, id_x0, gather=None, create_electric_average=None, create_magnetic_average=None):
    """Build the node planes `x_range` = (x0, x1) of a model of G.nx x G.ny x G.nz cells on a rank that holds only a slab of
    the geometry: G.solid / G.rigidE / G.rigidH carry the cell planes from `solid_x0` on, G.ID the node planes from `id_x0` on
    (they must cover the cell planes [x0 - 1, x1) and the node planes [x0, x1)).

    gather(list) -> list of lists: exchanges the slabs' distinct combinations (a few hundred small tuples), e.g.
    `torch.distributed.all_gather_object`; None for a single slab.  Every rank then resolves ALL combinations in the reference's
    order with the reference's own `create_*_average`, so `G.materials` grows identically everywhere and equals the reference's
    list, and the slabs' ID planes equal the reference's `G.ID[:, x0:x1]` (tests/test_yee_build.py, tests/test_sharded_cpu.py)."""
    if create_electric_average is None:
        from gprMax.yee_cell_build_ext import create_electric_average, create_magnetic_average
    L = _lib.lib()
    solid, rE, rH, ID = _arrays(G)
    x0, x1 = x_range
    dims = (int(G.nx), int(G.ny), int(G.nz), int(solid_x0), int(solid.shape[0]), int(id_x0), int(ID.shape[1]), int(x0), int(x1))
    if rE.shape[1] != solid.shape[0] or rH.shape[1] != solid.shape[0] or solid.shape[1:] != (G.ny, G.nz) or ID.shape[2:] != (G.ny + 1, G.nz + 1):
        raise GeneralError('slab arrays do not match the domain')
    cap = 4096
    while True:
        buf = (_lib.IdCombo * cap)()
        n = C.c_int(0)
        rc = L.gpb_ids_scan_slab(solid.ctypes.data, rE.ctypes.data, rH.ctypes.data, ID.ctypes.data, *dims, buf, cap, C.byref(n))
        if rc == 2:
            cap = n.value + 16
            continue
        if rc:
            raise GeneralError('gpb_ids_scan_slab failed ({}): do the slab arrays cover the planes {}..{}?'.format(rc, x0, x1))
        break
    mine = [(buf[q].comp, buf[q].i, buf[q].j, buf[q].k) + tuple(int(v) for v in buf[q].id) for q in range(n.value)]
    merged = merge_combos(gather(mine) if gather is not None else [mine])
    numid = resolve_materials(G.materials, merged, create_electric_average, create_magnetic_average)
    arr = (_lib.IdCombo * max(1, len(merged)))()
    for q, c in enumerate(merged):
        arr[q].comp, arr[q].i, arr[q].j, arr[q].k = c[:4]
        for t in range(4):
            arr[q].id[t] = c[4 + t]
    rc = L.gpb_ids_apply_slab(solid.ctypes.data, rE.ctypes.data, rH.ctypes.data, ID.ctypes.data, *dims, arr, numid.ctypes.data, len(merged))
    if rc:
        raise GeneralError('gpb_ids_apply_slab failed ({})'.format(rc))
    return len(merged)
