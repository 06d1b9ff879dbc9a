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
