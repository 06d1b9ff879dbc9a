"""Host-side build steps next to the hot path: the multi-threaded ID build (gprmax_b200/yee_build.py +
csrc/gpb_idbuild.cpp) and the vectorised PML build (gprmax_b200/pml_build.py: every R table bit-identical) against the reference's own
build_electric_components / build_magnetic_components (yee_cell_build_ext.pyx:110-257) on real models: the ID array and the
list of dielectric-smoothed materials (names, numbering, averaged properties) must be IDENTICAL.

Needs the vendored reference (baseline/_ref, built by baseline/install_ref.sh in the build container); CPU only."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'gprMax')), reason='baseline/_ref not installed')

SCRIPT = r'''
import os, sys, io, contextlib, tempfile, time
sys.path.insert(0, {root!r})
import numpy as np
import baseline
baseline.use_reference()
import gprMax.gprMax as top
import gprMax.model_build_run as mbr
from gprMax import yee_cell_build_ext as ref
from gprmax_b200 import yee_build

class Done(Exception):
    pass

res = {{}}
def electric(solid, rigidE, ID, G):
    ID0 = G.ID.copy()
    nmat0 = len(G.materials)
    t0 = time.perf_counter()
    ref.build_electric_components(G.solid, G.rigidE, G.ID, G)
    ref.build_magnetic_components(G.solid, G.rigidH, G.ID, G)
    res['t_ref'] = time.perf_counter() - t0
    ID_ref = G.ID.copy()
    mats_ref = [(m.numID, m.ID, m.type, float(m.er), float(m.se), float(m.mr), float(m.sm)) for m in G.materials]
    G.ID[...] = ID0
    del G.materials[nmat0:]
    t0 = time.perf_counter()
    yee_build.build_components(G)
    res['t_new'] = time.perf_counter() - t0
    mats_new = [(m.numID, m.ID, m.type, float(m.er), float(m.se), float(m.mr), float(m.sm)) for m in G.materials]
    res.update(same_id=bool(np.array_equal(G.ID, ID_ref)), same_mats=mats_new == mats_ref, nmat=len(mats_ref), nmat0=nmat0, cells=G.nx * G.ny * G.nz,
               averaged=int((ID_ref != ID0).sum()))
    # slab-wise: the same result when the planes are built in two separate calls (the combinations are scanned per slab and
    # resolved together in plane order, as a sharded build would)
    G.ID[...] = ID0
    del G.materials[nmat0:]
    cut = (G.nx + 1) // 2
    combos = yee_build.scan(G, (0, cut)) + yee_build.scan(G, (cut, G.nx + 1))
    seen, merged = set(), []
    for c in sorted(combos, key=lambda c: (c.comp, c.i, c.j, c.k)):
        key = (c.comp, tuple(c.id))
        if key not in seen:
            seen.add(key)
            merged.append(c)
    numid = yee_build.resolve(G, merged, ref.create_electric_average, ref.create_magnetic_average)
    yee_build.apply(G, merged, numid, (0, cut))
    yee_build.apply(G, merged, numid, (cut, G.nx + 1))
    res['same_id_slabs'] = bool(np.array_equal(G.ID, ID_ref))
    # slab-LOCAL: three "ranks" (threads) that each hold only their x range of solid / rigid / ID (plus the one cell plane to the
    # left that the edges of their first node plane look at), exchange their distinct combinations and resolve them on their own
    # copy of the material list (yee_build.build_slab)
    import copy, threading, types
    del G.materials[nmat0:]
    nranks = 1 if G.nx < 3 else 3
    cuts = [0] + [(G.nx + 1) * r // nranks for r in range(1, nranks)] + [G.nx + 1]
    slots, barrier = [None] * nranks, threading.Barrier(nranks)
    slabs = []
    for r in range(nranks):
        x0, x1 = cuts[r], cuts[r + 1]
        s0, s1 = max(x0 - 1, 0), min(x1, G.nx)
        slabs.append(types.SimpleNamespace(nx=G.nx, ny=G.ny, nz=G.nz, x0=x0, x1=x1, s0=s0,
                                           solid=np.ascontiguousarray(G.solid[s0:s1]), rigidE=np.ascontiguousarray(G.rigidE[:, s0:s1]),
                                           rigidH=np.ascontiguousarray(G.rigidH[:, s0:s1]), ID=np.ascontiguousarray(ID0[:, x0:x1]),
                                           materials=copy.deepcopy(G.materials)))
    errors = []
    def rank(r):
        def gather(mine):
            slots[r] = mine
            barrier.wait()
            return list(slots)
        try:
            sl = slabs[r]
            yee_build.build_slab(sl, (sl.x0, sl.x1), sl.s0, sl.x0, gather if nranks > 1 else None, ref.create_electric_average, ref.create_magnetic_average)
        except Exception as e:
            errors.append(repr(e))
            barrier.abort()
    threads = [threading.Thread(target=rank, args=(r,)) for r in range(nranks)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    same = not errors
    for sl in slabs:
        same = same and bool(np.array_equal(sl.ID, ID_ref[:, sl.x0:sl.x1]))
        same = same and [(m.numID, m.ID, m.type, float(m.er), float(m.se), float(m.mr), float(m.sm)) for m in sl.materials] == mats_ref
    res['same_id_slab_local'] = same
    res['slab_errors'] = errors
    raise Done()

from gprmax_b200 import pml_build
ref_build_pmls = mbr.build_pmls
def pmls(G, pbar):
    class Quiet(object):
        def update(self, *a):
            pass
    t0 = time.perf_counter()
    ref_build_pmls(G, Quiet())
    res['t_pml_ref'] = time.perf_counter() - t0
    names = ('ERA', 'ERB', 'ERE', 'ERF', 'HRA', 'HRB', 'HRE', 'HRF')
    tabs_ref = [[getattr(p, n).copy() for n in names] + [p.direction, p.xs, p.xf, p.ys, p.yf, p.zs, p.zf, p.thickness] for p in G.pmls]
    cfs_ref = [float(c.sigma.max) for c in G.cfs]
    del G.pmls[:]
    t0 = time.perf_counter()
    pml_build.build_pmls(G, pbar)
    res['t_pml_new'] = time.perf_counter() - t0
    tabs_new = [[getattr(p, n).copy() for n in names] + [p.direction, p.xs, p.xf, p.ys, p.yf, p.zs, p.zf, p.thickness] for p in G.pmls]
    same = len(tabs_ref) == len(tabs_new)
    for a, b in zip(tabs_ref, tabs_new):
        same = same and all(np.array_equal(x, y) for x, y in zip(a[:8], b[:8])) and a[8:] == b[8:]
    res['same_pml'] = bool(same) and cfs_ref == [float(c.sigma.max) for c in G.cfs]
mbr.build_pmls = pmls
mbr.build_electric_components = electric
work = tempfile.mkdtemp()
text = open(os.path.join(baseline.REF_DIR, {rel!r})).read()
for a, b in {repl!r}:
    assert a in text, a
    text = text.replace(a, b)
path = os.path.join(work, 'model.in')
open(path, 'w').write(text)
out = io.StringIO()
try:
    with contextlib.redirect_stdout(out):
        top.api(path, n=1)
except Done:
    pass
print('RESULT', res)
assert res['same_id'] and res['same_mats'] and res['same_id_slabs'] and res['same_id_slab_local'] and res['same_pml'], res
print('YEE_OK')
'''

MODELS = [
    ('user_models/cylinder_Ascan_2D.in', []),
    ('user_models/heterogeneous_soil.in', [('my_soil my_soil_box\n', 'my_soil my_soil_box 7\n'), ('0.065 0.080 my_soil_box\n', '0.065 0.080 my_soil_box 3\n'),
                                           ('#domain: 0.15 0.15 0.1', '#domain: 0.08 0.08 0.06'), ('#fractal_box: 0 0 0 0.15 0.15 0.070', '#fractal_box: 0 0 0 0.08 0.08 0.040'),
                                           ('#add_surface_roughness: 0 0 0.070 0.15 0.15 0.070 1.5 1 1 0.065 0.080', '#add_surface_roughness: 0 0 0.040 0.08 0.08 0.040 1.5 1 1 0.035 0.045'),
                                           ('#rx: 0.105 0.075 0.085', '#rx: 0.050 0.040 0.050'), ('#hertzian_dipole: y 0.045 0.075 0.085', '#hertzian_dipole: y 0.030 0.040 0.050'),
                                           ('#geometry_view', '##geometry_view')]),
    ('user_models/cylinder_Bscan_GSSI_1500.in', []),
]


@pytest.mark.parametrize('rel,repl', MODELS, ids=[m[0].split('/')[-1] for m in MODELS])
def test_id_build_identical_to_reference(rel, repl):
    r = subprocess.run([sys.executable, '-c', SCRIPT.format(root=ROOT, rel=rel, repl=repl)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       universal_newlines=True, timeout=900, cwd='/tmp')
    print(r.stdout[-1500:])
    assert r.returncode == 0 and 'YEE_OK' in r.stdout, r.stdout[-3000:]
