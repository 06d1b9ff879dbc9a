"""Host-side timing of the streaming VTK writers against the reference's own, CPU only:

    python profiles/vtk_writers_bench.py [n]      # n^3 cells (default 300), writes into a temporary directory

Prints one JSON line per file type: seconds with the reference's writer / with gprmax_b200.vtk_writers, identical or not,
and the peak resident memory of each (separate processes)."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys, time, types, resource, hashlib
sys.path.insert(0, {root!r})
import numpy as np
import baseline
baseline.use_reference()
from gprMax.geometry_outputs import GeometryView
from gprMax.snapshots import Snapshot
from gprMax.constants import floattype
from gprmax_b200 import vtk_writers
kind, impl, n, out = {kind!r}, {impl!r}, {n}, {out!r}
class Bar(object):
    def update(self, n=1):
        pass
rng = np.random.default_rng(1)
G = types.SimpleNamespace(nx=n, ny=n, nz=n, dx=0.001, dy=0.001, dz=0.001)
G.pmls = [types.SimpleNamespace(xs=0, xf=10, ys=0, yf=n, zs=0, zf=n)]
G.hertziandipoles = [types.SimpleNamespace(xcoord=n // 2, ycoord=n // 2, zcoord=n // 2, ID='src')]
G.magneticdipoles, G.voltagesources, G.transmissionlines = [], [], []
G.rxs = [types.SimpleNamespace(xcoord=n // 3, ycoord=n // 2, zcoord=n // 2, ID='rx')]
G.materials = [types.SimpleNamespace(ID='m{{}}'.format(q), numID=q) for q in range(8)]
base = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
if kind == 'vti':
    G.solid = rng.integers(0, 8, size=(n, n, n), dtype=np.uint32)
    v = GeometryView(0, 0, 0, n, n, n, 1, 1, 1, 'v', '.vti')
elif kind == 'vtp':
    G.ID = rng.integers(0, 8, size=(6, n + 1, n + 1, n + 1), dtype=np.uint32)
    v = GeometryView(0, 0, 0, n, n, n, 1, 1, 1, 'v', '.vtp')
else:
    v = Snapshot(0, 0, 0, n, n, n, 1, 1, 1, 10, 's')
    fields = [rng.standard_normal((n, n, n), dtype=np.float32).astype(floattype) for _ in range(6)]
v.filename = out
base = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
t0 = time.perf_counter()
if kind == 'snap':
    if impl == 'ref':
        # what solve_gpu of the reference does before the writer runs (snapshots.py:223-228)
        v.electric = np.stack(fields[:3]).reshape(-1, order='F')
        v.magnetic = np.stack(fields[3:]).reshape(-1, order='F')
        Snapshot.write_vtk_imagedata(v, Bar(), G)
    else:
        v.fields = fields
        vtk_writers.write_vtk_imagedata(v, Bar(), G)
else:
    (GeometryView.write_vtk if impl == 'ref' else vtk_writers.write_vtk)(v, G, Bar())
t = time.perf_counter() - t0
h = hashlib.sha256()
with open(out, 'rb') as f:
    for chunk in iter(lambda: f.read(1 << 24), b''):
        h.update(chunk)
print(t, (resource.getrusage(resource.RUSAGE_SELF).ru_maxrss - base) / 1024.0, os.path.getsize(out), h.hexdigest())
'''


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    with tempfile.TemporaryDirectory() as tmp:
        for kind, size in (('vti', n), ('vtp', max(8, n // 2)), ('snap', n)):
            row = {'file': kind, 'cells': size ** 3}
            digests = []
            for impl in ('ref', 'new'):
                out = os.path.join(tmp, '{}_{}.bin'.format(kind, impl))
                r = subprocess.run([sys.executable, '-c', CHILD.format(root=ROOT, kind=kind, impl=impl, n=size, out=out)], capture_output=True, text=True)
                if r.returncode:
                    sys.stderr.write(r.stderr)
                    return 1
                t, rss, nbytes, digest = r.stdout.split()[-4:]
                row[impl + '_s'] = round(float(t), 3)
                row[impl + '_extra_rss_mb'] = round(float(rss), 1)
                row['file_mb'] = round(int(nbytes) / 1e6, 1)
                digests.append(digest)
                os.remove(out)
            row['identical'] = digests[0] == digests[1]
            row['speedup'] = round(row['ref_s'] / row['new_s'], 1)
            print(json.dumps(row), flush=True)
    return 0


if __name__ == '__main__':
    sys.exit(main())
