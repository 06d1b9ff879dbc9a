"""Where the 300^3 step goes: the same box with 10-cell PML on all faces, on one pair of faces only, and with no PML at all."""
import sys
sys.path.insert(0, ".")
from gprmax_b200 import Solver
from benchkit.synthetic import homogeneous_model
size = int(sys.argv[1]) if len(sys.argv) > 1 else 300
its = 300
for label, keep in (('all six slabs', 'xyz'), ('x slabs only', 'x'), ('y slabs only', 'y'), ('z slabs only', 'z'), ('no PML', '')):
    G = homogeneous_model(size, iterations=its, pml_cells=10)
    G.pmls = [p for p in G.pmls if p.direction[0] in keep]
    sv = Solver(G, device_id=0)
    sv.run(); sv.reset(); sv.run()
    el = sv.elapsed
    sv.reset(); sv.profile(20); pr = sv.profile(100)
    print('size %d %-14s: %.1f Mcells/s (%.3f ms/iteration) %s' % (size, label, size**3 * its / el / 1e6, el / its * 1e3, {k: round(v / 100, 4) for k, v in pr.items()}))
    sv.close()
