import sys
sys.path.insert(0, ".")
import numpy as np
from gprmax_b200 import Solver
from benchkit.synthetic import bench_model
for size in (200, 300):
    G = bench_model(size, real=np.float64, iterations=200)
    sv = Solver(G, device_id=0)
    sv.run(); sv.reset(); sv.run()
    print('f64 size %d: %.1f Mcells/s (%.3f ms/iteration)' % (size, size**3 * 200 / sv.elapsed / 1e6, sv.elapsed / 200 * 1e3))
    sv.close()
