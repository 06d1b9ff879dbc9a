"""Quick device-side timing of the 300^3 benchmark loop (no CPU baseline, no e2e): prints Mcells/s and per-kernel ms."""
import sys
sys.path.insert(0, ".")
from gprmax_b200 import Solver
from benchkit.synthetic import bench_model
size = int(sys.argv[1]) if len(sys.argv) > 1 else 300
its = int(sys.argv[2]) if len(sys.argv) > 2 else 400
G = bench_model(size, iterations=its)
sv = Solver(G, device_id=0)
sv.run(); sv.reset(); sv.run()
t = sv.elapsed
print('size %d its %d: %.1f Mcells/s  (%.3f ms/iteration)' % (size, its, size**3 * its / t / 1e6, t / its * 1e3))
np_ = min(100, its - 20)
sv.reset(); sv.profile(20); pr = sv.profile(np_)
print({k: round(v / np_, 4) for k, v in pr.items()})
sv.close()
