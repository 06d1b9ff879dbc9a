"""Short profiling driver: 300^3 benchmark model, N plain-launch iterations (run under ncu)."""
import sys
sys.path.insert(0, ".")
from gprmax_b200 import Solver
from benchkit.synthetic import bench_model
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
size = int(sys.argv[2]) if len(sys.argv) > 2 else 300
G = bench_model(size, iterations=n)
sv = Solver(G, device_id=0)
sv.profile(n)
sv.close()
