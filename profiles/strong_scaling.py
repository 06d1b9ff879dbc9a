"""Strong scaling of ONE model over the GPUs of a box, one process (gpb_create_sharded: linked x-slabs, halo pushed over peer
memory, one CUDA graph per slab and iteration):  python profiles/strong_scaling.py [size] [iterations]
Prints one JSON line per device count with Mcells/s and whether receivers and final fields equal the 1-GPU run bit for bit."""
import json, sys
sys.path.insert(0, ".")
import numpy as np
from gprmax_b200 import Solver
from gprmax_b200.gpu import device_count
from benchkit.synthetic import bench_model

size = int(sys.argv[1]) if len(sys.argv) > 1 else 300
its = int(sys.argv[2]) if len(sys.argv) > 2 else 300
G = bench_model(size, iterations=its)
ref = None
n = 1
while n <= device_count():
    with Solver(G, devices=list(range(n))) as sv:
        sv.run(); sv.reset(); sv.run()
        t = sv.elapsed
        rx = sv.receivers()
        fields = [sv.get_field(c) for c in (1, 4)]
        path = sv.kernel_path
    if ref is None:
        ref = (rx, fields)
    same = bool(np.array_equal(rx, ref[0]) and all(np.array_equal(a, b) for a, b in zip(fields, ref[1])))
    print(json.dumps({'workload': 'bench_{0}x{0}x{0} ({1} iterations), one domain over {2} GPU(s)'.format(size, its, n), 'gpus': n,
                      'mcells_per_s': size**3 * its / t / 1e6, 'us_per_iteration': t / its * 1e6, 'bitexact_vs_1gpu': same, 'kernels': path}), flush=True)
    n *= 2
