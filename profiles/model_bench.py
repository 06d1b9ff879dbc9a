"""Device loop time of a fixture model (tests/golden/<name>_<variant>.npz): python profiles/model_bench.py name [variant]"""
import sys
sys.path.insert(0, ".")
from gprmax_b200 import Solver, load_model
name = sys.argv[1]; variant = sys.argv[2] if len(sys.argv) > 2 else 'f32'
G, _ = load_model('tests/golden/{}_{}.npz'.format(name, variant))
sv = Solver(G, device_id=0)
sv.run(); sv.reset(); sv.run()
cells = G.nx * G.ny * G.nz
print('{} {}: {} cells x {} its, maxpoles {}: loop {:.3f} s -> {:.0f} Mcells/s ({:.1f} us/iteration), {} launches'.format(
    name, variant, cells, G.iterations, G.maxpoles, sv.elapsed, cells * G.iterations / sv.elapsed / 1e6, sv.elapsed / G.iterations * 1e6, sv.kernel_launches))
sv.close()
