"""Small and mid-size grids: device loop time per iteration with the default kernels and with the opt-in cooperative whole-run kernel
(GPB_COOP=1).   python profiles/small_bench.py"""
import json, os, subprocess, sys
sys.path.insert(0, ".")
CODE = r'''
import sys, json
sys.path.insert(0, ".")
from gprmax_b200 import Solver, load_model
from benchkit.synthetic import bench_model
spec = sys.argv[1]
if spec.startswith('bench:'):
    n = int(spec[6:]); G = bench_model(n, iterations=400)
else:
    G, _ = load_model('tests/golden/' + spec + '.npz')
with Solver(G, device_id=0) as sv:
    path = sv.kernel_path
    sv.run(); sv.reset(); sv.run()
    t = sv.elapsed
cells = G.nx * G.ny * G.nz
print(json.dumps({'model': spec, 'cells': cells, 'iterations': G.iterations, 'us_per_iteration': t / G.iterations * 1e6, 'mcells_per_s': cells * G.iterations / t / 1e6, 'kernels': path}))
'''
for spec in ('cylinder_Ascan_2D_f32', 'bench:100', 'bench:120', 'bench:150', 'heterogeneous_soil_full_f32', 'bscan_gssi_trace1_f32'):
    for env in ({}, {'GPB_COOP': '1'}):
        r = subprocess.run([sys.executable, '-c', CODE, spec], stdout=subprocess.PIPE, stderr=subprocess.PIPE, universal_newlines=True, env=dict(os.environ, **env))
        line = (r.stdout.strip().splitlines() or ['{}'])[-1]
        try:
            d = json.loads(line)
            d['env'] = env
            print(json.dumps(d), flush=True)
        except Exception:
            print('FAILED', spec, env, r.stderr[-500:], flush=True)
