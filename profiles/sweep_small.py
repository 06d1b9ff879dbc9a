"""Small / mid-size grids: device loop time per iteration under a few launch-side switches, all in one process (the switches are
read in gpb_create).   python profiles/sweep_small.py > gpurun_out/sweep_small.jsonl"""
import json
import os
import sys
import time

sys.path.insert(0, ".")
import numpy as np
from gprmax_b200 import Solver, load_model
from benchkit.synthetic import bench_model

SWITCHES = ('GPB_PDL', 'GPB_FUSE_BEGIN', 'GPB_GRAPH_ITERS', 'GPB_V4_XCHUNK', 'GPB_NO_TMA', 'GPB_FORCE_TMA', 'GPB_TMA_XCHUNK', 'GPB_TMA_TZ', 'GPB_TMA_TY', 'GPB_NO_GRAPH')


def run(G, env, reference=None):
    for k in SWITCHES:
        os.environ.pop(k, None)
    os.environ.update(env)
    try:
        with Solver(G, device_id=0) as sv:
            path = sv.kernel_path
            sv.run()
            sv.reset()
            sv.run()
            t = sv.elapsed
            rx = sv.receivers()
    except Exception as e:   # a tile shape that is not instantiated, ...
        return {'env': env, 'error': str(e)[:200]}, None
    cells = G.nx * G.ny * G.nz
    d = {'env': env, 'us_per_iteration': round(t / G.iterations * 1e6, 2), 'mcells_per_s': round(cells * G.iterations / t / 1e6, 1), 'kernels': path}
    if reference is not None:
        d['bit_identical_with_default'] = bool(np.array_equal(rx, reference))
    return d, rx


def main():
    t00 = time.time()
    budget = float(os.environ.get('SWEEP_SECONDS', '100'))
    plan_name = os.environ.get('SWEEP_PLAN', 'pdl')
    graph = [{'GPB_GRAPH_ITERS': '1'}, {'GPB_GRAPH_ITERS': '4'}, {'GPB_NO_GRAPH': '1'}]
    v4 = [{'GPB_V4_XCHUNK': '2'}, {'GPB_V4_XCHUNK': '4'}, {'GPB_V4_XCHUNK': '8'}]
    pdl = [{'GPB_FUSE_BEGIN': '0'}, {'GPB_PDL': '0'}, {'GPB_PDL': '0', 'GPB_GRAPH_ITERS': '1'}]
    if plan_name == 'pdl':
        plan = [(m, pdl) for m in ('cylinder_Ascan_2D_f32', 'bench:100', 'bench:150', 'bench:200', 'bench:300', 'heterogeneous_soil_full_f32', 'transmission_line_f32')]
    else:
        plan = [
            ('cylinder_Ascan_2D_f32', graph + v4),
            ('bench:100', graph + v4 + [{'GPB_FORCE_TMA': '1'}]),
            ('bench:150', graph + [{'GPB_NO_TMA': '1'}, {'GPB_NO_TMA': '1', 'GPB_V4_XCHUNK': '8'}, {'GPB_TMA_XCHUNK': '2'}, {'GPB_TMA_XCHUNK': '8'}, {'GPB_TMA_TZ': '64', 'GPB_TMA_TY': '16'}]),
            ('bench:200', graph[:2] + [{'GPB_NO_TMA': '1'}, {'GPB_TMA_XCHUNK': '4'}, {'GPB_TMA_XCHUNK': '16'}]),
            ('bench:300', graph[:2]),
        ]
    for spec, envs in plan:
        if spec.startswith('bench:'):
            G = bench_model(int(spec[6:]), iterations=300)
        else:
            G, _ = load_model('tests/golden/' + spec + '.npz')
        base, ref = run(G, {})
        base['model'] = spec
        print(json.dumps(base), flush=True)
        for env in envs:
            if time.time() - t00 > budget:
                print(json.dumps({'stopped': 'time budget', 'model': spec}), flush=True)
                return
            d, _ = run(G, env, ref)
            d['model'] = spec
            print(json.dumps(d), flush=True)


if __name__ == '__main__':
    main()
