"""BASELINE.json configs[3] for real: user_models/cylinder_Bscan_GSSI_1500.in as a B-scan through the REFERENCE's own front end
(baseline/_ref: input parsing, antenna macro, geometry / material / PML build, run_model, output file per trace), farmed one
trace per GPU (gprmax_b200.farm.run_bscan), the time loop on this core.

    python profiles/bscan_farm.py [traces] [gpus]          default: 54 traces on all visible GPUs
Prints one JSON line: traces/hour, per-trace host time (parse + build + write) and solve time, and the parity of trace 1
against the golden of the unmodified reference CPU solver (tests/golden/bscan_gssi_trace1_f32.npz)."""
import json, os, shutil, sys, tempfile, time
sys.path.insert(0, ".")
import numpy as np
import baseline
from gprmax_b200.farm import run_bscan
from gprmax_b200.gpu import device_count

if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 54
    ngpu = int(sys.argv[2]) if len(sys.argv) > 2 else device_count()
    work = tempfile.mkdtemp(prefix='bscan_')
    src = os.path.join(baseline.REF_DIR, 'user_models', 'cylinder_Bscan_GSSI_1500.in')
    shutil.copy(src, work)
    inp = os.path.join(work, 'cylinder_Bscan_GSSI_1500.in')
    t0 = time.perf_counter()
    res = run_bscan(inp, n, list(range(ngpu)))
    wall = time.perf_counter() - t0
    ttot = [res[k]['ttotal'] for k in sorted(res)]
    tsol = [res[k]['tsolve'] for k in sorted(res)]
    parity = None
    gold = 'tests/golden/bscan_gssi_trace1_f32.npz'
    if os.path.exists(gold) and 1 in res:
        from baseline.standins import read_out
        from gprmax_b200.model_io import load_model
        _, g = load_model(gold)
        out = read_out(os.path.join(work, 'cylinder_Bscan_GSSI_15001.out'))
        ref, got = g['rx0_Ey'], out['data']['/rxs/rx1/Ey']
        parity = {'trace': 1, 'component': 'Ey', 'max_rel': float(np.abs(got - ref).max() / np.abs(ref).max()), 'tolerance': 1e-4,
                  'against': gold + ' (unmodified reference CPU solver)', 'datasets_in_out_file': sorted(out['data'])}
    cells, its = 480 * 148 * 235, 3117
    print(json.dumps({'workload': 'cylinder_Bscan_GSSI_1500.in -n {} farmed over {} GPU(s), reference front end per trace'.format(n, ngpu),
                      'traces': n, 'gpus': ngpu, 'wall_s': wall, 'traces_per_hour': n / wall * 3600,
                      'per_trace_s': {'run_model total (median)': float(np.median(ttot)), 'time loop (median)': float(np.median(tsol)),
                                      'host parse + build + write (median)': float(np.median(np.array(ttot) - np.array(tsol)))},
                      'solve_mcells_per_s': cells * its / float(np.median(tsol)) / 1e6, 'parity': parity,
                      'id_build': 'gprmax_b200.yee_build (all host cores)' if os.environ.get('GPRMAX_B200_REF_BUILD') != '1' else 'reference (single thread)'}))
    shutil.rmtree(work, ignore_errors=True)
