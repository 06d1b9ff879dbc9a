"""B-scan farm probe (BASELINE.json configs[3]): N traces of the GSSI 1.5 GHz model farmed one model per GPU.
The fixture holds trace 1 as built by the reference (the reference's geometry build cannot run on the GPU box),
so every trace here solves the same built model: this measures the farm + solver side (traces/hour), not the
host-side geometry build that the reference does per trace (9.4 s)."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np

FIX = os.path.join('tests', 'golden', 'bscan_gssi_trace1_f32.npz')


def build(k):
    from gprmax_b200.model_io import load_model
    G, _ = load_model(FIX)
    return G


if __name__ == '__main__':
    from gprmax_b200.farm import run_models
    from gprmax_b200.gpu import device_count
    from gprmax_b200.model_io import load_model
    ngpu = device_count()
    ntr = int(sys.argv[1]) if len(sys.argv) > 1 else 2 * ngpu
    t0 = time.perf_counter()
    out = run_models(build, ntr, list(range(ngpu)))
    wall = time.perf_counter() - t0
    _, golden = load_model(FIX)
    worst = max(float(np.abs(out[k]['rxs'][0]['Ey'] - golden['rx0_Ey']).max() / np.abs(golden['rx0_Ey']).max()) for k in out)
    ts = [out[k]['tsolve'] for k in out]
    print('farm: {} traces on {} GPU(s) in {:.1f} s wall -> {:.0f} traces/hour; solve {:.2f} s/trace (mean), load {:.2f} s/trace; '
          'worst |dEy|/peak vs reference golden {:.2e}; devices used {}'.format(
              ntr, ngpu, wall, ntr / wall * 3600, np.mean(ts), np.mean([out[k]['tbuild'] for k in out]), worst, sorted(set(out[k]['device'] for k in out))))
