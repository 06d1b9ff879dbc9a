#!/usr/bin/env python
"""Benchmark of the FDTD time-stepping hot path (BASELINE.json: Mcells/s, % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size S] [--iters I]

Metric (docs/source/benchmarking.rst:111-117 of the reference): Mcells/s = nx*ny*nz*iterations / (t*1e6).

N = 1   workload = tests/benchmarking/bench_300x300x300.in (BASELINE.json configs[1]): 300^3 free space,
        1 mm cells, 1559 iterations, x-directed Hertzian dipole + 1 receiver, 10-cell HORIPML on all
        faces, float32.  One "step" = one complete run of that model (all 1559 iterations).
          value : device-timed loop (CUDA events inside the library, inputs resident in HBM)
          e2e   : the drop-in call solve_gpu(1, 1, G) from HOST arrays: handle creation, H2D upload of
                  ID / coefficient / PML / waveform tables, the loop, D2H of the receiver traces.
N > 1   (torchrun, one rank per GPU) x-slab sharded synthetic lossy-dielectric domain, 256 x 2048 x 1024
        cells PER GPU (N = 8 is BASELINE.json configs[4], 2048 x 2048 x 1024), one-plane E/H halo
        exchange per half-step over NCCL; weak scaling.

--impl reference times the reference's own CPU kernels (oracle/_ref: the reference's Cython sources
compiled unmodified, OpenMP over all host cores; falls back to the plain-C port in oracle/ when
oracle/_ref is absent) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG_FP32 = 96.0  # algorithmic bytes per cell per time step, non-dispersive fp32 with uint32 IDs


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.stop = threading.Event()
        self.th = None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, universal_newlines=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(',')])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(self.rows)}


def host_threads():
    return int(os.environ.get('OMP_NUM_THREADS') or os.cpu_count() or 1)


def cpu_reference_run(G, iterations):
    """Time the reference's CPU kernels on the first `iterations` steps of G.  Returns (Mcells/s, kind, cores)."""
    from oracle.solver import have_ref, solve_cpu
    kind = 'reference' if have_ref() else 'port'
    cores = host_threads()
    # the reference pins its OpenMP threads this way (input_cmds_singleuse.py:78-80)
    os.environ.setdefault('OMP_PLACES', 'cores')
    os.environ.setdefault('OMP_PROC_BIND', 'TRUE')
    os.environ.setdefault('OMP_DYNAMIC', 'FALSE')
    out = solve_cpu(G, kernels='ref' if kind == 'reference' else 'oracle', nthreads=cores, iterations=iterations)
    mcells = G.nx * G.ny * G.nz * iterations / (out['tsolve'] * 1e6)
    return mcells, kind, cores, out['tsolve']


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path, all host threads."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    from gprmax_b200.synthetic import bench_model, homogeneous_model
    N = args.size
    sample_iters = args.iters or 12
    if args.gpus > 1:
        # the N-GPU arm runs the sharded lossy-dielectric domain (256 x 2048 x 1024 cells per GPU): the CPU solver gets a
        # bounded N^3 sample of the same recipe (material, PML, z dipole at the centre)
        c = N // 2 * 1e-3
        G = homogeneous_model((N, N, N), iterations=sample_iters, er=6.0, se=0.01, src=(c, c, c), src_pol='z', rxs=[(c + 0.02, c + 0.01, c)])
        workload = ('bounded sample of the sharded synthetic lossy-dielectric domain (er=6, sigma=0.01, z Hertzian dipole, 10-cell HORIPML): '
                    '{0}x{0}x{0} cells').format(N)
    else:
        G = bench_model(N, iterations=sample_iters)
        workload = 'tests/benchmarking/bench_{0}x{0}x{0}.in free-space cube, Hertzian dipole, 10-cell HORIPML'.format(N)
    vals, secs = [], []
    for s in range(args.warmup + args.steps):
        v, kind, cores, t = cpu_reference_run(G, sample_iters)
        if s >= args.warmup:
            vals.append(v)
            secs.append(t)
    value = float(np.mean(vals))
    line = {
        'impl': 'reference', 'metric': 'FDTD throughput', 'value': value, 'unit': 'Mcells/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': float(np.mean(secs) * 1e3), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload, 'cells': N**3, 'iterations_per_step': sample_iters},
        'cpu_baseline': {'value': value, 'unit': 'Mcells/s', 'cores': cores, 'kind': kind,
                         'sample': 'first {} iterations of a {}^3 model, {} OpenMP threads'.format(sample_iters, N, cores)},
        'e2e': {'value': value, 'unit': 'Mcells/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


def run_single_gpu(args):
    import ctypes
    from gprmax_b200 import GPU, Solver, solve_gpu
    from gprmax_b200.synthetic import bench_model

    N = args.size
    G = bench_model(N, iterations=args.iters)
    G.gpu = GPU(0)
    G.gpu.get_gpu_info()
    cells = G.nx * G.ny * G.nz
    its = G.iterations
    S = sum(p.thickness * {'x': G.ny * G.nz, 'y': G.nx * G.nz, 'z': G.nx * G.ny}[p.direction[0]] for p in G.pmls)
    b_alg = B_ALG_FP32 + 32.0 * len(G.cfs) * S / cells  # + PML Phi read+write
    peak, peak_src = measured_peaks()

    # ---- device-resident leg: handle created once, K timed full runs
    sv = Solver(G, device_id=0)
    times = []
    launches = 0
    with ClockSampler(0) as clocks:
        for s in range(args.warmup + args.steps):
            sv.reset()
            l0 = sv.kernel_launches
            sv.run()
            if s >= args.warmup:
                times.append(sv.elapsed)
                launches += sv.kernel_launches - l0
        # per-kernel device times (plain launches, CUDA events on the launching stream)
        sv.reset()
        nprof = min(its, 200)
        sv.profile(min(20, nprof))  # warm
        prof = sv.profile(nprof - min(20, nprof)) if nprof > 20 else sv.profile(0)
        nprof_timed = max(nprof - 20, 0)
    clk = clocks.summary()
    sv.close()
    t_step = float(np.mean(times))
    value = cells * its / (t_step * 1e6)

    # roofline of the dominant kernel
    dom = 'update_e' if prof['update_e'] >= prof['update_h'] else 'update_h'
    roof = None
    if nprof_timed > 0:
        t_launch = prof[dom] / nprof_timed * 1e-3
        alg_bytes = cells * b_alg / 2.0  # one half-step
        achieved = alg_bytes / t_launch / 1e9
        share = prof[dom] / max(sum(prof.values()), 1e-30)
        # DRAM bytes per launch from the committed ncu --set full capture of the same workload (profiles/README.md, r1p):
        # k_update_tma 843.4 MB read + 338.3 MB written at 300^3 (all PML slabs fused); only quoted for that exact workload
        traffic = 1.1817e9 if N == 300 else None
        roof = {'bound': 'hbm', 'kernel': 'k_update_tma<PHASE={}> ({}: base update + all six PML slabs in one launch)'.format(1 if dom == 'update_e' else 0, dom),
                'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'traffic_source': 'profiles/r1p_main_raw.csv (dram__bytes_read.sum + dram__bytes_write.sum)',
                'peak_source': peak_src, 'alg_bytes_per_launch': alg_bytes, 'launch_ms': t_launch * 1e3,
                'share_of_step': share, 'whole_step_frac': value * 1e6 * b_alg / 1e9 / peak,
                'kernel_ms_per_iteration': {k: v / nprof_timed for k, v in prof.items()}}

    # ---- end-to-end leg: the public drop-in call from host arrays
    # (1 untimed call, then `steps` timed calls, median reported, all samples kept.  Every call creates a solver from the
    # host arrays -- 654 MB of uint32 IDs go through pinned bounce buffers -- runs all iterations and copies the traces
    # back; device memory freed by the previous call is reused by the library's cache, as in a B-scan)
    e2e_t = []
    for s in range(1 + max(args.steps, 3)):
        t0 = time.perf_counter()
        tsolve, mem = solve_gpu(1, 1, G)
        dt = time.perf_counter() - t0
        if s >= 1:
            e2e_t.append(dt)
    e2e_value = cells * its / (float(np.median(e2e_t)) * 1e6)
    real = np.dtype(G.updatecoeffsE.dtype).itemsize
    h2d = G.ID.nbytes + G.updatecoeffsE.nbytes + G.updatecoeffsH.nbytes + sum(8 * p.ERA.nbytes for p in G.pmls) \
        + sum(s.waveformvalues_wholestep.nbytes for s in G.hertziandipoles) + 12 * len(G.rxs)
    d2h = 9 * its * len(G.rxs) * real

    # ---- CPU baseline beside it (bounded sample of the same workload)
    cpu = None
    if not args.no_cpu:
        sample = args.cpu_iters
        Gc = bench_model(N, iterations=sample)
        v, kind, cores, t = cpu_reference_run(Gc, sample)
        cpu = {'value': v, 'unit': 'Mcells/s', 'cores': cores, 'kind': kind,
               'sample': 'first {} of {} iterations of the same {}^3 model, {} OpenMP threads, {:.1f} s'.format(sample, its, N, cores, t)}

    line = {
        'metric': 'FDTD throughput', 'value': value, 'unit': 'Mcells/s', 'n_gpus': 1, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': t_step * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'tests/benchmarking/bench_{0}x{0}x{0}.in: {0}^3 free space, Hertzian dipole, 1 rx, 10-cell HORIPML x6, float32'.format(N),
                   'cells': cells, 'iterations_per_step': its, 'l2': 'working set {:.0f} MB >> 126 MB L2 (no flush needed)'.format(mem / 1e6),
                   'gpu': G.gpu.name, 'alg_bytes_per_cell_step': b_alg},
        'roofline': roof, 'cpu_baseline': cpu,
        'e2e': {'value': e2e_value, 'unit': 'Mcells/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'call': 'gprmax_b200.solve_gpu(1, 1, G) from host arrays', 'seconds_per_call': [round(t, 4) for t in e2e_t], 'statistic': 'median'},
        'gpu_launches': int(launches), 'clocks': clk,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--size', type=int, default=300, help='cube side of the single-GPU benchmark model')
    ap.add_argument('--iters', type=int, default=None, help='iterations per step (default: the model\'s own 1559)')
    ap.add_argument('--cpu-iters', type=int, default=40, help='iterations of the CPU baseline sample')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--workload', default='auto', choices=['auto', 'slab'], help="'slab' at N=1: run the sharded runs' per-GPU slab on one GPU")
    args = ap.parse_args()
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    if args.impl == 'reference':
        return run_reference_arm(args)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.gpus > 1 or world > 1 or args.workload == 'slab':
        # N = 1 with --workload slab: the per-GPU slab of the sharded runs on one GPU (weak-scaling baseline)
        if world == 1 and args.gpus == 1:
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            os.environ.setdefault('MASTER_PORT', '29517')
            os.environ.setdefault('RANK', '0')
            os.environ.setdefault('WORLD_SIZE', '1')
        from gprmax_b200.sharded import bench_sharded
        return bench_sharded(args)
    return run_single_gpu(args)


if __name__ == '__main__':
    sys.exit(main())
