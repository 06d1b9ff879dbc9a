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

PUBLISHED_300_F32 = 5777.0  # Mcells/s the reference publishes for bench_300x300x300.in (Tesla V100, PyCUDA solver)
B_ALG_FP32 = 96.0  # algorithmic bytes per cell per time step, non-dispersive fp32 with uint32 IDs


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


from benchkit.clocks import ClockSampler  # noqa: E402


def _usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# taken at import, before any OpenMP runtime is loaded: with OMP_PROC_BIND=TRUE (which the reference sets for every model,
# input_cmds_singleuse.py:78-80) libgomp binds the main thread to one core when it starts, and the affinity mask read after
# that says "1 core"
_CORES_AT_START = _usable_cores()


def host_threads():
    """Threads for the CPU legs: the cores this process may run on.  torchrun exports OMP_NUM_THREADS=1 to its workers; that
    is a launcher default, not a property of the host, so it is ignored (GPB_CPU_THREADS overrides)."""
    if os.environ.get('GPB_CPU_THREADS'):
        return int(os.environ['GPB_CPU_THREADS'])
    return _CORES_AT_START


def cpu_reference_run(G, iterations):
    """Time the reference's CPU kernels on the first `iterations` steps of G.  Returns (Mcells/s, kind, cores)."""
    from oracle.solver import have_ref, solve_cpu
    kind = 'reference' if have_ref() else 'port'
    cores = host_threads()
    out = solve_cpu(G, kernels='ref' if kind == 'reference' else 'oracle', nthreads=cores, iterations=iterations)
    mcells = G.nx * G.ny * G.nz * iterations / (out['tsolve'] * 1e6)
    return mcells, kind, cores, out['tsolve'], out


def pin_openmp():
    """OpenMP environment of the CPU legs, set before the OpenMP runtime loads: all host threads (see host_threads), pinned
    the way the reference pins them (input_cmds_singleuse.py:78-80)."""
    os.environ['OMP_NUM_THREADS'] = str(host_threads())
    os.environ.setdefault('OMP_PLACES', 'cores')
    os.environ.setdefault('OMP_PROC_BIND', 'TRUE')
    os.environ.setdefault('OMP_DYNAMIC', 'FALSE')


SLAB = (256, 2048, 1024)   # cells per GPU of the sharded weak-scaling workload (N = 8: BASELINE.json configs[4])
SLAB_SAMPLE = (64, 1024, 512)   # bounded CPU sample of the same recipe


def slab_model(n, iterations, x_range=None, build_id=False):
    """The sharded workload's recipe on an n = (nx, ny, nz) box: one lossy dielectric (er 6, sigma 0.01 S/m) filling the domain,
    z-directed Hertzian dipole at the centre, 10-cell HORIPML."""
    from benchkit.synthetic import homogeneous_model
    nx, ny, nz = n
    cx, cy, cz = nx // 2, ny // 2, nz // 2
    return homogeneous_model(n, iterations=iterations, er=6.0, se=0.01, src=(cx * 1e-3, cy * 1e-3, cz * 1e-3), src_pol='z',
                             rxs=[(cx * 1e-3, (cy + 20) * 1e-3, cz * 1e-3), ((cx + 7) * 1e-3, (cy + 10) * 1e-3, cz * 1e-3)],
                             x_range=x_range, build_id=build_id)


def bench_grid(N, iterations=None, dtype='f32'):
    """tests/benchmarking/bench_NxNxN.in as the drop-in receives it.  Built by the unmodified reference front end when
    baseline/_ref is present (parse + geometry / material / PML build, stopped at the solve_gpu seam), otherwise by the
    closed-form builder benchkit/synthetic.py (pinned bit-exactly against the reference's build in tests/test_synthetic.py)."""
    from benchkit import refmodel
    if dtype == 'f64':   # the vendored reference is the float32 build (constants.py:36-49 is a source-level switch)
        from benchkit.synthetic import bench_model
        return bench_model(N, real=np.float64, iterations=iterations), 'benchkit.synthetic (closed-form tables), float64'
    if iterations is None and os.environ.get('GPB_BENCH_MODEL', 'reference') == 'reference' and refmodel.reference_available():
        try:
            G = refmodel.build_with_reference(refmodel.reference_input('tests/benchmarking/bench_{0}x{0}x{0}.in'.format(N)))
            G.progressbars = False   # stdout carries exactly one JSON line
            return G, 'reference front end (baseline/_ref, gprMax v3.1.7 unmodified) stopped at the solve_gpu seam'
        except Exception as e:   # e.g. the input file of that size does not exist
            sys.stderr.write('reference front end unavailable for this model ({}); using benchkit.synthetic\n'.format(e))
    from benchkit.synthetic import bench_model
    return bench_model(N, iterations=iterations), 'benchkit.synthetic (closed-form tables, bit-identical to the reference build)'


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path (its Cython/OpenMP kernels compiled unmodified
    into oracle/_ref, driven in the reference's order of operations), all host threads, on a bounded sample of the workload
    the GPU arm runs at this N."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    pin_openmp()
    from benchkit.synthetic import bench_model
    sample_iters = args.iters or 12
    if args.gpus > 1:
        n = SLAB_SAMPLE
        G = slab_model(n, sample_iters, build_id=True)
        cells = n[0] * n[1] * n[2]
        workload = ('bounded sample of the sharded workload ({0} x {1} x {2} cells per GPU x {3} GPUs, lossy dielectric er=6 sigma=0.01, z Hertzian '
                    'dipole, 10-cell HORIPML): the same recipe on a {4} x {5} x {6} box').format(SLAB[0], SLAB[1], SLAB[2], args.gpus, *n)
        sample = 'first {} iterations of the {} x {} x {} sample'.format(sample_iters, *n)
    else:
        N = args.size
        G = bench_model(N, iterations=sample_iters)
        cells = N**3
        workload = 'tests/benchmarking/bench_{0}x{0}x{0}.in: {0}^3 free space, Hertzian dipole, 1 rx, 10-cell HORIPML x6, float32'.format(N)
        sample = 'first {} of 1559 iterations of the same {}^3 model'.format(sample_iters, N)
    vals, secs = [], []
    for s in range(args.warmup + args.steps):
        v, kind, cores, t, _ = cpu_reference_run(G, sample_iters)
        if s >= args.warmup:
            vals.append(v)
            secs.append(t)
    value = float(np.mean(vals))
    line = {
        'impl': 'reference', 'metric': 'FDTD throughput', 'value': value, 'unit': 'Mcells/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': float(np.mean(secs) * 1e3), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload, 'cells': cells, 'iterations_per_step': sample_iters},
        'cpu_baseline': {'value': value, 'unit': 'Mcells/s', 'cores': cores, 'kind': kind,
                         'sample': '{}, {} OpenMP threads (OMP_NUM_THREADS of the launcher ignored: {} usable cores)'.format(sample, cores, cores)},
        'e2e': {'value': value, 'unit': 'Mcells/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


def kernel_source_hash():
    """sha256 (first 16 hex digits) over the CUDA sources: ties a committed ncu traffic figure to the kernels it was taken from."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, 'gprmax_b200', 'csrc')
    # the device code: kernel headers, the item body and the TMA instantiation unit (host-only changes in gpb_core.cu do not
    # alter the kernels a capture was taken from)
    for name in sorted(os.listdir(d)):
        if not (name.endswith('.cuh') or name.endswith('.inc') or name in ('gpb_tma_inst.cu', 'gpb_tma.h')):
            continue
        with open(os.path.join(d, name), 'rb') as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def committed_traffic(N):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture, if that capture was taken from
    the kernels being run (profiles/traffic.json is written by profiles/ncu_traffic.py together with the capture)."""
    p = os.path.join(ROOT, 'profiles', 'traffic.json')
    if not os.path.exists(p):
        return None, 'no capture committed'
    with open(p) as f:
        t = json.load(f)
    key = 'bench_{}'.format(N)
    if key not in t:
        return None, 'no capture of this workload'
    if t[key].get('kernel_source_hash') != kernel_source_hash():
        return None, 'capture {} is from other kernel sources (hash {} != {})'.format(t[key].get('file'), t[key].get('kernel_source_hash'), kernel_source_hash())
    return float(t[key]['dram_bytes_per_launch']), t[key].get('file')


def trace_errors(rx, G, ref_outputs, iters):
    """{component: max |gpu - ref| / scale} over the first `iters` samples; scale = the component's own peak in `ref`, or 10 %
    of the strongest component of the same kind (E, H) at that receiver when that is larger (tests/parity.py: a component that
    is zero by symmetry only carries the rounding noise of the big ones)."""
    from gprmax_b200 import _lib
    out = {}
    for n, r in enumerate(G.rxs):
        for name in r.outputs:
            key = 'rx{}_{}'.format(n, name)
            if key not in ref_outputs:
                continue
            ref = np.asarray(ref_outputs[key], dtype=np.float64)[:iters]
            mine = np.asarray(rx[_lib.RX_ROWS.index(name), :iters, n] if isinstance(rx, np.ndarray) else rx[key], dtype=np.float64)[:iters]
            peaks = [float(np.abs(np.asarray(ref_outputs[k], dtype=np.float64)[:iters]).max()) for k in ref_outputs
                     if k.startswith('rx{}_'.format(n)) and k.split('_')[1][0] == name[0]]
            scale = max(float(np.abs(ref).max()), 0.1 * max(peaks))
            if scale > 0:
                out[key] = float(np.abs(mine - ref).max()) / scale
    return out


def trace_parity(rx, G, ref_outputs, iters, what):
    errs = trace_errors(rx, G, ref_outputs, iters)
    worst = max(errs, key=errs.get) if errs else None
    return {'max_rel': errs[worst] if worst else 0.0, 'component': worst, 'iters': int(iters), 'against': what,
            'per_component': {k: float('{:.3e}'.format(v)) for k, v in errs.items()}}


def run_single_gpu(args):
    from gprmax_b200 import GPU, Solver, solve_gpu

    N = args.size
    G, model_source = bench_grid(N, iterations=args.iters, dtype=args.dtype)
    f64 = args.dtype == 'f64'
    rbytes = 8 if f64 else 4
    G.gpu = GPU(0)
    G.gpu.get_gpu_info()
    cells = G.nx * G.ny * G.nz
    its = G.iterations
    S = sum(p.thickness * {'x': G.ny * G.nz, 'y': G.nx * G.nz, 'z': G.nx * G.ny}[p.direction[0]] for p in G.pmls)
    b_pml = 8.0 * rbytes * len(G.cfs) * S / cells
    b_alg = 18.0 * rbytes + 24.0 + b_pml  # fields + uint32 IDs + PML Phi read+write (96 B fp32, 168 B fp64)
    peak, peak_src = measured_peaks()

    # ---- device-resident leg: handle created once, K timed full runs
    sv = Solver(G, device_id=0)
    kpath = sv.kernel_path
    idbytes = 1 if G.updatecoeffsE.shape[0] <= 256 else (2 if G.updatecoeffsE.shape[0] <= 65536 else 4)
    b_moved = 18.0 * rbytes + 6.0 * idbytes + b_pml   # what the kernels have to move with the narrowed device IDs
    times = []
    launches = 0
    rx_gpu = None
    with ClockSampler(0) as clocks:
        for s in range(args.warmup + args.steps):
            sv.reset()
            l0 = sv.kernel_launches
            sv.run()
            if s >= args.warmup:
                times.append(sv.elapsed)
                launches += sv.kernel_launches - l0
        rx_gpu = sv.receivers()   # the trace of the last timed run (parity below)
        # per-kernel device times (plain launches, CUDA events on the launching stream)
        sv.reset()
        nprof = min(its, 200)
        sv.profile(min(20, nprof))  # warm
        prof = sv.profile(nprof - min(20, nprof)) if nprof > 20 else sv.profile(0)
        nprof_timed = max(nprof - 20, 0)
    clk = clocks.summary()
    mem = sv.mem_used
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
        traffic, traffic_src = committed_traffic(N) if not f64 else (None, 'no capture of the float64 kernels')
        kname = [k for k in kpath.split() if k.startswith('E:' if dom == 'update_e' else 'H:')][0][2:]
        roof = {'bound': 'hbm', 'kernel': '{} ({}: base update + PML slabs of one half-step)'.format(kname, dom),
                'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'definition': 'ALGORITHMIC bytes of SURVEY.md 8(d) (uint32 IDs: {:.1f} B per cell per step) / launch time; the device narrows IDs to '
                              '{} byte(s), so the bytes that must actually move are {:.1f} B per cell per step -> achieved_moved / frac_moved'.format(b_alg, idbytes, b_moved),
                'achieved_moved': cells * b_moved / 2.0 / t_launch / 1e9, 'frac_moved': cells * b_moved / 2.0 / t_launch / 1e9 / peak,
                'traffic': traffic, 'traffic_source': traffic_src,
                'peak_source': peak_src, 'alg_bytes_per_launch': alg_bytes, 'launch_ms': t_launch * 1e3,
                'share_of_step': share, 'whole_step_frac': value * 1e6 * b_alg / 1e9 / peak,
                'kernel_ms_per_iteration': {k: v / nprof_timed for k, v in prof.items()},
                'note': 'kernel_ms_per_iteration is measured with plain launches and an event pair around every kernel; the graph-replayed '
                        'iteration of the timed runs (ms_per_step / iterations) is a few % shorter than their sum'}

    # ---- end-to-end leg: the public drop-in call from host arrays
    # (1 untimed call, then `steps` timed calls, median reported, all samples kept.  Every call creates a solver from the
    # host arrays -- 654 MB of uint32 IDs go through pinned bounce buffers -- runs all iterations and copies the traces
    # back; device memory freed by the previous call is reused by the library's cache, as in a B-scan)
    e2e_t = []
    for s in range(1 + max(args.steps, 3)):
        t0 = time.perf_counter()
        tsolve, _ = solve_gpu(1, 1, G)
        dt = time.perf_counter() - t0
        if s >= 1:
            e2e_t.append(dt)
    e2e_value = cells * its / (float(np.median(e2e_t)) * 1e6)
    real = np.dtype(G.updatecoeffsE.dtype).itemsize
    h2d = G.ID.nbytes + G.updatecoeffsE.nbytes + G.updatecoeffsH.nbytes + sum(8 * p.ERA.nbytes for p in G.pmls) \
        + sum(s.waveformvalues_wholestep.nbytes for s in G.hertziandipoles) + 12 * len(G.rxs)
    d2h = 9 * its * len(G.rxs) * real
    rx_e2e = {'rx{}_{}'.format(n, k): np.asarray(v) for n, r in enumerate(G.rxs) for k, v in r.outputs.items()}

    # ---- CPU baseline beside it (bounded sample of the same workload) + parity of the GPU trace against it
    cpu, parity = None, {}
    if not args.no_cpu:
        sample = min(args.cpu_iters, its)
        v, kind, cores, t, ref_out = cpu_reference_run(G, sample)
        cpu = {'value': v, 'unit': 'Mcells/s', 'cores': cores, 'kind': kind,
               'sample': 'first {} of {} iterations of the same {}^3 model, {} OpenMP threads, {:.1f} s'.format(sample, its, N, cores, t)}
        parity['prefix'] = trace_parity(rx_gpu, G, ref_out, sample, 'the reference CPU kernels ({}) run here on the same grid, first {} iterations'.format(kind, sample))
    # the complete trace against the committed golden of the unmodified reference (tests/golden/make_golden.py: bench_300_trace)
    gold = os.path.join(ROOT, 'tests', 'golden', 'bench_{}_trace_{}.npz'.format(N, args.dtype))
    if os.path.exists(gold) and args.iters is None:
        z = np.load(gold)
        g32 = {k[len('golden_'):]: z[k] for k in z.files}
        parity['full'] = trace_parity(rx_gpu, G, g32, its, 'tests/golden/bench_{}_trace_{}.npz: all {} iterations, unmodified reference CPU solver'.format(N, args.dtype, its))
        t64 = gold.replace('_f32.npz', '_f64.npz')
        if os.path.exists(t64) and not f64:
            # criterion (b) of tests/parity.py, per component: against the reference's own float64 run, the GPU's float32 trace is
            # as close as the reference's float32 trace is (x3: two realisations of the same rounding noise)
            z64 = np.load(t64)
            g64 = {k[len('golden_'):]: z64[k] for k in z64.files}
            e_cuda = trace_errors(rx_gpu, G, g64, its)
            e_ref = trace_errors(g32, G, g64, its)
            parity['full'].update(gpu32_vs_ref64={k: float('{:.3e}'.format(v)) for k, v in e_cuda.items()},
                                  ref32_vs_ref64={k: float('{:.3e}'.format(v)) for k, v in e_ref.items()})
        # the end-to-end call must give the same bits as the device-resident run
        parity['e2e_equals_resident'] = bool(all(np.array_equal(rx_e2e[k], rx_gpu[_row(k), :, int(k[2:k.index('_')])]) for k in rx_e2e))
    tol = 1e-10 if f64 else 1e-4   # north_star: float32 1e-4 of trace peak, float64 1e-10
    verdicts = {}
    for k, v in parity.get('prefix', {}).get('per_component', {}).items():
        verdicts['prefix:' + k] = 'a' if v <= tol else 'FAIL'
    for k, v in parity.get('full', {}).get('per_component', {}).items():
        b_ok = 'gpu32_vs_ref64' in parity['full'] and parity['full']['gpu32_vs_ref64'].get(k, 1e9) <= 3 * parity['full']['ref32_vs_ref64'].get(k, 0.0) + 1e-5
        verdicts['full:' + k] = 'a' if v <= tol else ('b' if b_ok else 'FAIL')
    parity['per_component_criterion'] = verdicts
    parity['criterion'] = ('float64: max|gpu64 - ref64| <= 1e-10 of trace peak (north_star)' if f64 else 'a = max|gpu32 - ref32| <= 1e-4 of trace peak (north_star); b = |gpu32 - ref64| <= 3 |ref32 - ref64| + 1e-5: as close to the '
                           "reference's float64 result as the reference's own float32 result is (tests/parity.py)")
    checks = list(verdicts.values())
    ok_a = ok_b = False
    parity['tolerance'] = tol
    parity['ok'] = ('FAIL' not in checks) if checks else None

    # ---- weak-scaling baseline: the per-GPU slab of the sharded runs on this one GPU (so that v_N / (N v_1) is like for like)
    wsb = None
    if not args.no_slab and args.iters is None and not f64:
        try:
            Gs = slab_model(SLAB, 20 * 6)
            with Solver(Gs, device_id=0) as ss:
                ts = []
                for s in range(6):
                    e0 = ss.elapsed
                    ss.run(20)
                    if s >= 3:
                        ts.append(ss.elapsed - e0)
                wsb = {'value': SLAB[0] * SLAB[1] * SLAB[2] * 20 / (float(np.mean(ts)) * 1e6), 'unit': 'Mcells/s',
                       'workload': '{} x {} x {} cells on one GPU: the per-GPU slab of the N > 1 runs, same recipe, 3 x 20 timed iterations'.format(*SLAB),
                       'kernels': ss.kernel_path}
        except Exception as e:
            wsb = {'value': None, 'error': str(e)}

    line = {
        'metric': 'FDTD throughput', 'value': value, 'unit': 'Mcells/s', 'n_gpus': 1, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': t_step * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        # the reference's published figure for this metric on this model (300^3, float32) -- other hardware
        'vs_baseline': (value / PUBLISHED_300_F32) if (N == 300 and not f64) else None, 'dtype': args.dtype, 'data': 'synthetic',
        'config': {'workload': 'tests/benchmarking/bench_{0}x{0}x{0}.in: {0}^3 free space, Hertzian dipole, 1 rx, 10-cell HORIPML x6, {1}'.format(N, 'float64' if f64 else 'float32'),
                   'cells': cells, 'iterations_per_step': its, 'l2': 'working set {:.0f} MB >> 126 MB L2 (no flush needed)'.format(mem / 1e6),
                   'gpu': G.gpu.name, 'alg_bytes_per_cell_step': b_alg, 'model_built_by': model_source, 'kernels': kpath,
                   'vs_baseline_is': '{} Mcells/s, 300^3 float32 on an NVIDIA Tesla V100 with the reference PyCUDA solver (BASELINE.md; tests/benchmarking/results/gpu/NVIDIA.png)'.format(PUBLISHED_300_F32)},
        'roofline': roof, 'cpu_baseline': cpu, 'parity': parity, 'weak_scaling_baseline': wsb,
        'e2e': {'value': e2e_value, 'unit': 'Mcells/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'call': 'gprmax_b200.solve_gpu(1, 1, G) from host arrays', 'seconds_per_call': [round(t, 4) for t in e2e_t], 'statistic': 'median'},
        'gpu_launches': int(launches), 'clocks': clk,
    }
    print(json.dumps(line))
    if parity['ok'] is False:
        sys.stderr.write('PARITY FAILURE: {}\n'.format(json.dumps(parity)))
        return 1
    return 0


def _row(key):
    from gprmax_b200 import _lib
    return _lib.RX_ROWS.index(key.split('_')[1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--size', type=int, default=300, help='cube side of the single-GPU benchmark model')
    ap.add_argument('--iters', type=int, default=None, help='iterations per step (default: the model\'s own 1559)')
    ap.add_argument('--cpu-iters', type=int, default=300, help='iterations of the CPU baseline sample (about 10 s on 32 threads at 300^3)')
    ap.add_argument('--dtype', default='f32', choices=['f32', 'f64'], help='N = 1 only: floating type of the run (the reference default and the headline is f32)')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-slab', action='store_true', help='skip the weak-scaling baseline (the 256 x 2048 x 1024 slab on one GPU)')
    ap.add_argument('--workload', default='auto', choices=['auto', 'slab'], help="'slab' at N=1: run the sharded runs' per-GPU slab on one GPU")
    args = ap.parse_args()
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    if args.impl == 'reference':
        return run_reference_arm(args)
    pin_openmp()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.gpus > 1 or world > 1 or args.workload == 'slab':
        # N = 1 with --workload slab: the per-GPU slab of the sharded runs on one GPU (weak-scaling baseline)
        if world == 1 and args.gpus == 1:
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            os.environ.setdefault('MASTER_PORT', '29517')
            os.environ.setdefault('RANK', '0')
            os.environ.setdefault('WORLD_SIZE', '1')
        from benchkit.sharded_bench import bench_sharded
        return bench_sharded(args)
    return run_single_gpu(args)


if __name__ == '__main__':
    sys.exit(main())
