"""B-scan farm: independent model runs (traces) distributed one model per GPU.

Mirrors the reference's MPI task farm (gprMax/gprMax.py:327-588: master builds a work list of
`{'currentmodelrun': k}` items, workers pull until a StopIteration sentinel, worker w uses
`args.gpu[w % len(args.gpu)]`) without mpi4py: a local pool of worker processes, one per GPU, each
pulling model numbers from a shared queue.  Traces are replicas - no device-to-device traffic.

Two ways to use it:

  * inside gprMax (the reference's Python is importable): `run_bscan(inputfile, n, gpus)` makes every
    worker call the reference's own `run_model(...)` for its trace (host parse + geometry build stay
    the reference's), with `solve_gpu` / `detect_check_gpus` swapped for this package's drop-ins
    (INTEGRATION.md);
  * without gprMax: `run_models(build, n, gpus)` where `build(k)` returns a solver-ready grid
    (`SolverGrid`, e.g. loaded with `model_io.load_model` and with the source/receiver moved for
    trace k); each worker solves it with `solve_gpu` and returns the receiver outputs.
"""
import multiprocessing as mp
import os
import queue
import time
import traceback


def _worker_models(wid, device_id, build, tasks, results):
    try:
        from .gpu import GPU
        from .solver import solve_gpu
        gpu = GPU(device_id)
        gpu.get_gpu_info()
    except Exception as e:      # no device / no library: every task this worker would take is reported as failed
        results.put(('fatal', wid, device_id, '{}: {}'.format(type(e).__name__, e)))
        return
    while True:
        k = tasks.get()
        if k is None:           # sentinel, cf. StopIteration in gprMax.py:404-406
            break
        try:
            t0 = time.perf_counter()
            G = build(k)
            tbuild = time.perf_counter() - t0
            G.gpu = gpu
            tsolve, mem = solve_gpu(k, k, G)
            out = {n: {name: v for name, v in rx.outputs.items()} for n, rx in enumerate(G.rxs)}
            results.put(('ok', k, device_id, tbuild, tsolve, out))
        except Exception as e:
            results.put(('error', k, device_id, '{}: {}\n{}'.format(type(e).__name__, e, traceback.format_exc())))


def _collect(procs, results, n, what):
    """Gather n results; a worker that reports an error or dies (CUDA error, killed) raises here instead of leaving the parent
    blocked on the queue for ever."""
    out = {}
    while len(out) < n:
        try:
            rec = results.get(timeout=1.0)
        except queue.Empty:
            dead = [p for p in procs if not p.is_alive() and p.exitcode not in (0, None)]
            if dead:
                for p in procs:
                    p.terminate()
                raise RuntimeError('{}: worker process {} died with exit code {}'.format(what, dead[0].pid, dead[0].exitcode))
            if all(not p.is_alive() for p in procs) and results.empty():
                raise RuntimeError('{}: all workers finished but only {} of {} results arrived'.format(what, len(out), n))
            continue
        if rec[0] != 'ok':
            for p in procs:
                p.terminate()
            raise RuntimeError('{}: {} on device {} failed: {}'.format(what, 'worker start-up' if rec[0] == 'fatal' else 'model {}'.format(rec[1]), rec[2], rec[3]))
        out[rec[1]] = rec[2:]
    for p in procs:
        p.join()
    return out


def run_models(build, n, gpus, start=1):
    """Solve models start..start+n-1 (`build(k)` -> grid) on the given device ids, one worker per GPU.
    Returns {k: {'device': id, 'tbuild': s, 'tsolve': s, 'rxs': {rx index: {component: trace}}}}."""
    ctx = mp.get_context('spawn')
    tasks, results = ctx.Queue(), ctx.Queue()
    for k in range(start, start + n):
        tasks.put(k)
    for _ in gpus:
        tasks.put(None)
    procs = [ctx.Process(target=_worker_models, args=(w, dev, build, tasks, results)) for w, dev in enumerate(gpus)]
    for p in procs:
        p.start()
    got = _collect(procs, results, n, 'run_models')
    return {k: {'device': dev, 'tbuild': tb, 'tsolve': ts, 'rxs': rxs} for k, (dev, tb, ts, rxs) in got.items()}


def _worker_gprmax(wid, device_id, inputfile, n, tasks, results, extra_args):
    """One gprMax worker process bound to one GPU (cf. run_mpi_sim worker, gprMax.py:436-471): the reference's own
    run_model per trace -- input parsing, geometry / material / PML build, output file -- with the time loop on this core."""
    try:
        import argparse
        from . import detect_check_gpus
        from .dropin import install
        top, mbr = install()                           # the two assignments of INTEGRATION.md
        gpus, _ = detect_check_gpus([device_id])
        args = argparse.Namespace(inputfile=inputfile, n=n, task=None, restart=None, mpi=False, mpi_no_spawn=False, mpicomm=None,
                                  gpu=gpus[0], benchmark=False, geometry_only=False, geometry_fixed=False, write_processed=False,
                                  opt_taguchi=False)
        for k, v in (extra_args or {}).items():
            setattr(args, k, v)
        from gprMax.constants import c, e0, m0, z0
    except Exception as e:
        results.put(('fatal', wid, device_id, '{}: {}\n{}'.format(type(e).__name__, e, traceback.format_exc())))
        return
    while True:
        k = tasks.get()
        if k is None:
            break
        try:
            t0 = time.perf_counter()
            with open(inputfile) as f:
                usernamespace = {'c': c, 'e0': e0, 'm0': m0, 'z0': z0, 'number_model_runs': n, 'inputfile': os.path.abspath(inputfile)}
                tsolve = mbr.run_model(args, k, n, n, f, usernamespace)
            results.put(('ok', k, device_id, time.perf_counter() - t0, tsolve))
        except Exception as e:
            results.put(('error', k, device_id, '{}: {}\n{}'.format(type(e).__name__, e, traceback.format_exc())))


def run_bscan(inputfile, n, gpus, extra_args=None):
    """`python -m gprMax inputfile -n N -gpu ...` farmed one trace per GPU; needs the reference importable (installed, or the
    vendored baseline/_ref).  Output files (`<name><k>.out`) are written by the reference's own writer, ready for
    tools/outputfiles_merge.py.  Returns {k: {'device': id, 'ttotal': host seconds of run_model (parse + build + solve +
    write), 'tsolve': seconds of the time loop}}."""
    ctx = mp.get_context('spawn')
    tasks, results = ctx.Queue(), ctx.Queue()
    for k in range(1, n + 1):
        tasks.put(k)
    for _ in gpus:
        tasks.put(None)
    procs = [ctx.Process(target=_worker_gprmax, args=(w, dev, inputfile, n, tasks, results, extra_args)) for w, dev in enumerate(gpus)]
    for p in procs:
        p.start()
    got = _collect(procs, results, n, 'run_bscan')
    return {k: {'device': dev, 'ttotal': tt, 'tsolve': ts} for k, (dev, tt, ts) in got.items()}
