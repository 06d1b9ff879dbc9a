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
import time


def _worker_models(wid, device_id, build, tasks, results):
    from .gpu import GPU
    from .solver import solve_gpu
    gpu = GPU(device_id)
    gpu.get_gpu_info()
    while True:
        k = tasks.get()
        if k is None:           # sentinel, cf. StopIteration in gprMax.py:404-406
            break
        t0 = time.perf_counter()
        G = build(k)
        tbuild = time.perf_counter() - t0
        G.gpu = gpu
        tsolve, mem = solve_gpu(k, k, G)
        out = {n: {name: v for name, v in rx.outputs.items()} for n, rx in enumerate(G.rxs)}
        results.put((k, device_id, tbuild, tsolve, out))


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
    out = {}
    for _ in range(n):
        k, dev, tb, ts, rxs = results.get()
        out[k] = {'device': dev, 'tbuild': tb, 'tsolve': ts, 'rxs': rxs}
    for p in procs:
        p.join()
    return out


def _worker_gprmax(wid, device_id, inputfile, n, tasks, results, extra_args):
    """One gprMax worker process bound to one GPU (cf. run_mpi_sim worker, gprMax.py:436-471)."""
    import argparse
    from . import detect_check_gpus, solve_gpu
    import gprMax.gprMax as top
    import gprMax.model_build_run as mbr
    mbr.solve_gpu = solve_gpu                      # the drop-in
    gpus, _ = detect_check_gpus([device_id])
    args = argparse.Namespace(inputfile=inputfile, n=n, task=None, restart=None, mpi=False, mpi_no_spawn=False, mpicomm=None,
                              gpu=gpus[0], benchmark=False, geometry_only=False, geometry_fixed=False, write_processed=False,
                              opt_taguchi=False)
    for k, v in (extra_args or {}).items():
        setattr(args, k, v)
    from gprMax.constants import c, e0, m0, z0
    while True:
        k = tasks.get()
        if k is None:
            break
        with open(inputfile) as f:
            usernamespace = {'c': c, 'e0': e0, 'm0': m0, 'z0': z0, 'number_model_runs': n, 'inputfile': os.path.abspath(inputfile)}
            tsolve = mbr.run_model(args, k, n, n, f, usernamespace)
        results.put((k, device_id, tsolve))


def run_bscan(inputfile, n, gpus, extra_args=None):
    """`python -m gprMax inputfile -n N -gpu ...` farmed one trace per GPU; needs the reference importable.
    Output files (`<name><k>.out`) are written by the reference's own writer, ready for
    tools/outputfiles_merge.py."""
    ctx = mp.get_context('spawn')
    tasks, results = ctx.Queue(), ctx.Queue()
    for k in range(1, n + 1):
        tasks.put(k)
    for _ in gpus:
        tasks.put(None)
    procs = [ctx.Process(target=_worker_gprmax, args=(w, dev, inputfile, n, tasks, results, extra_args)) for w, dev in enumerate(gpus)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(n):
        k, dev, ts = results.get()
        out[k] = {'device': dev, 'tsolve': ts}
    for p in procs:
        p.join()
    return out
