"""Where the end-to-end time of solve_gpu() goes for the 300^3 benchmark model (host wall clock), with the calling thread
free or pinned to one core the way the OpenMP runtime pins it inside gprMax (OMP_PROC_BIND=TRUE, input_cmds_singleuse.py:78-80).

    python profiles/e2e_breakdown.py [pin]        (GPB_TIMING=1 prints the library's set-up phases)"""
import os, sys, time
sys.path.insert(0, ".")
if 'pin' in sys.argv:
    os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[0]})   # before CUDA starts, like libgomp does
import numpy as np
from gprmax_b200 import GPU, Solver, solve_gpu
from gprmax_b200.solver import PackedModel, store_results
from benchkit.synthetic import bench_model
G = bench_model(300)
G.gpu = GPU(0); G.gpu.get_gpu_info()
print('affinity of the calling thread: {} core(s)'.format(len(os.sched_getaffinity(0))))
for rep in range(3):
    t0 = time.perf_counter(); sv = Solver(G, device_id=0); t1 = time.perf_counter()
    sv.run(); t2 = time.perf_counter()
    store_results(G, sv); t3 = time.perf_counter()
    el = sv.elapsed
    sv.close(); t4 = time.perf_counter()
    print('create %.1f ms | run %.1f ms (device %.1f) | results %.1f ms | destroy %.1f ms | total %.1f ms' % ((t1-t0)*1e3, (t2-t1)*1e3, el*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t4-t0)*1e3))
for rep in range(3):
    t0 = time.perf_counter(); ts, mem = solve_gpu(1, 1, G); print('solve_gpu %.1f ms (tsolve %.1f ms)' % ((time.perf_counter()-t0)*1e3, ts*1e3))
