"""Host side of config 4 (cylinder_Bscan_GSSI_1500.in, 16.7 M cells, with its geometry view switched on) through the gprMax command
line of this package, `--geometry-only`: parse + geometry build + ID build + PML build + `.vti` geometry view, once with the host
rows of SURVEY.md 8(f) installed (multi-threaded ID build, vectorised PML build, streaming writer) and once with the reference's
own (GPRMAX_B200_REF_BUILD=1 GPRMAX_B200_REF_WRITERS=1).  CPU only.   python profiles/host_path_bench.py"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')


def main():
    out = {}
    digests = {}
    with tempfile.TemporaryDirectory() as tmp:
        for mode in ('reference_rows', 'b200_rows'):
            d = os.path.join(tmp, mode)
            os.mkdir(d)
            text = open(os.path.join(REF, 'user_models', 'cylinder_Bscan_GSSI_1500.in')).read().replace('\ngeometry_view:', '\n#geometry_view:')
            open(os.path.join(d, 'model.in'), 'w').write(text)
            for f in ('GSSI.py', 'MALA.py'):
                shutil.copy(os.path.join(REF, 'user_libs', 'antennas', f), d)
            env = dict(os.environ, PYTHONPATH=ROOT)
            if mode == 'reference_rows':
                env.update(GPRMAX_B200_REF_BUILD='1', GPRMAX_B200_REF_WRITERS='1')
            best = None
            for rep in range(3):
                t0 = time.perf_counter()
                r = subprocess.run([sys.executable, '-m', 'gprmax_b200', 'model.in', '--geometry-only'], cwd=d, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
                t = time.perf_counter() - t0
                assert r.returncode == 0, r.stdout[-2000:]
                best = t if best is None else min(best, t)
            out[mode] = {'wall_s': round(best, 2)}
            import hashlib
            digests[mode] = hashlib.sha256(open(os.path.join(d, 'cylinder_GSSI_1500.vti'), 'rb').read()).hexdigest()
    out['geometry_view_identical'] = digests['reference_rows'] == digests['b200_rows']
    out['speedup'] = round(out['reference_rows']['wall_s'] / out['b200_rows']['wall_s'], 2)
    out['host_cores'] = len(os.sched_getaffinity(0))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
