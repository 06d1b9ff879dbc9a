import os, sys, subprocess
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == 'run':
    from gprmax_b200 import Solver
    from sharded_worker import build
    G = build('synthetic:160,144,128,' + sys.argv[3])
    with Solver(G, device_id=0) as sv:
        sv.run()
        np.savez(sys.argv[2], rx=sv.receivers(), **{'f%d' % c: sv.get_field(c) for c in range(6)})
    sys.exit(0)
def run(tag, env, its):
    subprocess.run([sys.executable, __file__, 'run', '/tmp/pd_%s.npz' % tag, str(its)], env=dict(os.environ, **env), check=True)
    return np.load('/tmp/pd_%s.npz' % tag)
def cmp(a, b, label):
    out = []
    for c, n in enumerate(('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz')):
        x, y = a['f%d' % c], b['f%d' % c]
        d = np.argwhere(x != y)
        if len(d):
            out.append('%s %d cells i[%d,%d] j[%d,%d] k[%d,%d] rel %.1e' % (n, len(d), d[:, 0].min(), d[:, 0].max(), d[:, 1].min(), d[:, 1].max(), d[:, 2].min(), d[:, 2].max(), float(np.abs(x - y).max() / max(np.abs(y).max(), 1e-30))))
    print(label, 'IDENTICAL' if not out else ' | '.join(out), flush=True)
for its in (5, 10, 20, 40):
    t1 = run('t1', {}, its); t2 = run('t2', {}, its); v = run('v4', {'GPB_NO_TMA': '1'}, its); sc = run('sc', {'GPB_SCALAR': '1'}, its); p0 = run('p0', {'GPB_TMA_PW': '0'}, its)
    print('--- iterations', its)
    cmp(t1, t2, 'tma vs tma again :')
    cmp(t1, v, 'tma vs v4        :')
    cmp(t1, p0, 'tma vs tma pw0   :')
    cmp(v, sc, 'v4  vs scalar    :')
