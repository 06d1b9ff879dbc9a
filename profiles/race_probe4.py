import os, sys, subprocess
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
NS = (17, 18, 19, 20, 24, 40)
if len(sys.argv) > 1 and sys.argv[1] == 'run':
    from gprmax_b200 import Solver
    from sharded_worker import build
    G = build('synthetic:160,144,128,40')
    out = {}
    with Solver(G, device_id=0) as sv:
        for n in NS:
            sv.reset()
            sv.run(n)
            for c in range(6):
                out['n%d_f%d' % (n, c)] = sv.get_field(c)
    np.savez(sys.argv[2], **out)
    sys.exit(0)
def run(tag, env):
    subprocess.run([sys.executable, __file__, 'run', '/tmp/rs_%s.npz' % tag], env=dict(os.environ, **env), check=True)
    return np.load('/tmp/rs_%s.npz' % tag)
ref = run('np', {'GPB_TMA_NOPERSIST': '1'})
per = run('p', {})
names = ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz')
shown = 0
for n in NS:
    for c in range(6):
        a, b = per['n%d_f%d' % (n, c)], ref['n%d_f%d' % (n, c)]
        d = np.argwhere(a != b)
        if len(d) and shown < 14:
            shown += 1
            print('run(%d)' % n, names[c], len(d), 'cells differ:', d[:6].tolist(), 'persist', ['%.6e' % a[tuple(x)] for x in d[:6]], 'ref', ['%.6e' % b[tuple(x)] for x in d[:6]], 'absmax %.3e' % np.abs(b).max(), flush=True)
    print('run(%d) compared' % n, flush=True)
