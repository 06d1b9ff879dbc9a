import os, sys, subprocess
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == 'run':
    from gprmax_b200 import Solver
    from sharded_worker import build
    G = build('synthetic:160,144,128,40')
    out = {}
    with Solver(G, device_id=0) as sv:
        for n in range(1, 21):
            sv.run(1)
            for c in range(6):
                out['n%d_f%d' % (n, c)] = sv.get_field(c)
    np.savez(sys.argv[2], **out)
    sys.exit(0)
def run(tag, env):
    subprocess.run([sys.executable, __file__, 'run', '/tmp/rr_%s.npz' % tag], env=dict(os.environ, **env), check=True)
    return np.load('/tmp/rr_%s.npz' % tag)
ref = run('np', {'GPB_TMA_NOPERSIST': '1'})
per = run('p', {})
names = ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz')
shown = 0
for n in range(1, 21):
    for c in range(6):
        a, b = per['n%d_f%d' % (n, c)], ref['n%d_f%d' % (n, c)]
        d = np.argwhere(a != b)
        if len(d) and shown < 12:
            shown += 1
            i, j, k = d[0]
            print('it', n, names[c], len(d), 'cells differ; first', d[:4].tolist(), 'persist', [float(a[tuple(x)]) for x in d[:4]], 'ref', [float(b[tuple(x)]) for x in d[:4]],
                  'field absmax', float(np.abs(b).max()), 'smallest nonzero |ref|', float(np.abs(b[b != 0]).min()) if (b != 0).any() else 0, flush=True)
print('done')
