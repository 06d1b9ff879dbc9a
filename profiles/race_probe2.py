"""Random initial fields, then run(n) (single full-range launches, graph for n > 1): persistent vs one-CTA-per-item."""
import os, sys, subprocess
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == 'run':
    from gprmax_b200 import Solver
    from sharded_worker import build
    G = build('synthetic:160,144,128,10')
    rng = np.random.default_rng(1)
    init = [rng.standard_normal((G.nx + 1, G.ny + 1, G.nz + 1)).astype(np.float32) for c in range(6)]
    out = {}
    with Solver(G, device_id=0) as sv:
        for n in (1, 2, 4):
            sv.reset()
            for c in range(6):
                sv.set_field(c, init[c])
            sv.run(n)
            for c in range(6):
                out['n%d_f%d' % (n, c)] = sv.get_field(c)
    np.savez(sys.argv[2], **out)
    sys.exit(0)
def run(tag, env):
    subprocess.run([sys.executable, __file__, 'run', '/tmp/rq_%s.npz' % tag], env=dict(os.environ, **env), check=True)
    return np.load('/tmp/rq_%s.npz' % tag)
ref = run('np', {'GPB_TMA_NOPERSIST': '1'})
names = ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz')
for label, env in (('persist', {}), ('persist again', {}), ('persist nograph', {'GPB_NO_GRAPH': '1'})):
    per = run('p', env)
    for n in (1, 2, 4):
        for c in range(6):
            a, b = per['n%d_f%d' % (n, c)], ref['n%d_f%d' % (n, c)]
            d = np.argwhere(a != b)
            if len(d):
                print(label, 'run(%d)' % n, names[c], len(d), 'cells; i%8:', np.bincount(d[:, 0] % 8, minlength=8).tolist(), ' j%14:', np.bincount(d[:, 1] % 14, minlength=14).tolist(),
                      ' (k%64)//4:', np.bincount((d[:, 2] % 64) // 4, minlength=16).tolist(), 'i range', d[:, 0].min(), d[:, 0].max(), 'first', d[:3].tolist(), flush=True)
        print(label, 'run(%d) compared' % n, flush=True)
