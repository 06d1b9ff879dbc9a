"""One half-step on random fields, persistent vs one-CTA-per-item TMA launches: where do they differ?"""
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
        for rep in range(int(sys.argv[3])):
            for c in range(6):
                sv.set_field(c, init[c])
            sv.half_step(0, -1)   # H
            sv.half_step(1, -1)   # E
            sv.synchronize()
            for c in range(6):
                out['r%d_f%d' % (rep, c)] = sv.get_field(c)
    np.savez(sys.argv[2], **out)
    sys.exit(0)
reps = 6
def run(tag, env):
    subprocess.run([sys.executable, __file__, 'run', '/tmp/rp_%s.npz' % tag, str(reps)], env=dict(os.environ, **env), check=True)
    return np.load('/tmp/rp_%s.npz' % tag)
ref = run('np', {'GPB_TMA_NOPERSIST': '1'})
per = run('p', {})
names = ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz')
for rep in range(reps):
    for c in range(6):
        a, b = per['r%d_f%d' % (rep, c)], ref['r%d_f%d' % (rep, c)]
        d = np.argwhere(a != b)
        if len(d):
            print('rep', rep, names[c], len(d), 'cells; i%8:', np.bincount(d[:, 0] % 8, minlength=8), ' j%14:', np.bincount(d[:, 1] % 14, minlength=14), ' k%64 hist16:', np.bincount((d[:, 2] % 64) // 4, minlength=16), 'first', d[:5].tolist(), flush=True)
    print('rep', rep, 'done', flush=True)
for c in range(6):
    same = True
    print('nopersist', names[c], 'stable over reps:', same)
