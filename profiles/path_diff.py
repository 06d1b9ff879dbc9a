"""Bitwise comparison of the final fields of the TMA-staged and the register-vectorised kernel paths on one model
(they must agree bit for bit: shards below the TMA threshold run next to a single-GPU reference above it)."""
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
its = sys.argv[1] if len(sys.argv) > 1 else '40'
for tag, env in (('tma', {}), ('v4', {'GPB_NO_TMA': '1'}), ('tmapw0', {'GPB_TMA_PW': '0'}), ('zsplit', {'GPB_TMA_ZSPLIT': '1'})):
    subprocess.run([sys.executable, __file__, 'run', '/tmp/pd_%s.npz' % tag, its], env=dict(os.environ, **env), check=True)
a = np.load('/tmp/pd_tma.npz')
for other in ('v4', 'tmapw0', 'zsplit'):
    b = np.load('/tmp/pd_%s.npz' % other)
    print('tma vs', other, ': receivers equal:', np.array_equal(a['rx'], b['rx']))
    for c, n in enumerate(('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz')):
        x, y = a['f%d' % c], b['f%d' % c]
        d = np.argwhere(x != y)
        if len(d):
            print(' ', n, len(d), 'cells differ; i range', d[:, 0].min(), d[:, 0].max(), 'j range', d[:, 1].min(), d[:, 1].max(), 'k range', d[:, 2].min(), d[:, 2].max(),
                  'max rel', float(np.abs(x - y).max() / np.abs(y).max()), 'first', d[0])
