"""Run the same model twice per configuration and compare the final fields bit for bit (race detector)."""
import os, sys, subprocess
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == 'run':
    from gprmax_b200 import Solver
    from sharded_worker import build
    G = build('synthetic:160,144,128,' + sys.argv[3])
    with Solver(G, device_id=0) as sv:
        if os.environ.get('PROBE_RESET'):
            sv.reset()
        if os.environ.get('PROBE_SLEEP'):
            import time; time.sleep(1.0)
        sv.run()
        np.savez(sys.argv[2], **{'f%d' % c: sv.get_field(c) for c in range(6)})
    sys.exit(0)
its = 40
def run(tag, env):
    subprocess.run([sys.executable, __file__, 'run', '/tmp/dt_%s.npz' % tag, str(its)], env=dict(os.environ, **env), check=True)
    z = np.load('/tmp/dt_%s.npz' % tag)
    return [z['f%d' % c] for c in range(6)]
ref = run('v4', {'GPB_NO_TMA': '1'})
for label, env in (('default', {}), ('reset first', {'PROBE_RESET': '1'}), ('sleep first', {'PROBE_SLEEP': '1'}), ('nopool', {'GPB_NO_POOL': '1'}), ('default', {})):
    a = run('a', env); b = run('b', env)
    same = all(np.array_equal(x, y) for x, y in zip(a, b))
    vs = all(np.array_equal(x, y) for x, y in zip(a, ref))
    nd = sum(int((x != y).sum()) for x, y in zip(a, b))
    print('%-10s run-to-run identical: %-5s (%d cells differ)   identical to v4: %s' % (label, same, nd, vs), flush=True)
