"""Worker for tests/test_gpu_linked.py (fresh process: CUDA_DEVICE_MAX_CONNECTIONS must be set before CUDA starts).

    python linked_worker.py <fixture> <nslabs> <devices: 'same' | 'all'> [tma] [tl=<plane>] [its=<n>]

Runs the model on one device and as `nslabs` LINKED x-slabs inside this process (gpb_create_sharded: halo planes pushed over
peer memory, flag-synchronised, one CUDA graph per slab and iteration) and compares receivers, transmission-line totals,
snapshots and all six final field arrays bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    fixture, nslabs, devmode = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    opts = sys.argv[4:]
    if 'tma' in opts:
        os.environ['GPB_FORCE_TMA'] = '1'
    from gprmax_b200 import Solver
    from gprmax_b200.gpu import device_count
    from sharded_worker import build
    G = build(fixture)
    for o in opts:
        if o.startswith('tl='):
            for t in G.transmissionlines:
                t.xcoord = int(o[3:])
        if o.startswith('its='):
            G.iterations = min(G.iterations, int(o[4:]))
            for lst in (G.hertziandipoles, G.magneticdipoles, G.voltagesources, G.transmissionlines):
                for s in lst:
                    s.waveformvalues_wholestep = s.waveformvalues_wholestep[:G.iterations]
                    s.waveformvalues_halfstep = s.waveformvalues_halfstep[:G.iterations]
            for sn in G.snapshots:
                sn.time = min(sn.time, G.iterations - 1)

    def outputs(sv):
        sv.run()
        return dict(rx=sv.receivers(), fields=[sv.get_field(c) for c in range(6)],
                    snaps=[sv.snapshot(n) for n in range(len(G.snapshots))], tls=[sv.tline(n) for n in range(len(G.transmissionlines))])

    with Solver(G, device_id=0) as sv:
        ref = outputs(sv)
    assert np.abs(ref['rx']).max() > 0
    if devmode == 'same':
        os.environ['GPB_ALLOW_SAME_DEVICE'] = '1'
        devices = [0] * nslabs
    else:
        n = device_count()
        devices = [d % n for d in range(nslabs)]
        if nslabs > n:
            os.environ['GPB_ALLOW_SAME_DEVICE'] = '1'
    with Solver(G, devices=devices) as sv:
        path = sv.kernel_path
        out = outputs(sv)
        # a second run after reset must give the same bits (flags and line state re-armed)
        sv.reset()
        out2 = outputs(sv)
    for o in (out, out2):
        assert np.array_equal(o['rx'], ref['rx']), 'receivers differ'
        for c in range(6):
            assert np.array_equal(o['fields'][c], ref['fields'][c]), 'field {} differs'.format(c)
        for a, b in zip(o['snaps'], ref['snaps']):
            for c in range(6):
                assert np.array_equal(a[c], b[c]), 'snapshot component {} differs'.format(c)
        for a, b in zip(o['tls'], ref['tls']):
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), 'transmission line differs'
    if G.snapshots:
        assert max(float(np.abs(c).max()) for sn in ref['snaps'] for c in sn) > 0
    print('LINKED_OK', path)


if __name__ == '__main__':
    main()
