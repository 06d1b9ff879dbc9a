"""The reference's analytical known-answer test (`tests/models_basic/hertzian_dipole_fs_analytical` compared with
`tests/analytical_solutions.py:hertzian_dipole_fs` by its `tests/test_models.py`): evaluate the reference's own analytic
solution of a Hertzian dipole in free space for the source / receiver geometry of the `hertzian_dipole_fs` fixture (the two
.in files are identical) and store it as `tests/golden/hertzian_dipole_fs_analytic.npz`.

    python tests/golden/make_analytic.py          (needs the reference: baseline/_ref or /root/reference)
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main():
    import baseline
    ref_dir = baseline.use_reference()            # makes `gprMax` importable (the analytic module imports its waveforms)
    path = os.path.join(ref_dir, 'tests', 'analytical_solutions.py')
    if not os.path.exists(path):
        path = os.path.join(os.environ.get('GPRMAX_REFERENCE', '/root/reference'), 'tests', 'analytical_solutions.py')
    spec = importlib.util.spec_from_file_location('ref_analytical_solutions', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from gprmax_b200.model_io import load_model
    G, _ = load_model(os.path.join(HERE, 'hertzian_dipole_fs_f32.npz'))
    src, rx = G.hertziandipoles[0], G.rxs[0]
    rel = ((rx.xcoord - src.xcoord) * G.dx, (rx.ycoord - src.ycoord) * G.dy, (rx.zcoord - src.zcoord) * G.dz)
    fields = mod.hertzian_dipole_fs(G.iterations, G.dt, (G.dx, G.dy, G.dz), rel)
    out = os.path.join(HERE, 'hertzian_dipole_fs_analytic.npz')
    np.savez_compressed(out, fields=fields, rel=np.array(rel), dt=np.array(G.dt), iterations=np.array(G.iterations))
    print('wrote', out, fields.shape)


if __name__ == '__main__':
    main()
