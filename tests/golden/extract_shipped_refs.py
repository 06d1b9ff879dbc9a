"""Extract the receiver traces of the golden outputs the reference ships with its own regression tests
(`tests/models_basic/<model>/<model>_ref.out`, written by gprMax 3.0.13 / 3.1.0b1; the reference's `tests/test_models.py`
compares a fresh run with them) into `tests/golden/shipped_ref_traces.npz`, so that the parity tests can use them where
/root/reference does not exist.  HDF5 is read with the pure-Python reader next to this file (no h5py in this image).

    python tests/golden/extract_shipped_refs.py [/root/reference]
"""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from h5v0 import H5File  # noqa: E402

OUT = os.path.join(HERE, 'shipped_ref_traces.npz')


def extract(ref_root):
    data = {}
    for path in sorted(glob.glob(os.path.join(ref_root, 'tests', 'models_basic', '*', '*_ref.out'))):
        model = os.path.basename(path)[:-len('_ref.out')]
        f = H5File(path)
        a = f.attrs('/')
        data[model + '/version'] = np.array(str(a['gprMax']))
        data[model + '/iterations'] = np.array(int(a['Iterations']))
        data[model + '/dt'] = np.array(float(a['dt']))
        for rx in f.keys('/rxs'):
            data['{}/{}/position'.format(model, rx)] = np.asarray(f.attrs('/rxs/' + rx)['Position'], dtype=np.float64)
            for comp in f.keys('/rxs/' + rx):
                data['{}/{}/{}'.format(model, rx, comp)] = f.dataset('/rxs/{}/{}'.format(rx, comp))
    return data


if __name__ == '__main__':
    root = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    d = extract(root)
    np.savez_compressed(OUT, **d)
    print('wrote', OUT, len(d), 'arrays,', os.path.getsize(OUT), 'bytes')
