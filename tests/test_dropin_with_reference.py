"""Drop-in seam against the REAL reference (only where /root/reference exists, i.e. the build container).

Runs the reference's own front end (`gprMax.run(file, gpu=[0])` -> run_main -> run_model) with the two patched
symbols INTEGRATION.md describes.  There is no GPU here, so the test stops at the boundary: it checks that the
reference's fully built FDTDGrid *in GPU mode* (no host field arrays, no PML Phi arrays, `G.gpu` set) is accepted by
the packer that feeds the C ABI, that the packed model equals the fixture the CPU run produced, and that without a
device the product path raises (no fallback).
"""
import os
import subprocess
import sys

import pytest

from conftest import ROOT, have_ref_kernels

REF = os.environ.get('GPRMAX_REFERENCE', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.isdir(REF) or not have_ref_kernels('f32'), reason='reference tree / oracle/_ref not available')

SCRIPT = r'''
import os, sys, tempfile, shutil
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests', 'golden'))
import numpy as np
from ref_import import import_reference
gprMax = import_reference('f32')
import gprMax.gprMax as top
import gprMax.model_build_run as mbr
import gprmax_b200
from gprmax_b200.solver import PackedModel
from gprmax_b200.model_io import load_model

class FakeGPU(gprmax_b200.GPU):
    def get_gpu_info(self, drv=None):
        self.name, self.pcibusID, self.constmem, self.totalmem = 'no device in this container', '0', 65536, 180 * 2**30

def fake_detect(ids):
    g = FakeGPU(0); g.get_gpu_info(); return [g], ['0 - fake']

seen = {{}}
def boundary(currentmodelrun, modelend, G):
    assert G.gpu is not None and not hasattr(G, 'Ex')          # GPU mode: model_build_run.py:183-184
    pm = PackedModel(G)
    m = pm.model
    seen.update(nx=m.nx, ny=m.ny, nz=m.nz, its=m.iterations, npml=m.npml, nsrc=m.nsources, nrx=m.nrx, nmat=m.nmaterials,
                order=m.pml_order, idsum=int(np.asarray(G.ID, dtype=np.uint64).sum()))
    # without a device the real drop-in must raise, not fall back
    try:
        gprmax_b200.solve_gpu(currentmodelrun, modelend, G)
    except gprmax_b200.GeneralError as e:
        seen['error'] = str(e)
    return 0.0, 0

top.detect_check_gpus = fake_detect            # gprMax.py:136
mbr.solve_gpu = boundary                       # model_build_run.py:373
mbr.write_hdf5_outputfile = lambda f, G: None
work = tempfile.mkdtemp()
shutil.copy(os.path.join({ref!r}, 'user_models', 'cylinder_Ascan_2D.in'), work)
gprMax.run(os.path.join(work, 'cylinder_Ascan_2D.in'), gpu=[0])
G, _ = load_model(os.path.join({root!r}, 'tests', 'golden', 'cylinder_Ascan_2D_f32.npz'))
assert (seen['nx'], seen['ny'], seen['nz'], seen['its']) == (G.nx, G.ny, G.nz, G.iterations), seen
assert seen['npml'] == len(G.pmls) and seen['nrx'] == 1 and seen['nsrc'] == 1 and seen['nmat'] == G.updatecoeffsE.shape[0]
assert seen['idsum'] == int(np.asarray(G.ID, dtype=np.uint64).sum())
assert 'error' in seen and ('CUDA' in seen['error'] or 'device' in seen['error'] or 'GPU' in seen['error']), seen
print('DROPIN_OK', seen['error'])
'''


def test_reference_front_end_reaches_the_boundary():
    r = subprocess.run([sys.executable, '-c', SCRIPT.format(root=ROOT, ref=REF)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       universal_newlines=True, timeout=600, cwd='/tmp')
    assert r.returncode == 0 and 'DROPIN_OK' in r.stdout, r.stdout[-3000:]
