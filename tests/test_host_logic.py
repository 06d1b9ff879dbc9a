"""Host-side logic that needs no GPU: device selection semantics of detect_check_gpus (utilities.py:369-413) incl.
CUDA_VISIBLE_DEVICES, the composite GPU of a sharded run, slab partitioning, the packed model of a slab that points into the
global ID array, the h5py stand-in used when the reference writes output files on a box without h5py."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, golden_path


@pytest.fixture
def fake_devices(monkeypatch):
    from gprmax_b200 import gpu

    def fake_info(self, drv=None):
        self.name, self.pcibusID, self.constmem, self.totalmem, self.smcount = 'FAKE B200', '0000:0{}:00.0'.format(self.ordinal), 65536, 180 * 2**30, 148
    monkeypatch.setattr(gpu, 'device_count', lambda: 4)
    monkeypatch.setattr(gpu.GPU, 'get_gpu_info', fake_info)
    monkeypatch.delenv('CUDA_VISIBLE_DEVICES', raising=False)
    monkeypatch.delenv('GPRMAX_B200_SHARD', raising=False)
    monkeypatch.setattr(sys, 'argv', ['gprMax', 'model.in'])
    return gpu


def test_default_device_and_listing(fake_devices):
    gpus, text = fake_devices.detect_check_gpus([])
    assert [g.deviceID for g in gpus] == [0] and len(text) == 4 and text[0].startswith('0 - FAKE B200, 180GiB')
    with pytest.raises(fake_devices.GeneralError, match='GPU with device ID 7 does not exist'):
        fake_devices.detect_check_gpus([7])


def test_cuda_visible_devices_ids_are_physical(fake_devices, monkeypatch):
    """utilities.py:386-390: with CUDA_VISIBLE_DEVICES set the IDs a user may name are the ones listed there."""
    monkeypatch.setenv('CUDA_VISIBLE_DEVICES', '4,5,6,7')
    gpus, text = fake_devices.detect_check_gpus([6])
    assert gpus[0].deviceID == 6 and gpus[0].ordinal == 2 and text[2].startswith('6 - ')
    with pytest.raises(fake_devices.GeneralError):
        fake_devices.detect_check_gpus([0])


def test_device_list_means_one_sharded_model(fake_devices):
    gpus, _ = fake_devices.detect_check_gpus([0, 1, 3])
    head = gpus[0]
    assert head.shard_ordinals == [0, 1, 3] and head.totalmem == 3 * 180 * 2**30 and len(gpus) == 3
    assert gpus[1].shard_ordinals == [1]      # the others stay plain


def test_farm_modes_keep_the_reference_meaning(fake_devices, monkeypatch):
    monkeypatch.setattr(sys, 'argv', ['gprMax', 'model.in', '-n', '8', '-mpi', '5'])
    gpus, _ = fake_devices.detect_check_gpus([0, 1])
    assert gpus[0].shard_ordinals == [0] and gpus[0].totalmem == 180 * 2**30
    monkeypatch.setattr(sys, 'argv', ['gprMax', 'model.in'])
    monkeypatch.setenv('GPRMAX_B200_SHARD', '0')
    gpus, _ = fake_devices.detect_check_gpus([0, 1])
    assert gpus[0].shard_ordinals == [0]


def test_partition_planes_balanced_and_contiguous():
    from gprmax_b200.sharded import partition_planes
    for nx, world in ((300, 1), (300, 8), (40, 3), (7, 8)):
        parts = partition_planes(nx, world)
        assert parts[0][0] == 0 and sum(n for _, n in parts) == nx + 1
        assert all(a + n == b for (a, n), (b, _) in zip(parts, parts[1:]))
        assert max(n for _, n in parts) - min(n for _, n in parts) <= 1
    with pytest.raises(ValueError):
        partition_planes(3, 5)


def test_slab_model_points_into_the_global_id_array():
    """A slab handed the GLOBAL G.ID gets a pointer into it and the global component stride (no second host copy)."""
    from gprmax_b200.model_io import load_model
    from gprmax_b200.solver import PackedModel
    G, _ = load_model(golden_path('pml_HORIPML_1', 'f32'))
    ID = np.ascontiguousarray(G.ID, dtype=np.uint32)
    G.ID = ID
    pm = PackedModel(G, x_start=10, nx_planes=12)
    m = pm.model
    assert m.ID == ID.ctypes.data + 10 * ID.strides[1] and m.id_comp_stride == ID.strides[0] // 4
    pm = PackedModel(G)
    assert pm.model.ID == ID.ctypes.data and pm.model.id_comp_stride == 0
    pm = PackedModel(G, x_start=3, nx_planes=5, ID=np.ascontiguousarray(ID[:, 3:8]))
    assert pm.model.id_comp_stride == 0
    pm = PackedModel(G, with_id=False)
    assert not pm.model.ID


def test_h5py_standin_round_trip(tmp_path):
    sys.path.insert(0, ROOT)
    from baseline import standins
    f = standins.File(str(tmp_path / 'a.out'), 'w')
    f.attrs['Iterations'] = 7
    g = f.create_group('/rxs/rx1')
    g.attrs['Position'] = (0.1, 0.2, 0.3)
    f['/rxs/rx1/Ez'] = np.arange(7, dtype=np.float32)
    g['Hx'] = np.ones(7, dtype=np.float32)
    f.close()
    out = standins.read_out(str(tmp_path / 'a.out'))
    assert int(out['attrs']['/']['Iterations']) == 7 and out['attrs']['/rxs/rx1']['Position'].tolist() == [0.1, 0.2, 0.3]
    assert np.array_equal(out['data']['/rxs/rx1/Ez'], np.arange(7, dtype=np.float32)) and '/rxs/rx1/Hx' in out['data']


def test_committed_ncu_capture_belongs_to_these_kernel_sources():
    """bench.py quotes `roofline.traffic` from profiles/traffic.json only while the capture's hash over the device-code files equals
    the sources being built; a kernel edit without a new `ncu --set full` capture (profiles/ncu_r2.sh) turns the figure into null."""
    import sys
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    import bench
    traffic, source = bench.committed_traffic(300)
    assert traffic is not None, source
    assert 1.0e9 < traffic < 1.4e9          # 300^3 half-step: 1.18 GB moved, 1.38 GB algorithmic
