"""GPU discovery -- drop-in for `GPU` / `detect_check_gpus` of the reference
(gprMax/utilities.py:341-413), backed by libgprmax_b200.so instead of PyCUDA.

Two things go beyond the reference, both behind the same call:

  * device IDs follow the reference's convention under CUDA_VISIBLE_DEVICES (utilities.py:386-388: the IDs a user may name
    are the physical IDs listed there); each `GPU` additionally carries `.ordinal`, the CUDA runtime ordinal the library
    needs (the position in that list);
  * `-gpu 0 1 2 3` outside the reference's farm modes (-mpi, --mpi-no-spawn, -benchmark) means ONE model sharded as
    x-slabs over those devices.  The reference keeps only `gpus[0]` there (gprMax.py:141-144) and checks the model against
    `gpus[0].totalmem` (grid.py:239-241), so `gpus[0]` is returned as a composite: `.totalmem` is the sum over the listed
    devices and `.shard_deviceIDs` / `.shard_ordinals` remember all of them for `solve_gpu`.
"""
import ctypes as C
import os
import sys

from . import _lib
from .exceptions import GeneralError


def human_size(size, a_kilobyte_is_1024_bytes=True):
    """Byte count as text, e.g. '179GiB'.  Inside gprMax the reference's own formatter (utilities.py:122-146) is used, so the
    text is the reference's by construction; stand-alone, an equivalent three-significant-digit rendering."""
    if 'gprMax.utilities' in sys.modules:
        return sys.modules['gprMax.utilities'].human_size(size, a_kilobyte_is_1024_bytes=a_kilobyte_is_1024_bytes)
    if size < 0:
        raise ValueError('Number must be non-negative.')
    base, units = (1024, ('KiB', 'MiB', 'GiB', 'TiB', 'PiB', 'EiB')) if a_kilobyte_is_1024_bytes else (1000, ('KB', 'MB', 'GB', 'TB', 'PB', 'EB'))
    value = float(size)
    for unit in units:
        value /= base
        if value < base:
            return '{:.3g}{}'.format(value, unit)
    raise ValueError('Number is too large.')


class GPU(object):
    """GPU information (utilities.py:341-366).  Same attributes; `get_gpu_info` needs no driver object."""

    def __init__(self, deviceID, ordinal=None):
        self.deviceID = deviceID
        self.ordinal = deviceID if ordinal is None else ordinal   # CUDA runtime ordinal (differs under CUDA_VISIBLE_DEVICES)
        self.name = None
        self.pcibusID = None
        self.constmem = None
        self.totalmem = None
        self.smcount = None
        # devices of an x-slab sharded run (single GPU: just this one); read by solve_gpu
        self.shard_deviceIDs = [deviceID]
        self.shard_ordinals = [self.ordinal]

    def get_gpu_info(self, drv=None):
        L = _lib.lib()
        info = _lib.DeviceInfo()
        if L.gpb_device_info(int(self.ordinal), C.byref(info)):
            raise GeneralError(_lib.last_error())
        self.name = info.name.decode('utf-8', 'replace')
        self.pcibusID = info.pci_bus_id.decode('utf-8', 'replace')
        self.constmem = int(info.const_mem)
        self.totalmem = int(info.total_mem)
        self.smcount = int(info.sm_count)


def device_count():
    L = _lib.lib()
    n = C.c_int(0)
    if L.gpb_device_count(C.byref(n)):
        return 0
    return n.value


def _farm_mode():
    """True when the reference will hand the GPUs out one per model (gprMax.py:141-142: -mpi / --mpi-no-spawn / -benchmark,
    and the workers those modes spawn)."""
    flags = ('-mpi', '--mpi-no-spawn', '-benchmark', '--mpi-worker')
    return any(a in flags for a in sys.argv[1:]) or os.environ.get('GPRMAX_B200_SHARD', '1') == '0'


def detect_check_gpus(deviceIDs):
    """Get information about Nvidia GPU(s) (utilities.py:369-413).

    Args:
        deviceIDs (list): List of integers of device IDs.

    Returns:
        gpus (list): Detected GPU(s) object(s).
        allgpustext (list): one line of text per visible device.
    """
    count = device_count()
    if count == 0:
        raise GeneralError('No NVIDIA CUDA-Enabled GPUs detected (https://developer.nvidia.com/cuda-gpus)')
    # utilities.py:386-390: with CUDA_VISIBLE_DEVICES set, the IDs on offer are the ones listed there
    available = list(range(count))
    visible = os.environ.get('CUDA_VISIBLE_DEVICES')
    if visible is not None:
        try:
            listed = [int(s) for s in visible.split(',') if s.strip() != '']
            if len(listed) == count:
                available = listed
        except ValueError:      # GPU-<uuid> entries: fall back to runtime ordinals
            pass

    if not deviceIDs:           # utilities.py:393-395
        deviceIDs = [0] if 0 in available else [available[0]]
    for ID in deviceIDs:        # utilities.py:398-401
        if ID not in available:
            raise GeneralError('GPU with device ID {} does not exist'.format(ID))

    gpus, allgpustext = [], []
    for ordinal, ID in enumerate(available):
        gpu = GPU(deviceID=ID, ordinal=ordinal)
        gpu.get_gpu_info()
        if ID in deviceIDs:
            gpus.append(gpu)
        allgpustext.append('{} - {}, {}'.format(gpu.deviceID, gpu.name, human_size(gpu.totalmem, a_kilobyte_is_1024_bytes=True)))

    if len(gpus) > 1 and not _farm_mode():
        # one model over all listed devices: gpus[0] stands for the set (see the module docstring)
        head = gpus[0]
        head.shard_deviceIDs = [g.deviceID for g in gpus]
        head.shard_ordinals = [g.ordinal for g in gpus]
        head.totalmem = sum(g.totalmem for g in gpus)
        head.name = '{} x {} (x-slab sharded: {})'.format(len(gpus), head.name, ', '.join(str(g.deviceID) for g in gpus))
    return gpus, allgpustext
