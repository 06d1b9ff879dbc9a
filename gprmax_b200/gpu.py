"""GPU discovery -- drop-in for `GPU` / `detect_check_gpus` of the reference
(gprMax/utilities.py:341-413), backed by libgprmax_b200.so instead of PyCUDA.
"""
import ctypes as C

from . import _lib
from .exceptions import GeneralError


def human_size(size, a_kilobyte_is_1024_bytes=True):
    """Same rendering as the reference's utilities.human_size (utilities.py:122-146)."""
    suffixes = {1000: ['KB', 'MB', 'GB', 'TB', 'PB', 'EB', 'ZB', 'YB'],
                1024: ['KiB', 'MiB', 'GiB', 'TiB', 'PiB', 'EiB', 'ZiB', 'YiB']}
    if size < 0:
        raise ValueError('Number must be non-negative.')
    multiple = 1024 if a_kilobyte_is_1024_bytes else 1000
    for suffix in suffixes[multiple]:
        size /= multiple
        if size < multiple:
            return '{:.3g}{}'.format(size, suffix)
    raise ValueError('Number is too large.')


class GPU(object):
    """GPU information (utilities.py:341-366).  Same attributes; `get_gpu_info` needs no driver object."""

    def __init__(self, deviceID):
        self.deviceID = deviceID
        self.name = None
        self.pcibusID = None
        self.constmem = None
        self.totalmem = None
        self.smcount = None
        # device IDs of an x-slab sharded run (single-GPU: just this one); see sharded.py
        self.shard_deviceIDs = [deviceID]

    def get_gpu_info(self, drv=None):
        L = _lib.lib()
        info = _lib.DeviceInfo()
        if L.gpb_device_info(int(self.deviceID), C.byref(info)):
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


def detect_check_gpus(deviceIDs):
    """Get information about Nvidia GPU(s) (utilities.py:369-413).

    Args:
        deviceIDs (list): List of integers of device IDs.

    Returns:
        gpus (list): Detected GPU(s) object(s).
        allgpustext (list): one line of text per visible device.

    Device IDs are CUDA runtime ordinals (0..count-1 after any CUDA_VISIBLE_DEVICES remapping).
    """
    count = device_count()
    if count == 0:
        raise GeneralError('No NVIDIA CUDA-Enabled GPUs detected (https://developer.nvidia.com/cuda-gpus)')
    deviceIDsavail = range(count)

    # If no device ID is given use default of 0
    if not deviceIDs:
        deviceIDs = [0]

    for ID in deviceIDs:
        if ID not in deviceIDsavail:
            raise GeneralError('GPU with device ID {} does not exist'.format(ID))

    gpus = []
    allgpustext = []
    for ID in deviceIDsavail:
        gpu = GPU(deviceID=ID)
        gpu.get_gpu_info()
        if ID in deviceIDs:
            gpus.append(gpu)
        allgpustext.append('{} - {}, {}'.format(gpu.deviceID, gpu.name, human_size(gpu.totalmem, a_kilobyte_is_1024_bytes=True)))

    return gpus, allgpustext
