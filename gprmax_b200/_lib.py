"""ctypes binding of libgprmax_b200.so -- mirrors include/gprmax_b200.h field for field.

There is no fallback: if the shared library is missing, `lib()` raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get('GPB_LIB') or os.path.join(HERE, 'libgprmax_b200.so')

GPB_ABI_VERSION = 3
GPB_F32, GPB_F64 = 0, 1
GPB_HORIPML, GPB_MRIPML = 0, 1
GPB_SRC_HERTZIAN, GPB_SRC_MAGNETIC, GPB_SRC_VOLTAGE = 0, 1, 2
GPB_NRXOUT = 9
RX_ROWS = ['Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz', 'Ix', 'Iy', 'Iz']

# every symbol include/gprmax_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = ['gpb_device_count', 'gpb_device_info', 'gpb_create', 'gpb_destroy', 'gpb_run', 'gpb_iteration',
           'gpb_elapsed_seconds', 'gpb_mem_used', 'gpb_kernel_launches', 'gpb_reset', 'gpb_set_points', 'gpb_profile', 'gpb_kernel_path', 'gpb_half_step', 'gpb_halo',
           'gpb_stream', 'gpb_synchronize', 'gpb_create_sharded', 'gpb_link_info', 'gpb_link', 'gpb_get_receivers', 'gpb_get_snapshot', 'gpb_get_tline',
           'gpb_get_field', 'gpb_set_field', 'gpb_release_cached', 'gpb_ids_scan', 'gpb_ids_apply', 'gpb_ids_scan_slab', 'gpb_ids_apply_slab', 'gpb_vtk_transpose', 'gpb_last_error', 'gpb_version']


class DeviceInfo(C.Structure):
    _fields_ = [('device_id', C.c_int32), ('name', C.c_char * 256), ('pci_bus_id', C.c_char * 32),
                ('total_mem', C.c_uint64), ('const_mem', C.c_uint64), ('sm_count', C.c_int32),
                ('cc_major', C.c_int32), ('cc_minor', C.c_int32)]


class Pml(C.Structure):
    _fields_ = [('direction', C.c_int32),
                ('xs', C.c_int32), ('xf', C.c_int32), ('ys', C.c_int32), ('yf', C.c_int32), ('zs', C.c_int32), ('zf', C.c_int32),
                ('thickness', C.c_int32), ('d', C.c_double),
                ('ERA', C.c_void_p), ('ERB', C.c_void_p), ('ERE', C.c_void_p), ('ERF', C.c_void_p),
                ('HRA', C.c_void_p), ('HRB', C.c_void_p), ('HRE', C.c_void_p), ('HRF', C.c_void_p)]


class Source(C.Structure):
    _fields_ = [('kind', C.c_int32), ('i', C.c_int32), ('j', C.c_int32), ('k', C.c_int32), ('polarisation', C.c_int32),
                ('it_first', C.c_int32), ('it_last', C.c_int32), ('param', C.c_double), ('waveform', C.c_void_p)]


class TLine(C.Structure):
    _fields_ = [('i', C.c_int32), ('j', C.c_int32), ('k', C.c_int32), ('polarisation', C.c_int32),
                ('it_first', C.c_int32), ('it_last', C.c_int32),
                ('nl', C.c_int32), ('srcpos', C.c_int32), ('antpos', C.c_int32),
                ('resistance', C.c_double), ('dl', C.c_double), ('abcv0', C.c_double), ('abcv1', C.c_double),
                ('voltage0', C.c_void_p), ('current0', C.c_void_p), ('wave_whole', C.c_void_p), ('wave_half', C.c_void_p)]


class Snapshot(C.Structure):
    _fields_ = [('xs', C.c_int32), ('ys', C.c_int32), ('zs', C.c_int32), ('xf', C.c_int32), ('yf', C.c_int32), ('zf', C.c_int32),
                ('dx', C.c_int32), ('dy', C.c_int32), ('dz', C.c_int32),
                ('nx', C.c_int32), ('ny', C.c_int32), ('nz', C.c_int32), ('time', C.c_int32)]


class IdCombo(C.Structure):
    _fields_ = [('id', C.c_uint32 * 4), ('comp', C.c_int32), ('i', C.c_int32), ('j', C.c_int32), ('k', C.c_int32)]


class Link(C.Structure):
    _fields_ = [('process_id', C.c_uint64), ('device_id', C.c_int32), ('dtype', C.c_int32),
                ('x_start', C.c_int32), ('nx_planes', C.c_int32), ('ny', C.c_int32), ('nz', C.c_int32),
                ('plane_elems', C.c_uint64), ('array_elems', C.c_uint64), ('fields_ptr', C.c_uint64), ('flags_ptr', C.c_uint64),
                ('fields_ipc', C.c_ubyte * 64), ('flags_ipc', C.c_ubyte * 64),
                ('fields_ipc_offset', C.c_uint64), ('flags_ipc_offset', C.c_uint64)]


class Model(C.Structure):
    _fields_ = [('abi_version', C.c_int32), ('dtype', C.c_int32),
                ('nx', C.c_int32), ('ny', C.c_int32), ('nz', C.c_int32),
                ('x_start', C.c_int32), ('nx_planes', C.c_int32),
                ('dx', C.c_double), ('dy', C.c_double), ('dz', C.c_double), ('dt', C.c_double),
                ('iterations', C.c_int32), ('nmaterials', C.c_int32),
                ('ID', C.c_void_p), ('id_comp_stride', C.c_int64), ('uniform_id', C.c_int32), ('updatecoeffsE', C.c_void_p), ('updatecoeffsH', C.c_void_p),
                ('maxpoles', C.c_int32), ('updatecoeffsdispersive', C.c_void_p),
                ('pml_formulation', C.c_int32), ('pml_order', C.c_int32),
                ('npml', C.c_int32), ('pmls', C.POINTER(Pml)),
                ('nsources', C.c_int32), ('sources', C.POINTER(Source)),
                ('ntlines', C.c_int32), ('tlines', C.POINTER(TLine)),
                ('nrx', C.c_int32), ('rxcoords', C.c_void_p),
                ('nsnapshots', C.c_int32), ('snapshots', C.POINTER(Snapshot))]


_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise RuntimeError('libgprmax_b200.so is not built ({}); run `python -m gprmax_b200.build` '
                           '(needs nvcc). There is no CPU fallback.'.format(LIBPATH))
    L = C.CDLL(LIBPATH)
    H = C.c_void_p
    L.gpb_last_error.restype = C.c_char_p
    L.gpb_version.restype = C.c_char_p
    L.gpb_device_count.argtypes = [C.POINTER(C.c_int)]
    L.gpb_device_info.argtypes = [C.c_int, C.POINTER(DeviceInfo)]
    L.gpb_create.argtypes = [C.POINTER(Model), C.c_int, C.POINTER(H)]
    L.gpb_create_sharded.argtypes = [C.POINTER(Model), C.POINTER(C.c_int), C.c_int, C.POINTER(H)]
    L.gpb_link_info.argtypes = [H, C.POINTER(Link)]
    L.gpb_link.argtypes = [H, C.POINTER(Link), C.POINTER(Link)]
    P = C.c_void_p
    L.gpb_ids_scan.argtypes = [P, P, P, P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(IdCombo), C.c_int, C.POINTER(C.c_int)]
    L.gpb_ids_apply.argtypes = [P, P, P, P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(IdCombo), P, C.c_int]
    L.gpb_ids_scan_slab.argtypes = [P, P, P, P] + [C.c_int] * 9 + [C.POINTER(IdCombo), C.c_int, C.POINTER(C.c_int)]
    L.gpb_ids_apply_slab.argtypes = [P, P, P, P] + [C.c_int] * 9 + [C.POINTER(IdCombo), P, C.c_int]
    L.gpb_vtk_transpose.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), P]
    L.gpb_destroy.argtypes = [H]
    L.gpb_run.argtypes = [H, C.c_int]
    L.gpb_half_step.argtypes = [H, C.c_int, C.c_int]
    L.gpb_reset.argtypes = [H]
    L.gpb_set_points.argtypes = [H, C.POINTER(Model)]
    L.gpb_profile.argtypes = [H, C.c_int, C.POINTER(C.c_double)]
    L.gpb_kernel_path.argtypes = [H, C.c_char_p, C.c_size_t]
    L.gpb_iteration.argtypes = [H, C.POINTER(C.c_int)]
    L.gpb_elapsed_seconds.argtypes = [H, C.POINTER(C.c_double)]
    L.gpb_mem_used.argtypes = [H, C.POINTER(C.c_uint64)]
    L.gpb_kernel_launches.argtypes = [H, C.POINTER(C.c_uint64)]
    L.gpb_get_receivers.argtypes = [H, C.c_void_p, C.c_size_t]
    L.gpb_get_snapshot.argtypes = [H, C.c_int, C.POINTER(C.c_void_p), C.c_size_t]
    L.gpb_get_tline.argtypes = [H, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    L.gpb_get_field.argtypes = [H, C.c_int, C.c_void_p, C.c_size_t]
    L.gpb_set_field.argtypes = [H, C.c_int, C.c_void_p, C.c_size_t]
    L.gpb_halo.argtypes = [H, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.gpb_stream.argtypes = [H, C.POINTER(C.c_void_p)]
    L.gpb_synchronize.argtypes = [H]
    for name in SYMBOLS:
        f = getattr(L, name)
        if name not in ('gpb_last_error', 'gpb_version'):
            f.restype = C.c_int
    _lib = L
    return L


def last_error():
    return (lib().gpb_last_error() or b'').decode('utf-8', 'replace')
