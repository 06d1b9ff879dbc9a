"""The C-ABI shared library loads and exports every symbol include/gprmax_b200.h declares
(no compute calls here: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from gprmax_b200 import _lib
from gprmax_b200.build import build_library


@pytest.fixture(scope='module')
def library():
    build_library()
    return _lib.lib()


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'gprmax_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gpb_[a-z_]+)\s*\(', text)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_symbol(library):
    for name in header_symbols():
        assert hasattr(library, name), name


def test_struct_sizes_match_header(library):
    # natural C layout of the structs in the header (x86-64 / sm_100a host ABI)
    assert ctypes.sizeof(_lib.Pml) == 8 * 4 + 8 + 8 * 8
    assert ctypes.sizeof(_lib.Source) == 7 * 4 + 4 + 8 + 8
    assert ctypes.sizeof(_lib.Snapshot) == 13 * 4
    assert ctypes.sizeof(_lib.TLine) == 9 * 4 + 4 + 4 * 8 + 4 * 8
    assert _lib.Model.abi_version.offset == 0
    assert _lib.Model.dx.offset % 8 == 0


def test_no_gpu_fails_loudly(library):
    """Without a CUDA device the product path must raise, never fall back."""
    import numpy as np
    from gprmax_b200 import GeneralError, detect_check_gpus
    from gprmax_b200.gpu import device_count
    if device_count() > 0:
        pytest.skip('a GPU is present')
    with pytest.raises(GeneralError):
        detect_check_gpus([0])
    from gprmax_b200.model_io import load_model
    from gprmax_b200.solver import solve_gpu
    from conftest import golden_path
    G, _ = load_model(golden_path('cylinder_Ascan_2D'))
    with pytest.raises(GeneralError):
        solve_gpu(1, 1, G)
    assert b'version' not in library.gpb_version() and library.gpb_version().startswith(b'gprmax_b200')
