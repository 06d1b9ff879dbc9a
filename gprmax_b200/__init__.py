"""gprmax_b200 -- B200-native (sm_100a) FDTD time-stepping core for gprMax.

A drop-in for the reference's PyCUDA solver path only (model_build_run.solve_gpu and the
helpers it calls); everything above the time loop stays the reference's own Python.

    from gprmax_b200 import solve_gpu, detect_check_gpus, GPU

The compute lives in libgprmax_b200.so (hand-written CUDA, C ABI in include/gprmax_b200.h,
bound with ctypes).  There is no CPU fallback.
"""
from .exceptions import GeneralError
from .gpu import GPU, detect_check_gpus
from .solver import Solver, solve_gpu
from .model_io import SolverGrid, load_model, save_model

__all__ = ['GeneralError', 'GPU', 'detect_check_gpus', 'Solver', 'solve_gpu', 'SolverGrid', 'load_model', 'save_model']
__version__ = '0.1.0'
