"""Use the B200 core from the gprMax command line without touching the gprMax tree.

    python -m gprmax_b200 model.in -gpu            one GPU
    python -m gprmax_b200 model.in -gpu 0 1 2 3    ONE model sharded as x-slabs over four GPUs
    python -m gprmax_b200 model.in -n 54 -gpu 0    B-scan, traces one after the other
    (every other gprMax option is passed through unchanged)

`install()` replaces the two call sites SURVEY.md 8(b) names -- `gprMax.gprMax.detect_check_gpus` (gprMax.py:136) and
`gprMax.model_build_run.solve_gpu` (model_build_run.py:373) -- by this package's drop-ins, exactly the two assignments
INTEGRATION.md gives a maintainer.  It also lifts the two input restrictions the reference only has because its PyCUDA
solver lacks the features: `#transmission_line` under `-gpu` (input_cmds_multiuse.py:317-319) and the Ix/Iy/Iz receiver
outputs (:417-418); both are CPU-only in the reference and are supported by this core.  The input commands are parsed by
the reference's own `process_multicmds`; it merely does not see `G.gpu` while it parses.  Finally the per-edge ID build
(`build_electric_components` / `build_magnetic_components`, model_build_run.py:216-218) runs on all host cores with an identical
result (gprmax_b200/yee_build.py; GPRMAX_B200_REF_BUILD=1 keeps the reference's single-threaded loop), and the snapshot and
geometry-view files are written by the streaming writers of gprmax_b200/vtk_writers.py (byte-identical files;
GPRMAX_B200_REF_WRITERS=1 keeps the reference's).
"""
import os
import sys


def _import_reference():
    try:
        import gprMax  # noqa: F401
    except ImportError:
        # not installed: the vendored copy of this repository (baseline/_ref, see baseline/install_ref.sh)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        if root not in sys.path:
            sys.path.insert(0, root)
        import baseline
        baseline.use_reference()
        import gprMax  # noqa: F401
    import gprMax.gprMax as top
    import gprMax.model_build_run as mbr
    return top, mbr


_installed = False


def install():
    """Point the running gprMax at this core (idempotent).  Returns (gprMax.gprMax, gprMax.model_build_run)."""
    global _installed
    top, mbr = _import_reference()
    if _installed:
        return top, mbr
    from . import detect_check_gpus, solve_gpu
    # errors leave this package as the reference's own GeneralError (exceptions.py:29-36), whatever was imported first
    from gprMax.exceptions import GeneralError as RefError
    import gprmax_b200
    from . import exceptions, gpu, solver
    for mod in (gprmax_b200, exceptions, gpu, solver):
        mod.GeneralError = RefError
    top.detect_check_gpus = detect_check_gpus
    mbr.solve_gpu = solve_gpu

    parse = mbr.process_multicmds

    def process_multicmds(multicmds, G):
        gpu, G.gpu = G.gpu, None    # transmission lines and Ix/Iy/Iz outputs are fine on this core
        try:
            return parse(multicmds, G)
        finally:
            G.gpu = gpu
    mbr.process_multicmds = process_multicmds

    # per-edge material IDs on all host cores (yee_build.py; identical G.ID and G.materials, tests/test_yee_build.py).  The
    # reference calls the electric and the magnetic build back to back (model_build_run.py:216-218): the first call does both.
    if os.environ.get('GPRMAX_B200_REF_BUILD') != '1':
        from . import yee_build
        from gprMax.yee_cell_build_ext import create_electric_average, create_magnetic_average

        def build_electric_components(solid, rigidE, ID, G):
            yee_build.build_components(G, create_electric_average, create_magnetic_average)

        def build_magnetic_components(solid, rigidH, ID, G):
            pass
        mbr.build_electric_components = build_electric_components
        mbr.build_magnetic_components = build_magnetic_components
        # PML face averages without the per-cell Python search through the material list (pml_build.py; identical tables)
        from .pml_build import build_pmls
        mbr.build_pmls = build_pmls
    # snapshot and geometry-view files streamed through a bounded buffer, byte-identical (vtk_writers.py)
    if os.environ.get('GPRMAX_B200_REF_WRITERS') != '1':
        from . import vtk_writers
        vtk_writers.install()
    _installed = True
    return top, mbr


def main(argv=None):
    top, _ = install()
    if argv is not None:
        sys.argv = [sys.argv[0]] + list(argv)
    return top.main()
