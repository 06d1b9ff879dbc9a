"""Models built by the UNMODIFIED reference front end (baseline/_ref, see baseline/install_ref.sh).

`build_with_reference(inputfile)` runs the reference exactly as `python -m gprMax inputfile -gpu` would -- input
parsing, geometry / material / PML build, `run_model` (model_build_run.py:84-404) -- and stops at the seam where the
reference calls `solve_gpu(currentmodelrun, modelend, G)` (model_build_run.py:373), returning the fully built FDTDGrid.
Nothing of this is on the measured path: bench.py and the tests use it to obtain the very object the drop-in receives
inside gprMax, on a box where only baseline/_ref (not /root/reference) exists.
"""
import contextlib
import io
import os


class _Captured(Exception):
    pass


def reference_available():
    import baseline
    return baseline.have_reference()


def reference_input(relpath):
    """Path of an input file of the vendored reference, e.g. 'tests/benchmarking/bench_300x300x300.in'."""
    import baseline
    return os.path.join(baseline.REF_DIR, relpath)


def build_with_reference(inputfile, currentmodelrun=1, n=1, device=None, quiet=True):
    """The FDTDGrid `G` the reference hands to solve_gpu for model run `currentmodelrun` of `n`.

    device: a GPU-like object (`.deviceID .name .totalmem .constmem`) to build for; default: a description of device 0
    from the library when a CUDA device is present, else a placeholder (CPU-only container: nothing is run on it)."""
    import baseline
    baseline.use_reference()
    import gprMax.gprMax as top
    import gprMax.model_build_run as mbr
    from gprmax_b200 import GPU
    from gprmax_b200.gpu import device_count

    if device is None:
        device = GPU(0)
        if device_count() > 0:
            device.get_gpu_info()
        else:
            device.name, device.pcibusID, device.constmem, device.totalmem = 'no CUDA device (build only)', '0', 65536, 180 * 2**30
    cap = {}

    def seam(cur, end, G):
        cap['G'] = G
        raise _Captured()

    saved = (top.detect_check_gpus, mbr.solve_gpu)
    top.detect_check_gpus = lambda ids: ([device], ['{} - {}'.format(device.deviceID, device.name)])
    mbr.solve_gpu = seam
    out = io.StringIO()
    try:
        with (contextlib.redirect_stdout(out) if quiet else contextlib.nullcontext()):
            try:
                top.api(inputfile, n=n, restart=currentmodelrun if currentmodelrun != 1 else None, gpu=[int(device.deviceID)])
            except _Captured:
                pass
    finally:
        top.detect_check_gpus, mbr.solve_gpu = saved
    if 'G' not in cap:
        raise RuntimeError('the reference front end did not reach solve_gpu for {}:\n{}'.format(inputfile, out.getvalue()[-2000:]))
    G = cap['G']
    from gprMax.materials import Material
    G.maxpoles = Material.maxpoles
    return G
