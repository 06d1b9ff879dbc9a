"""B-scan farm (gprmax_b200/farm.py): independent traces, one model per GPU, pulled from a shared queue -- the reference's MPI
task farm (gprMax.py:397-471) without mpi4py.  Also the device-resident solver across the traces of a fixed geometry
(the reference's --geometry-fixed, model_build_run.py:109, 288-330) and the drop-in command line."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, golden_path

pytestmark = pytest.mark.gpu


def test_run_models_matches_direct_solves():
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from farm_build import build
    from gprmax_b200 import GPU, solve_gpu
    from gprmax_b200.farm import run_models
    from gprmax_b200.gpu import device_count
    gpus = list(range(min(device_count(), 4)))
    n = 5
    out = run_models(build, n, gpus * (2 if len(gpus) == 1 else 1))   # two workers on a single-GPU box
    assert sorted(out) == list(range(1, n + 1))
    for k in range(1, n + 1):
        G = build(k)
        G.gpu = GPU(0)
        solve_gpu(k, n, G)
        for r, rx in enumerate(G.rxs):
            for name, v in rx.outputs.items():
                assert np.array_equal(out[k]['rxs'][r][name], v), (k, r, name)
    assert len({out[k]['device'] for k in out}) == len(gpus)


def _bad_build(k):
    raise ValueError('no such trace {}'.format(k))


def test_worker_failure_is_reported_not_hung():
    """a worker that raises (or dies) must surface in the parent instead of blocking it on the result queue."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import test_gpu_farm
    from gprmax_b200.farm import run_models
    with pytest.raises(RuntimeError, match='no such trace'):
        run_models(test_gpu_farm._bad_build, 2, [0])


def test_resident_solver_across_fixed_geometry_traces():
    """The same grid object solved again with stepped sources / receivers (what --geometry-fixed does) reuses the solver that
    is resident on the device (gpb_set_points): results identical to fresh solvers, ID array uploaded once."""
    from gprmax_b200 import GPU, solve_gpu, solver as solver_mod
    from gprmax_b200.model_io import load_model
    G, _ = load_model(golden_path('sources_mixed', 'f32'))
    G.gpu = GPU(0)
    G.srcsteps = [1, 0, 0]
    G.rxsteps = [1, 0, 0]
    created = []
    orig = solver_mod.Solver.__init__

    def counting(self, *a, **k):
        created.append(1)
        return orig(self, *a, **k)
    solver_mod.Solver.__init__ = counting
    try:
        resident = []
        for k in range(1, 4):
            if k > 1:   # model_build_run.py:303-330: sources and receivers step, nothing else changes
                for s in G.hertziandipoles + G.magneticdipoles:
                    s.xcoord += 1
                for rx in G.rxs:
                    rx.xcoord += 1
            solve_gpu(k, 3, G)
            resident.append({(r, n): v.copy() for r, rx in enumerate(G.rxs) for n, v in rx.outputs.items()})
        assert len(created) == 1 and not solver_mod._RESIDENT    # one solver for the three traces, released after the last
    finally:
        solver_mod.Solver.__init__ = orig
    # the same three traces with a fresh solver each
    G2, _ = load_model(golden_path('sources_mixed', 'f32'))
    G2.gpu = GPU(0)
    for k in range(1, 4):
        if k > 1:
            for s in G2.hertziandipoles + G2.magneticdipoles:
                s.xcoord += 1
            for rx in G2.rxs:
                rx.xcoord += 1
        solve_gpu(1, 1, G2)
        for (r, n), v in resident[k - 1].items():
            assert np.array_equal(G2.rxs[r].outputs[n], v), (k, r, n)


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'gprMax')), reason='baseline/_ref not installed')
def test_dropin_command_line_with_the_reference_front_end(tmp_path):
    """`python -m gprmax_b200 cylinder_Ascan_2D.in -gpu` = the reference's own CLI, parser, geometry build and .out writer with
    the time loop on this core; the receiver traces in the output file match the reference CPU solver's (golden)."""
    import shutil
    from gprmax_b200.model_io import load_model
    src = os.path.join(ROOT, 'baseline', '_ref', 'user_models', 'cylinder_Ascan_2D.in')
    shutil.copy(src, tmp_path)
    r = subprocess.run([sys.executable, '-m', 'gprmax_b200', str(tmp_path / 'cylinder_Ascan_2D.in'), '-gpu'], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       universal_newlines=True, timeout=600, cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT))
    assert r.returncode == 0, r.stdout[-3000:]
    assert 'GPU solving using' in r.stdout
    sys.path.insert(0, ROOT)
    from baseline.standins import read_out
    out = read_out(str(tmp_path / 'cylinder_Ascan_2D.out'))
    _, golden = load_model(golden_path('cylinder_Ascan_2D', 'f32'))
    peak = max(np.abs(golden['rx0_' + c]).max() for c in ('Ex', 'Ey', 'Ez'))
    for c in ('Ez', 'Hx', 'Hy'):
        got, ref = out['data']['/rxs/rx1/' + c], golden['rx0_' + c]
        scale = np.abs(ref).max()
        assert scale > 0 and np.abs(got - ref).max() <= 1e-4 * scale, c


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'gprMax')), reason='baseline/_ref not installed')
def test_command_line_snapshots_and_geometry_views_on_the_gpu(tmp_path):
    """`python -m gprmax_b200 model.in -gpu` with `#snapshot` and `#geometry_view` commands: the device snapshots reach the
    streaming writer as component arrays (solver.store_results -> vtk_writers.write_vtk_imagedata).  Against the same command
    on the reference's CPU solver with the reference's own writers: geometry files byte-identical, snapshot files identical in
    layout and within float32 tolerance in the field values."""
    from test_vtk_writers import MODEL
    files = {}
    for mode in ('gpu', 'cpu'):
        d = tmp_path / mode
        d.mkdir()
        (d / 'model.in').write_text(MODEL)
        env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS='4')
        if mode == 'cpu':
            env['GPRMAX_B200_REF_WRITERS'] = '1'
        r = subprocess.run([sys.executable, '-m', 'gprmax_b200', str(d / 'model.in')] + (['-gpu'] if mode == 'gpu' else []), stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, universal_newlines=True, timeout=600, cwd=str(d), env=env)
        assert r.returncode == 0, r.stdout[-3000:]
        assert ('GPU solving using' in r.stdout) == (mode == 'gpu')
        found = {}
        for base, _, names in os.walk(str(d)):
            for n in names:
                if n.endswith(('.vti', '.vtp')):
                    with open(os.path.join(base, n), 'rb') as f:
                        found[os.path.relpath(os.path.join(base, n), str(d))] = f.read()
        files[mode] = found
    assert sorted(files['gpu']) == sorted(files['cpu']) and len(files['gpu']) == 5
    for name, ref in files['cpu'].items():
        got = files['gpu'][name]
        if 'snap' not in name:
            assert got == ref, name
            continue
        assert len(got) == len(ref), name
        start = ref.index(b'<AppendedData encoding="raw">\n_') + len(b'<AppendedData encoding="raw">\n_')
        assert got[:start] == ref[:start], name
        n = int(np.frombuffer(ref[start:start + 4], dtype=np.uint32)[0])
        for off in (start, start + 4 + n):
            assert got[off:off + 4] == ref[off:off + 4]
            a = np.frombuffer(got[off + 4:off + 4 + n], dtype=np.float32)
            b = np.frombuffer(ref[off + 4:off + 4 + n], dtype=np.float32)
            assert np.abs(b).max() > 0
            assert np.abs(a - b).max() <= 1e-4 * np.abs(b).max(), (name, off)
