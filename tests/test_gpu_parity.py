"""GPU parity: the CUDA core (through the C ABI, via the drop-in solve_gpu) against
  * the golden vectors written by the UNMODIFIED reference CPU solver (tests/golden/*.npz), and
  * the CPU oracle run here on the same model.
Every fixture covers a different part of the hot path: 2-D modes, 3-D,
HORIPML/MRIPML orders 1-2, dispersive 1-pole/multi-pole, magnetic dipole / voltage sources /
transmission line, Ix-Iz receivers, snapshots, 50-material fractal soil.
float64: <= 1e-10 of trace peak.  float32: see parity.compare_f32_with_truth.
"""
import os

import numpy as np
import pytest

from conftest import golden_names, golden_path
from parity import collect_outputs, compare_f32_with_truth, compare_traces, oracle_outputs_as_golden, tolerance

pytestmark = pytest.mark.gpu


def _solve(G):
    from gprmax_b200 import GPU, solve_gpu
    G.gpu = GPU(0)
    tsolve, mem = solve_gpu(1, 1, G)
    assert tsolve > 0 and mem > 0
    return collect_outputs(G)


@pytest.mark.parametrize('name', golden_names('f64'))
def test_f64_against_reference_golden(name):
    from gprmax_b200.model_io import load_model
    G, golden = load_model(golden_path(name, 'f64'))
    out = _solve(G)
    worst, rep = compare_traces(out, golden, np.float64, tol=tolerance(G, np.float64), model=name)
    print(rep)
    assert worst <= 1.0, rep


@pytest.mark.parametrize('name', golden_names('f32'))
def test_f32_against_reference_golden(name):
    from gprmax_b200.model_io import load_model
    G, golden = load_model(golden_path(name, 'f32'))
    out = _solve(G)
    if os.path.exists(golden_path(name, 'f64')):
        _, golden64 = load_model(golden_path(name, 'f64'))
        ok, rep = compare_f32_with_truth(out, golden, golden64, model=name)
        print(rep)
        assert ok, rep
    else:
        worst, rep = compare_traces(out, golden, np.float32, model=name)
        print(rep)
        assert worst <= 1.0, rep


@pytest.mark.parametrize('name,variant', [('pml_HORIPML_1', 'f64'), ('pml_MRIPML_2', 'f64'), ('sources_mixed', 'f64'),
                                          ('dispersive_multipole', 'f64'), ('cylinder_Ascan_2D', 'f64'),
                                          ('transmission_line', 'f64'), ('pml_HORIPML_2', 'f64'), ('snapshots', 'f64'), ('heterogeneous_soil_small', 'f64')])
def test_against_oracle_run_here(name, variant, oracle_built):
    """Same model through the CUDA core and through oracle/ on this box (no stored vectors)."""
    from gprmax_b200.model_io import load_model
    from oracle.solver import solve_cpu
    G, _ = load_model(golden_path(name, variant))
    ref = oracle_outputs_as_golden(G, solve_cpu(G, kernels='oracle'))
    out = _solve(G)
    real = np.float64 if variant == 'f64' else np.float32
    worst, rep = compare_traces(out, ref, real, tol=tolerance(G, real))
    print(rep)
    assert worst <= 1.0, rep


def test_final_fields_match_oracle(oracle_built):
    """Whole-volume check (not just receiver points): final E/H of a PML + dielectric model in float64."""
    from gprmax_b200 import Solver
    from gprmax_b200.model_io import load_model
    from oracle.solver import solve_cpu
    G, _ = load_model(golden_path('pml_MRIPML_2', 'f64'))
    S = solve_cpu(G, kernels='oracle', keep_state=True)['state']
    with Solver(G, device_id=0) as sv:
        sv.run()
        for c, n in enumerate(('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz')):
            dev = sv.get_field(c)
            ref = getattr(S, n)
            assert np.abs(dev - ref).max() <= 1e-10 * np.abs(ref).max(), n


def test_restart_and_chunked_run_bit_identical():
    """gpb_reset + running in chunks reproduces a single run bit for bit (graph replay vs plain launches)."""
    from gprmax_b200 import Solver
    from gprmax_b200.model_io import load_model
    G, _ = load_model(golden_path('sources_mixed', 'f32'))
    with Solver(G, device_id=0) as sv:
        sv.run()
        a = sv.receivers()
        sv.reset()
        for n in (1, 7, 100, G.iterations - 108):
            sv.run(n)
        b = sv.receivers()
        assert sv.kernel_launches > 0
    assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['sources_mixed', 'transmission_line', 'snapshots', 'pml_MRIPML_2', 'hertzian_dipole_dispersive'])
def test_launch_side_switches_bit_identical(name, monkeypatch):
    """How the step is launched does not change a bit of the result: programmatic dependent launches on / off (GPB_PDL),
    16 / 5 / 1 iterations per CUDA graph (GPB_GRAPH_ITERS; with more than one the electric-phase sources and the next
    iteration's prologue share a launch unless GPB_FUSE_BEGIN=0), plain launches (GPB_NO_GRAPH), each on the register-vectorised
    and on the TMA kernels, run in one piece and in ragged chunks (snapshot iterations leave the graph)."""
    from gprmax_b200 import Solver
    from gprmax_b200.model_io import load_model
    G, _ = load_model(golden_path(name, 'f32'))

    def run(env, chunks=None):
        for k in ('GPB_PDL', 'GPB_GRAPH_ITERS', 'GPB_FUSE_BEGIN', 'GPB_NO_GRAPH', 'GPB_FORCE_TMA'):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with Solver(G, device_id=0) as sv:
            if chunks:
                done = 0
                for n in chunks:
                    n = min(n, G.iterations - done)
                    sv.run(n)
                    done += n
                sv.run(G.iterations - done)
            else:
                sv.run()
            out = [sv.receivers()] + [np.stack(sv.snapshot(q)) for q in range(len(G.snapshots))]
            out += [np.stack(sv.tline(q)) for q in range(len(G.transmissionlines))]
        return out

    ref = run({'GPB_PDL': '0', 'GPB_NO_GRAPH': '1'})
    for tma in ({}, {'GPB_FORCE_TMA': '1'}):
        base = run(dict(tma, GPB_PDL='0', GPB_NO_GRAPH='1'))
        for env, chunks in (({}, None), ({}, (1, 17, 3, 40, 16, 33)), ({'GPB_PDL': '0'}, None), ({'GPB_GRAPH_ITERS': '1'}, None),
                            ({'GPB_GRAPH_ITERS': '5'}, (7, 23)), ({'GPB_FUSE_BEGIN': '0'}, None), ({'GPB_NO_GRAPH': '1'}, None)):
            got = run(dict(tma, **env), chunks)
            assert len(got) == len(base)
            for a, b in zip(got, base):
                assert np.array_equal(a, b), (name, tma, env, chunks)
    assert len(ref) == len(base)


def test_config2_bench_300_full_trace():
    """BASELINE.json configs[1], the headline benchmark at its own size: tests/benchmarking/bench_300x300x300.in, all 1559
    iterations, against the trace of the unmodified reference CPU solver (tests/golden/make_golden.py `bench_300_trace`, float32
    and float64 runs, 3 and 6 minutes on 8 cores).  The grid is rebuilt by benchkit.synthetic (bit-identical to the reference
    build) -- or by the reference front end itself when baseline/_ref is present."""
    p32, p64 = golden_path('bench_300_trace', 'f32'), golden_path('bench_300_trace', 'f64')
    if not (os.path.exists(p32) and os.path.exists(p64)):
        pytest.skip('fixture not present')
    from benchkit.synthetic import bench_model
    z32, z64 = np.load(p32), np.load(p64)
    golden = {k[len('golden_'):]: z32[k] for k in z32.files}
    golden64 = {k[len('golden_'):]: z64[k] for k in z64.files}
    G = bench_model(300)
    assert G.iterations == 1559 and len(golden['rx0_Ex']) == 1559
    out = _solve(G)
    ok, rep = compare_f32_with_truth(out, golden, golden64, model='bench_300 (config 2, full size, 1559 iterations)')
    print(rep)
    assert ok, rep


def test_config4_gssi_bscan_trace(tmp_path):
    """BASELINE.json configs[3]: user_models/cylinder_Bscan_GSSI_1500.in, trace 1 of the B-scan at full size
    (480 x 148 x 235 cells, 3117 iterations, GSSI 1.5 GHz antenna model: 24 materials, PEC plates, a 230-ohm
    voltage source, one Ey receiver).  Golden = the unmodified reference CPU solver (408 s on 8 cores here)."""
    import time
    from gprmax_b200.model_io import load_model
    path = golden_path('bscan_gssi_trace1', 'f32')
    if not os.path.exists(path):
        pytest.skip('fixture not present')
    G, golden = load_model(path)
    t0 = time.perf_counter()
    out = _solve(G)
    print('config 4 trace: {:.2f} s end to end, {} cells x {} iterations'.format(time.perf_counter() - t0, G.nx * G.ny * G.nz, G.iterations))
    worst, rep = compare_traces(out, golden, np.float32, model='bscan_gssi_trace1 (config 4, full size)')
    print(rep)
    assert worst <= 1.0, rep


def test_config3_heterogeneous_soil_full_size():
    """BASELINE.json configs[2] at full size with explicit fractal seeds: 150 x 150 x 100 cells, 3117 iterations,
    50-bin Peplinski soil (1-pole Debye everywhere), rough surface, 6 PML slabs -- the dispersive update path."""
    from gprmax_b200.model_io import load_model
    p32 = golden_path('heterogeneous_soil_full', 'f32')
    p64 = p32.replace('_f32.npz', '_f64_truth.npz')   # float64 reference traces only (no second copy of the 8 MB ID array)
    if not (os.path.exists(p32) and os.path.exists(p64)):
        pytest.skip('fixture not present')
    import time
    G, golden = load_model(p32)
    z = np.load(p64)
    golden64 = {k[len('golden_'):]: z[k] for k in z.files}
    t0 = time.perf_counter()
    out = _solve(G)
    print('config 3: {:.2f} s end to end, {} cells x {} iterations, {} materials'.format(time.perf_counter() - t0, G.nx * G.ny * G.nz, G.iterations, G.updatecoeffsE.shape[0]))
    ok, rep = compare_f32_with_truth(out, golden, golden64, model='heterogeneous_soil_full (config 3, full size)')
    print(rep)
    assert ok, rep


TMA_FIXTURES = ['pml_HORIPML_1', 'pml_HORIPML_2', 'pml_MRIPML_1', 'pml_MRIPML_2', 'sources_mixed', 'transmission_line',
                'snapshots', 'hertzian_dipole_hs', 'dispersive_multipole', 'hertzian_dipole_dispersive', 'heterogeneous_soil_small']


@pytest.mark.parametrize('name', TMA_FIXTURES)
@pytest.mark.parametrize('mode', ['tma', 'tma_pair', 'tma_nopersist', 'tma_pw0', 'tma_tile8x128', 'tma_zsplit', 'tma_znocoop', 'tma_ids16', 'tma_ids32', 'v4_ids16', 'scalar',
                                  'tma_tpf1', 'tma_tpf0', 'tma_disp_v4', 'tma_disp_complex', 'v4_disp_complex'])
def test_f64_every_kernel_path(name, mode, monkeypatch):
    """The small fixtures normally run on the register-vectorised kernels (the TMA kernels are only selected above
    2.5 M nodes); force each kernel family in turn so that all of them are held to the 1e-10 bar:
    TMA-staged persistent CTAs with a producer warp (default), one CTA per work item, producer = thread 0 of a 16 x 64 tile,
    another tile shape, z-slab PML in its own kernel, z-slab PML per thread instead of one cell per lane, generic scalar."""
    from gprmax_b200.model_io import load_model
    if mode.startswith('tma'):
        monkeypatch.setenv('GPB_FORCE_TMA', '1')
    if mode == 'tma_pair':
        monkeypatch.setenv('GPB_PAIR', '1')
    if mode == 'tma_nopersist':
        monkeypatch.setenv('GPB_TMA_NOPERSIST', '1')
    if mode == 'tma_pw0':
        monkeypatch.setenv('GPB_TMA_PW', '0')
    if mode == 'tma_zsplit':
        monkeypatch.setenv('GPB_TMA_ZSPLIT', '1')
    if mode == 'tma_znocoop':
        monkeypatch.setenv('GPB_TMA_ZNOCOOP', '1')
    if mode.endswith('ids16'):   # 16- / 32-bit device IDs (models with more than 256 / 65536 materials)
        monkeypatch.setenv('GPB_ID_BYTES', '2')
    if mode.endswith('ids32'):
        monkeypatch.setenv('GPB_ID_BYTES', '4')
    if mode == 'tma_tile8x128':
        monkeypatch.setenv('GPB_TMA_TZ', '128')
    if mode == 'scalar':
        monkeypatch.setenv('GPB_SCALAR', '1')
    # dispersive E half-step: T prefetch distance 1 / direct loads, register-vectorised E kernel next to the TMA H kernel,
    # complex T even for Debye media
    if mode in ('tma_tpf1', 'tma_tpf0', 'tma_disp_v4', 'tma_disp_complex', 'v4_disp_complex'):
        if name not in ('dispersive_multipole', 'hertzian_dipole_dispersive', 'heterogeneous_soil_small'):
            pytest.skip('dispersive models only')
        if mode.startswith('tma_tpf'):
            monkeypatch.setenv('GPB_TMA_TPF', mode[-1])
        if mode == 'tma_disp_v4':
            monkeypatch.setenv('GPB_DISP_V4', '1')
        if mode.endswith('disp_complex'):
            monkeypatch.setenv('GPB_DISP_COMPLEX', '1')
    G, golden = load_model(golden_path(name, 'f64'))
    out = _solve(G)
    worst, rep = compare_traces(out, golden, np.float64, tol=tolerance(G, np.float64))
    assert worst <= 1.0, rep


def test_kernel_families_bit_identical_on_heterogeneous_grid(monkeypatch):
    """A grid large enough for the persistent TMA kernels (many work items per CTA) with a random mix of materials on
    every edge: the TMA-staged kernels (persistent / one CTA per item / thread-0 producer / z slabs in their own kernel)
    and the register-vectorised kernels must produce the same bits, twice in a row.  Catches stage-release races (a
    consumer reading material IDs after handing the stage back read the next tile's IDs -- harmless on homogeneous
    benchmarks, timing-dependent here) and any difference in FMA placement between kernel families, which a sharded run
    next to its single-GPU reference would expose."""
    sys_path = os.path.join(os.path.dirname(os.path.abspath(__file__)))
    import sys
    sys.path.insert(0, sys_path)
    from gprmax_b200 import Solver
    from benchkit.synthetic import homogeneous_model, material_rows
    nx, ny, nz, its = 160, 144, 128, 70
    # the source sits 3 cells outside the x0 / ymax / z0 PML slabs, so that within the run a non-zero field fills the edges and
    # the corner where two and three slab corrections hit the same cell (every family must apply them in the same order)
    G = homogeneous_model((nx, ny, nz), iterations=its, er=6.0, se=0.01, src=(13e-3, (ny - 13) * 1e-3, 13e-3), src_pol='z',
                          rxs=[(5e-3, (ny - 4) * 1e-3, 3e-3), (16e-3, (ny - 15) * 1e-3, 14e-3)])
    real = G.updatecoeffsE.dtype
    rows = [material_rows(er, se, 1.0, 0.0, G.dx, G.dy, G.dz, G.dt, real) for er, se in ((3.0, 0.001), (9.0, 0.02), (4.5, 0.0), (12.0, 0.05))]
    G.updatecoeffsE = np.concatenate([G.updatecoeffsE, np.stack([r[0] for r in rows])])
    G.updatecoeffsH = np.concatenate([G.updatecoeffsH, np.stack([r[1] for r in rows])])
    rng = np.random.default_rng(7)
    G.ID = rng.integers(2, G.updatecoeffsE.shape[0], size=G.ID.shape, dtype=np.uint32)

    def run(env):
        for k in ('GPB_NO_TMA', 'GPB_TMA_NOPERSIST', 'GPB_TMA_PW', 'GPB_TMA_ZSPLIT', 'GPB_TMA_ZNOCOOP', 'GPB_PAIR'):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with Solver(G, device_id=0) as sv:
            sv.run()
            return [sv.get_field(c) for c in range(6)] + [sv.receivers()]

    ref = run({'GPB_NO_TMA': '1'})
    assert np.abs(ref[-1]).max() > 0
    # the field has reached the x0 / ymax / z0 corner region of the PML (all three slabs overlap there)
    for c in range(6):
        assert np.abs(ref[c][1:9, ny - 9:ny - 1, 1:9]).max() > 0, c
    # GPB_PAIR: both half-steps of an iteration in one launch (k_update_pair)
    for env in ({}, {}, {'GPB_PAIR': '1'}, {'GPB_TMA_NOPERSIST': '1'}, {'GPB_TMA_PW': '0'}, {'GPB_TMA_ZSPLIT': '1'}, {'GPB_TMA_ZNOCOOP': '1'}):
        out = run(env)
        for c, (a, b) in enumerate(zip(out, ref)):
            assert np.array_equal(a, b), (env, c, int((a != b).sum()))


def test_dispersive_families_and_real_T_bit_identical(monkeypatch):
    """Dispersive E half-step: the TMA-staged kernel (T prefetched per thread, distance 2 / 1 / direct), the register-vectorised
    kernel and the scalar kernel, with real-valued T (Debye media) and with complex T, must all produce the same bits: the
    per-cell arithmetic is one explicitly rounded routine (disp_cell_r / disp_cell_c), and with real coefficients Im(T) stays zero.
    Model: the 50-material Peplinski soil fixture (1-pole Debye, rough surface, PML) in float32, fields compared everywhere."""
    from gprmax_b200 import Solver
    from gprmax_b200.model_io import load_model
    G, _ = load_model(golden_path('heterogeneous_soil_small', 'f32'))
    keys = ('GPB_FORCE_TMA', 'GPB_NO_TMA', 'GPB_SCALAR', 'GPB_DISP_V4', 'GPB_DISP_COMPLEX', 'GPB_TMA_TPF')

    def run(env):
        for k in keys:
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with Solver(G, device_id=0) as sv:
            path = sv.kernel_path
            sv.run()
            return path, [sv.get_field(c) for c in range(6)] + [sv.receivers()]

    pref, ref = run({'GPB_FORCE_TMA': '1'})
    assert 'k_update_tma' in pref and 'DISP=real' in pref, pref
    assert np.abs(ref[-1]).max() > 0
    for env in ({'GPB_FORCE_TMA': '1', 'GPB_TMA_TPF': '1'}, {'GPB_FORCE_TMA': '1', 'GPB_TMA_TPF': '0'}, {'GPB_FORCE_TMA': '1', 'GPB_DISP_COMPLEX': '1'},
                {'GPB_FORCE_TMA': '1', 'GPB_DISP_V4': '1'}, {'GPB_NO_TMA': '1'}, {'GPB_NO_TMA': '1', 'GPB_DISP_COMPLEX': '1'}):
        path, out = run(env)
        for c, (a, b) in enumerate(zip(out, ref)):
            assert np.array_equal(a, b), (env, path, c, int((a != b).sum()))


@pytest.mark.parametrize('name', ['pml_HORIPML_2', 'pml_MRIPML_1', 'hertzian_dipole_dispersive', 'heterogeneous_soil_small', 'bench_100', 'snapshots', 'sources_mixed'])
def test_pair_kernel_bit_identical(name, monkeypatch):
    """TMA kernels, opt-in GPB_PAIR=1: both half-steps of an iteration run in ONE launch (k_update_pair: H and E work items from one
    queue, an E item loaded only after the H items it depends on were published through per-chunk progress counters, so that
    E finds its operands in L2).  Same bits as one launch per half-step and as the register-vectorised kernels, three runs in a
    row (a missed dependency would show as run-to-run jitter)."""
    from gprmax_b200 import Solver
    from gprmax_b200.model_io import load_model
    G, _ = load_model(golden_path(name, 'f32'))

    def run(env):
        for k in ('GPB_FORCE_TMA', 'GPB_NO_TMA', 'GPB_PAIR', 'GPB_PAIR_XCHUNK', 'GPB_PAIR_LAG'):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with Solver(G, device_id=0) as sv:
            path = sv.kernel_path
            sv.run()
            return path, [sv.get_field(c) for c in range(6)] + [sv.receivers()]

    _, ref = run({'GPB_NO_TMA': '1'})
    pseq, seq = run({'GPB_FORCE_TMA': '1'})
    assert 'k_update_tma' in pseq and 'k_update_pair' not in pseq
    for rep in range(3):
        pcon, con = run({'GPB_FORCE_TMA': '1', 'GPB_PAIR': '1', 'GPB_PAIR_XCHUNK': ('4', '2', '8')[rep], 'GPB_PAIR_LAG': ('1', '0', '3')[rep]})
        # (a magnetic dipole acts between the half-steps, other tile shapes have no pair kernel: those keep two launches)
        if name not in ('sources_mixed', 'heterogeneous_soil_small'):
            assert 'k_update_pair' in pcon, pcon
        for c, (a, b, d) in enumerate(zip(con, seq, ref)):
            assert np.array_equal(a, b) and np.array_equal(a, d), (name, rep, c)


@pytest.mark.parametrize('name,variant', [('cylinder_Ascan_2D', 'f32'), ('2D_ExHyHz', 'f64'), ('pml_HORIPML_2', 'f32'), ('pml_MRIPML_2', 'f64'), ('sources_mixed', 'f32'),
                                          ('snapshots', 'f32'), ('hertzian_dipole_dispersive', 'f32'), ('dispersive_multipole', 'f64'),
                                          ('heterogeneous_soil_small', 'f32'), ('bench_100', 'f32')])
def test_cooperative_whole_run_kernel_bit_identical(name, variant, monkeypatch):
    """Opt-in GPB_COOP=1: small grids run n iterations in ONE cooperative launch (k_run_coop: grid-wide barrier between the half-steps; z-slab PML,
    point sources and receiver samples done by the block that owns the cells).  Same bits as the kernel-per-half-step graph
    path -- receivers (all nine rows), snapshots, final fields -- also when the run is cut into uneven pieces."""
    from gprmax_b200 import Solver
    from gprmax_b200.model_io import load_model
    G, _ = load_model(golden_path(name, variant))

    def run(env, pieces=None):
        for k in ('GPB_COOP', 'GPB_COOP_XCHUNK'):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with Solver(G, device_id=0) as sv:
            path = sv.kernel_path
            for n in (pieces or [G.iterations]):
                sv.run(n)
            assert sv.iteration == G.iterations
            return path, [sv.get_field(c) for c in range(6)] + [sv.receivers()] + [a for k in range(len(G.snapshots)) for a in sv.snapshot(k)]

    pref, ref = run({})
    assert 'k_run_coop' not in pref
    assert np.abs(ref[6]).max() > 0
    its = G.iterations
    for env, pieces in (({'GPB_COOP': '1'}, None), ({'GPB_COOP': '1'}, [1, 2, 37, its - 40]), ({'GPB_COOP': '1', 'GPB_COOP_XCHUNK': '1'}, None),
                        ({'GPB_COOP': '1', 'GPB_COOP_XCHUNK': '16'}, [its // 2, its - its // 2])):
        path, out = run(env, pieces)
        assert 'k_run_coop' in path, path
        for c, (a, b) in enumerate(zip(out, ref)):
            assert np.array_equal(a, b), (name, env, pieces, c, int((a != b).sum()))


def test_device_memory_cache_reuse_and_release():
    """gpb_destroy keeps the device blocks for the next gpb_create of the same size (a B-scan creates one solver per trace):
    a second solver built from stale cached blocks must give the same bits as the first, and gpb_release_cached must work."""
    from gprmax_b200 import Solver, _lib
    from gprmax_b200.model_io import load_model
    G, _ = load_model(golden_path('pml_HORIPML_2', 'f32'))
    outs = []
    for rep in range(3):
        with Solver(G, device_id=0) as sv:
            sv.run()
            outs.append(sv.receivers())
        if rep == 1:
            assert _lib.lib().gpb_release_cached() == 0
    assert np.abs(outs[0]).max() > 0
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
