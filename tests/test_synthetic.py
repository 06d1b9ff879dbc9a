"""The reference-free builder of the benchmark models (gprmax_b200/synthetic.py) must reproduce,
bit for bit, what the reference's own model build produces for tests/benchmarking/bench_100x100x100.in
(golden fixture written by tests/golden/make_golden.py): ID array, coefficient tables, PML tables,
pre-sampled waveforms, source/receiver cells, time step."""
import numpy as np
import pytest

from conftest import golden_path
from gprmax_b200.model_io import load_model
from benchkit.synthetic import bench_model


@pytest.mark.parametrize('variant', ['f32', 'f64'])
def test_bench_model_matches_reference_build(variant):
    G, _ = load_model(golden_path('bench_100', variant))
    real = np.float32 if variant == 'f32' else np.float64
    S = bench_model(100, real=real, iterations=G.iterations)
    assert (S.nx, S.ny, S.nz, S.iterations) == (G.nx, G.ny, G.nz, G.iterations)
    assert S.dt == G.dt and (S.dx, S.dy, S.dz) == (G.dx, G.dy, G.dz)
    assert np.array_equal(S.ID, G.ID)
    assert np.array_equal(S.updatecoeffsE, G.updatecoeffsE) and S.updatecoeffsE.dtype == G.updatecoeffsE.dtype
    assert np.array_equal(S.updatecoeffsH, G.updatecoeffsH)
    assert len(S.pmls) == len(G.pmls) == 6
    for a, b in zip(S.pmls, G.pmls):
        assert (a.direction, a.xs, a.xf, a.ys, a.yf, a.zs, a.zf, a.thickness, a.d) == (b.direction, b.xs, b.xf, b.ys, b.yf, b.zs, b.zf, b.thickness, b.d)
        for t in ('ERA', 'ERB', 'ERE', 'ERF', 'HRA', 'HRB', 'HRE', 'HRF'):
            assert np.array_equal(getattr(a, t), getattr(b, t)), (a.direction, t)
    a, b = S.hertziandipoles[0], G.hertziandipoles[0]
    assert (a.xcoord, a.ycoord, a.zcoord, a.polarisation, a.dl) == (b.xcoord, b.ycoord, b.zcoord, b.polarisation, b.dl)
    assert np.array_equal(a.waveformvalues_wholestep, b.waveformvalues_wholestep)
    assert np.array_equal(a.waveformvalues_halfstep, b.waveformvalues_halfstep)
    assert (S.rxs[0].xcoord, S.rxs[0].ycoord, S.rxs[0].zcoord) == (G.rxs[0].xcoord, G.rxs[0].ycoord, G.rxs[0].zcoord)


def test_full_length_benchmark_has_published_iteration_count():
    # docs/source/benchmarking.rst: 3 ns window -> 1559 iterations at 1 mm cells
    S = bench_model(20)
    assert S.iterations == 1559
