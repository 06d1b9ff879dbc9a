"""Pin the oracle against the golden vectors generated from the UNMODIFIED reference
(tests/golden/make_golden.py ran gprMax v3.1.7's own CPU solver on each model).

  * every float64 fixture: the oracle must reproduce the reference to <= 1e-10 of peak
    (only FMA-contraction differences remain, ~1e-13);
  * float32 fixtures: <= 1e-4 of peak, or -- on the models where float32 rounding noise alone
    exceeds that (see parity.compare_f32_with_truth) -- as close to the reference's float64
    result as the reference's own float32 result is;
  * with the reference's compiled kernels as back-end (oracle/_ref) the restated time loop,
    sources, receivers, transmission line and snapshots must be BIT-EXACT against the goldens.
Long models are truncated: the golden trace prefix is compared.
"""
import numpy as np
import pytest

from conftest import golden_names, golden_path, have_ref_kernels
from gprmax_b200.model_io import load_model
from oracle.solver import solve_cpu
from parity import compare_f32_with_truth, compare_traces, oracle_outputs_as_golden, tolerance

MAX_ITERS = 400   # keeps the CPU suite to a few minutes


def _run(name, variant, kernels):
    G, golden = load_model(golden_path(name, variant))
    nit = min(G.iterations, MAX_ITERS)
    # snapshots / full-length-only outputs need every iteration
    if G.snapshots:
        nit = max(nit, max(s.time for s in G.snapshots))
    out = oracle_outputs_as_golden(G, solve_cpu(G, kernels=kernels, iterations=nit, nthreads=4))
    gold = {}
    for k, v in golden.items():
        gold[k] = v[:nit] if (k.startswith('rx') or k.startswith('tl')) else v
    return G, out, gold


@pytest.mark.parametrize('name', golden_names('f64'))
def test_oracle_f64(name, oracle_built):
    G, out, gold = _run(name, 'f64', 'oracle')
    worst, rep = compare_traces(out, gold, np.float64, tol=tolerance(G, np.float64))
    assert worst <= 1.0, rep


@pytest.mark.parametrize('name', golden_names('f32'))
def test_oracle_f32(name, oracle_built):
    G, out, gold = _run(name, 'f32', 'oracle')
    _, gold64 = load_model(golden_path(name, 'f64'))
    nit = len(next(v for k, v in gold.items() if k.startswith('rx')))
    gold64 = {k: (v[:nit] if (k.startswith('rx') or k.startswith('tl')) else v) for k, v in gold64.items()}
    ok, rep = compare_f32_with_truth(out, gold, gold64)
    assert ok, rep


@pytest.mark.skipif(not have_ref_kernels('f32'), reason='oracle/_ref not built')
@pytest.mark.parametrize('name', ['cylinder_Ascan_2D', 'pml_MRIPML_2', 'sources_mixed', 'transmission_line', 'dispersive_multipole', 'snapshots'])
def test_loop_restatement_bit_exact_with_reference_kernels(name):
    G, out, gold = _run(name, 'f32', 'ref')
    for k, v in gold.items():
        assert np.array_equal(np.asarray(out[k]).reshape(v.shape), v), k
