"""Picklable model builder for the farm test (worker processes are spawned): trace k of a small B-scan = the `sources_mixed`
fixture with its Hertzian dipole and second receiver moved k cells along x."""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(k):
    from gprmax_b200.model_io import load_model
    G, _ = load_model(os.path.join(ROOT, 'tests', 'golden', 'sources_mixed_f32.npz'))
    for s in G.hertziandipoles:
        s.xcoord += k
    G.rxs[1].xcoord += k
    return G
