"""torchrun worker for tests/test_gpu_sharded.py: runs one model sharded over all ranks (NCCL),
rank 0 stores the receiver array to the path given on the command line."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(spec):
    """A fixture path, or 'synthetic:nx,ny,nz,iterations' = homogeneous lossy dielectric box (the recipe of the sharded
    benchmark: er 6, sigma 0.01, z dipole at the centre, one receiver on the middle cut plane, one off-centre)."""
    from gprmax_b200.model_io import load_model
    if not spec.startswith('synthetic'):
        return load_model(spec)[0]
    from benchkit.synthetic import homogeneous_model
    nx, ny, nz, its = [int(v) for v in spec.split(':')[1].split(',')]
    if spec.startswith('synthetic_cut:'):
        # y-directed Hertzian dipole ON the first plane of the rank that owns the middle of the domain for 2 and 4 ranks: its
        # term must be in the plane before the plane is sent to the left neighbour
        from gprmax_b200.sharded import partition_planes
        cut = partition_planes(nx, 2)[1][0]
        assert cut == partition_planes(nx, 4)[2][0]
        return homogeneous_model((nx, ny, nz), iterations=its, er=4.0, se=0.005, src=(cut * 1e-3, ny // 2 * 1e-3, nz // 2 * 1e-3), src_pol='y',
                                 rxs=[((cut - 6) * 1e-3, (ny // 2 + 3) * 1e-3, nz // 2 * 1e-3), ((cut + 5) * 1e-3, (ny // 2 - 4) * 1e-3, (nz // 2 + 2) * 1e-3)])
    return homogeneous_model((nx, ny, nz), iterations=its, er=6.0, se=0.01, src=(nx // 2 * 1e-3, ny // 2 * 1e-3, nz // 2 * 1e-3), src_pol='z',
                             rxs=[((nx // 2 + 1) * 1e-3, (ny // 2 + 9) * 1e-3, nz // 2 * 1e-3), ((nx // 4) * 1e-3, (ny // 3) * 1e-3, (nz // 2 + 5) * 1e-3)])


def main():
    import torch.distributed as dist
    from gprmax_b200.model_io import load_model
    from gprmax_b200.sharded import solve_gpu_sharded
    fixture, out, overlap = sys.argv[1], sys.argv[2], sys.argv[3] == '1'
    transport = sys.argv[4] if len(sys.argv) > 4 else 'nccl'
    dist.init_process_group('nccl')
    G = build(fixture)
    res = {}
    rxs, seconds = solve_gpu_sharded(G, overlap=overlap, transport=transport, results=res)
    if dist.get_rank() == 0:
        np.save(out, rxs)
        if res.get('snapshots'):
            np.savez(out + '.snaps.npz', **{'s{}_{}'.format(k, c): a for k, sn in enumerate(res['snapshots']) for c, a in enumerate(sn)})
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
