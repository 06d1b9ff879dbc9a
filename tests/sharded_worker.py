"""torchrun worker for tests/test_gpu_sharded.py: runs one model sharded over all ranks (NCCL),
rank 0 stores the receiver array to the path given on the command line."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist
    from gprmax_b200.model_io import load_model
    from gprmax_b200.sharded import solve_gpu_sharded
    fixture, out, overlap = sys.argv[1], sys.argv[2], sys.argv[3] == '1'
    dist.init_process_group('nccl')
    G, _ = load_model(fixture)
    rxs, seconds = solve_gpu_sharded(G, overlap=overlap)
    if dist.get_rank() == 0:
        np.save(out, rxs)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
