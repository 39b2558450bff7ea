"""torchrun worker for tests/test_gpu_rollout.py::test_two_gpu_closed_loop_equals_one_gpu (2 ranks, NCCL)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributedes_b200.engine import RolloutEngine      # noqa: E402


def synthetic_theta():
    # same as oracle.nes_oracle.synthetic_theta(3, 64, 1) without importing the oracle in a product-side process
    from oracle import nes_oracle as orc
    return orc.synthetic_theta(3, 64, 1)


if __name__ == '__main__':
    rank = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl')
    eng = RolloutEngine(hidden=64, pop_size=37, theta0=synthetic_theta(), sigma=0.1, learning_rate=0.1, seed=3)
    eng.generation()
    fit0, stats0 = eng.fitness_all.cpu().numpy().copy(), eng.obs_stats.cpu().numpy().copy()
    eng.generation()
    np.savez(os.path.join(sys.argv[1], 'rank%d.npz' % rank), theta=eng.theta_numpy(), stats=eng.obs_stats.cpu().numpy(),
             fit=eng.fitness_all.cpu().numpy(), fit0=fit0, stats0=stats0)
    dist.destroy_process_group()
