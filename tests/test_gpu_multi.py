"""Two real GPUs, NCCL: the sharded generation equals the single-GPU generation bit for bit in fitness and to
fp32 summation order in the update.  Skipped on a one-GPU box (the CPU gloo test covers the host logic there)."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip('torch')
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, N, precision, use_graph, outdir, gens, comm):
    sys.path.insert(0, REPO)
    os.environ['DES_COMM'] = comm           # 'peer': kernels of this library over NVLink peer memory; 'nccl': two all-reduces
    from distributedes_b200.engine import NESEngine
    from oracle import nes_oracle as orc
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    try:
        d0, H, A, T = 24, 64, 4, 256
        obs, target = orc.synthetic_tape(T, d0, A)
        eng = NESEngine(state_dim=d0, hidden=H, action_dim=A, pop_size=N, theta0=orc.synthetic_theta(d0, H, A), obs=obs,
                        target=target, sigma=0.1, learning_rate=0.1, clip=1.0, seed=21, precision=precision,
                        device='cuda:%d' % rank, use_graph=use_graph)
        fit0 = theta1 = None
        for g in range(gens):
            eng.generation()
            if g == 0:
                fit0 = eng.fitness_all.cpu().numpy()
                theta1 = eng.theta.cpu().numpy()
        torch.cuda.synchronize()
        # results go through files: a SimpleQueue pipe (64 KB) would block the child while the parent joins
        assert (eng.comm is not None) == (comm == 'peer'), 'exchange path %r was requested' % comm
        np.savez(os.path.join(outdir, 'rank%d.npz' % rank), theta=eng.theta.cpu().numpy(), fitness=fit0, theta1=theta1)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('N,precision,use_graph,comm', [(1000, 'f16x3', False, 'peer'), (1001, 'fp32', False, 'nccl'),
                                                        (4096, 'f16', True, 'peer'), (1001, 'fp32', True, 'peer')])
def test_two_gpu_generation_matches_one_gpu(N, precision, use_graph, comm):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    sys.path.insert(0, REPO)
    from distributedes_b200.engine import NESEngine
    from oracle import nes_oracle as orc
    import tempfile
    with tempfile.TemporaryDirectory() as outdir:
        mp.spawn(_worker, args=(2, 29700 + N % 50 + (7 if use_graph else 0), N, precision, use_graph, outdir, 3, comm), nprocs=2, join=True)
        res = []
        for r in range(2):
            z = np.load(os.path.join(outdir, 'rank%d.npz' % r))
            res.append((r, z['theta'], z['fitness'], z['theta1']))
    assert np.array_equal(res[0][1], res[1][1])            # identical parameters on both ranks, no broadcast
    assert np.array_equal(res[0][2], res[1][2])
    d0, H, A, T = 24, 64, 4, 256
    obs, target = orc.synthetic_tape(T, d0, A)
    one = NESEngine(state_dim=d0, hidden=H, action_dim=A, pop_size=N, theta0=orc.synthetic_theta(d0, H, A), obs=obs,
                    target=target, sigma=0.1, learning_rate=0.1, clip=1.0, seed=21, precision=precision, device='cuda:0')
    one.generation()
    # generation-0 fitness of a member does not depend on the sharding: bit-identical
    assert np.array_equal(one.fitness_all.cpu().numpy(), res[0][2])
    # the first update differs only by the order of the cross-shard fp32 sum.  (Later generations are not compared:
    # parameters that differ in the last bit flip near-tied ranks, which moves the update by ~5/N^1.5 per flip.)
    th0 = orc.synthetic_theta(d0, H, A)
    th1 = one.theta.cpu().numpy()
    assert np.linalg.norm(th1 - res[0][3]) <= 1e-5 * np.linalg.norm(th1 - th0)


def _cma_worker(rank, world, port, outdir):
    sys.path.insert(0, REPO)
    from distributedes_b200.cma_es import CMAEvolutionStrategy
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    try:
        n, lam = 512, 130                                   # ragged shards
        m0 = np.random.RandomState(0).randn(n)
        es = CMAEvolutionStrategy(m0, 1.0, lam, seed=6, device='cuda:%d' % rank)
        X = es.ask()
        cost = es.gather_cost((X.double() ** 2).sum(1).float())
        es.tell(X, cost)
        torch.cuda.synchronize()
        np.savez(os.path.join(outdir, 'cma%d.npz' % rank), C=es.C.cpu().numpy(), m=es.m.cpu().numpy(), sigma=es.sigma,
                 X=X.cpu().numpy(), cost=cost.cpu().numpy())
    finally:
        dist.destroy_process_group()


def test_two_gpu_cma_generation_matches_one_gpu():
    """cma_es.CMAEvolutionStrategy sharded over 2 GPUs (all-reduce of the [n,n] rank-mu partials) against the same
    generation on one GPU: generation 0 has B = I, so both sample identical solutions from the counter noise."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    sys.path.insert(0, REPO)
    from distributedes_b200.cma_es import CMAEvolutionStrategy
    import tempfile
    with tempfile.TemporaryDirectory() as outdir:
        mp.spawn(_cma_worker, args=(2, 29761, outdir), nprocs=2, join=True)
        r = [np.load(os.path.join(outdir, 'cma%d.npz' % k)) for k in range(2)]
    for k in ('C', 'm', 'sigma', 'cost'):
        assert np.array_equal(r[0][k], r[1][k]), k
    n, lam = 512, 130
    one = CMAEvolutionStrategy(np.random.RandomState(0).randn(n), 1.0, lam, seed=6, device='cuda:0')
    X = one.ask()
    assert np.array_equal(X.cpu().numpy(), np.concatenate([r[0]['X'], r[1]['X']]))
    cost = (X.double() ** 2).sum(1).float()
    one.tell(X, cost)
    C1 = one.C.cpu().numpy()
    assert np.linalg.norm(C1 - r[0]['C']) <= 1e-6 * np.linalg.norm(C1)          # order of the cross-shard fp32 sum
    assert np.linalg.norm(one.m.cpu().numpy() - r[0]['m']) <= 1e-12 * np.linalg.norm(r[0]['m'])
