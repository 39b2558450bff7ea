"""CMA-ES rank-mu covariance update: CUDA path vs the fp64 restatement (oracle/cma_oracle.py).
The oracle is parity-UNPINNED (pycma is not available; see its header) — the bar is 1e-5 relative."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import cma_oracle as cma

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def both_norms(got, ref, tol=1e-5):
    got = np.asarray(got, dtype=np.float64)
    assert np.linalg.norm(got - ref) <= tol * np.linalg.norm(ref)
    assert np.max(np.abs(got - ref)) <= tol * np.max(np.abs(ref))


@pytest.mark.parametrize('n,lam', [(1024, 256), (257, 37), (64, 5), (130, 200), (3000, 64), (4096, 1024)])
def test_rank_mu_matches_restatement(n, lam):
    from distributedes_b200 import ops
    rs = np.random.RandomState(n + lam)
    k = cma.cma_constants(n, lam)
    m_old = rs.randn(n)
    X = m_old + 0.7 * rs.randn(lam, n)
    Y, _ = cma.sort_and_scale(X, cma.sphere(X), m_old, 0.7)
    ref = cma.rank_mu_delta(Y, k['w'])
    Yt = torch.from_numpy(Y.astype(np.float32)).to(DEV)
    wt = torch.from_numpy(k['w'].astype(np.float32)).to(DEV)
    dC = ops.cma_rank_mu(Yt, wt)
    got = dC.cpu().numpy()
    assert np.array_equal(got, got.T)                      # exactly symmetric
    both_norms(got, ref)
    # full covariance update with a rank-one term
    C0 = np.eye(n) + 0.01 * np.cov(rs.randn(n, 2 * n))
    C0 = 0.5 * (C0 + C0.T)
    pc = rs.randn(n)
    Cref, decay = cma.cov_update(C0, ref, pc, k['c1'], k['cmu'], k['w'].sum())
    Ct = torch.from_numpy(C0.astype(np.float32)).to(DEV)
    ops.cma_cov_apply(Ct, dC, torch.from_numpy(pc.astype(np.float32)).to(DEV), decay=decay, c1=k['c1'], cmu=k['cmu'])
    both_norms(Ct.cpu().numpy(), Cref)


def test_rank_mu_shards_sum_to_whole_and_active_weights():
    """Population sharding: sum of per-shard partials == whole (the all-reduce contract), and signed weights."""
    from distributedes_b200 import ops
    n, lam = 512, 96
    rs = np.random.RandomState(0)
    k = cma.cma_constants(n, lam, active=True)
    Y = rs.randn(lam, n)
    ref = cma.rank_mu_delta(Y, k['w'])
    Yt = torch.from_numpy(Y.astype(np.float32)).to(DEV)
    wt = torch.from_numpy(k['w'].astype(np.float32)).to(DEV)
    parts = [ops.cma_rank_mu(Yt[a:b].contiguous(), wt[a:b].contiguous()) for a, b in [(0, 24), (24, 48), (48, 96)]]
    both_norms(sum(p.cpu().numpy().astype(np.float64) for p in parts), ref)
    empty = ops.cma_rank_mu(Yt[:0].contiguous(), wt[:0].contiguous())
    assert float(empty.abs().max()) == 0.0


def test_rank_mu_eight_shards_of_baseline_config5():
    """BASELINE configs[4]: n=4096, lambda=1024 split over 8 shards of 128 members; the all-reduce contract is
    sum(partials) == whole, checked here on one GPU by summing the eight partials in fp32 like NCCL would."""
    from distributedes_b200 import ops
    n, lam = 4096, 1024
    rs = np.random.RandomState(8)
    k = cma.cma_constants(n, lam)
    Y = rs.randn(lam, n)
    ref = cma.rank_mu_delta(Y, k['w'])
    Yt = torch.from_numpy(Y.astype(np.float32)).to(DEV)
    wt = torch.from_numpy(k['w'].astype(np.float32)).to(DEV)
    total = torch.zeros((n, n), dtype=torch.float32, device=DEV)
    for r in range(8):
        total += ops.cma_rank_mu(Yt[r * 128:(r + 1) * 128].contiguous(), wt[r * 128:(r + 1) * 128].contiguous())
    both_norms(total.cpu().numpy(), ref)
