"""BASELINE.json's full sizes, checked through size-independent properties (the fp64 oracle cannot run N*P = 4.8e9
normals in seconds): bijection / sortedness of ranks, linearity and shard-additivity of the fitness x noise reduction,
shard invariance and cross-precision agreement of the fused evaluation, trace / symmetry identities of the rank-mu term."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
N, D0, H, A, T = 65536, 24, 256, 4, 256        # configs[3]: NES population 65536, 2x256 MLP, obs dim 24


def _inputs():
    from distributedes_b200.envs import TapeEnv
    from distributedes_b200.model import StandardFCNet
    env = TapeEnv(D0, A, T)
    theta = StandardFCNet(D0, A, H, seed=0).get_weight()
    return (torch.from_numpy(theta).to(DEV), torch.from_numpy(env.obs).to(DEV), torch.from_numpy(env.target).to(DEV))


def test_ranks_are_a_sorted_bijection_at_65536():
    from distributedes_b200 import ops
    g = torch.Generator(device='cpu').manual_seed(1)
    f = torch.randn(N, generator=g).to(DEV)
    f[::1000] = f[7]                                            # some exact ties
    shaped, ranks = ops.centered_rank(f, return_ranks=True)
    r = ranks.long()
    assert torch.equal(torch.sort(r).values, torch.arange(N, device=DEV))          # a permutation of 0..N-1
    by_rank = f[torch.argsort(r)]
    assert bool((by_rank[1:] >= by_rank[:-1]).all())                                # sortedness
    tie = (f == f[7]).nonzero().flatten()
    assert bool((r[tie][1:] > r[tie][:-1]).all())                                   # ties keep index order
    assert abs(float(shaped.double().sum())) < 1e-2 and float(shaped.min()) == -0.5 and float(shaped.max()) == 0.5
    # shards (8 GPUs' worth) reproduce the global ranks
    for k in (0, 3, 7):
        _, rs = ops.centered_rank(f, member_offset=k * 8192, n_local=8192, return_ranks=True)
        assert torch.equal(rs, ranks[k * 8192:(k + 1) * 8192])


def test_reduction_is_linear_and_shard_additive_at_full_size():
    from distributedes_b200 import ops
    P = ops.param_count(D0, H, A)
    g = torch.Generator(device='cpu').manual_seed(2)
    s1 = (torch.rand(N, generator=g) - 0.5).to(DEV)
    s2 = (torch.rand(N, generator=g) - 0.5).to(DEV)
    kw = dict(seed=11, generation=3)
    p1 = ops.nes_grad_partial(s1, P, **kw).double()
    p2 = ops.nes_grad_partial(s2, P, **kw).double()
    p12 = ops.nes_grad_partial((s1 + 2 * s2).contiguous(), P, **kw).double()
    ref = p1 + 2 * p2
    assert float((p12 - ref).norm() / ref.norm()) < 2e-6                              # linearity in the shaped fitness
    parts = sum(ops.nes_grad_partial(s1[k * 8192:(k + 1) * 8192].contiguous(), P, member_offset=k * 8192, **kw).double()
                for k in range(8))
    assert float((parts - p1).norm() / p1.norm()) < 2e-6                              # all-reduce contract: shards add up
    # E[eps] = 0, Var = 1: a constant weight gives sum_i eps_ij ~ N(0, N)
    ones = ops.nes_grad_partial(torch.ones(N, device=DEV), P, **kw).double()
    assert abs(float(ones.mean())) < 5 * (N / P) ** 0.5 and abs(float(ones.std()) / N ** 0.5 - 1) < 0.02


def test_eval_shard_invariance_and_precision_agreement_at_full_size():
    from distributedes_b200 import ops
    theta, obs, target = _inputs()
    kw = dict(hidden=H, sigma=0.1, clip=1.0, seed=5, generation=9)
    full = ops.nes_eval(theta, obs, target, member_offset=0, n_local=N, precision='f16x3', **kw)
    assert bool(torch.isfinite(full).all())
    # a shard evaluated alone (other CTA <-> member assignment) gives the same bits
    for off, n in ((8192 * 5, 8192), (65000, 536), (12345, 7)):
        part = ops.nes_eval(theta, obs, target, member_offset=off, n_local=n, precision='f16x3', **kw)
        assert torch.equal(part, full[off:off + n])
    # the three arithmetic paths agree member by member on a slice
    sl = slice(30000, 30512)
    f32 = ops.nes_eval(theta, obs, target, member_offset=sl.start, n_local=512, precision='fp32', **kw)
    f16 = ops.nes_eval(theta, obs, target, member_offset=sl.start, n_local=512, precision='f16', **kw)
    assert float(((full[sl] - f32).abs() / f32.abs()).max()) < 3e-5
    assert float(((f16 - f32).abs() / f32.abs()).max()) < 4e-3
    # different noise for different generations / seeds
    other = ops.nes_eval(theta, obs, target, member_offset=0, n_local=64, precision='f16x3', hidden=H, sigma=0.1, clip=1.0,
                         seed=5, generation=10)
    assert not torch.equal(other, full[:64])


def test_rank_mu_identities_at_4096_by_1024():
    from distributedes_b200 import ops
    n, lam = 4096, 1024                                        # configs[4]
    g = torch.Generator(device='cpu').manual_seed(3)
    Y = torch.randn(lam, n, generator=g).to(DEV)
    w = torch.rand(lam, generator=g).to(DEV) / lam
    dC = ops.cma_rank_mu(Y, w)
    assert torch.equal(dC, dC.T)                                                      # exactly symmetric
    tr_ref = float((w.double() * (Y.double() ** 2).sum(1)).sum())
    # tr(sum w y y^T) = sum w |y|^2: a sum of 4096 x 1024 POSITIVE terms — the tensor cores' truncating fp32 accumulation
    # shows up here as a systematic few-1e-6 deficit (the FFMA path gives 1e-7); the 1e-5 contract of the op holds
    assert abs(float(torch.trace(dC.double())) - tr_ref) <= 6e-6 * tr_ref
    v = torch.randn(n, generator=g).to(DEV).double()
    quad_ref = float((w.double() * (Y.double() @ v) ** 2).sum())                      # v^T dC v = sum w (y.v)^2 >= 0
    assert abs(float(v @ (dC.double() @ v)) - quad_ref) <= 1e-5 * quad_ref
    halves = ops.cma_rank_mu(Y[:512].contiguous(), w[:512].contiguous()) + ops.cma_rank_mu(Y[512:].contiguous(), w[512:].contiguous())
    assert float((halves - dC).norm() / dC.norm()) < 3e-6      # same terms, different fp32 accumulation grouping on the tensor cores
