"""des_rollout_eval (closed-loop Pendulum-v0 on the device) through the C ABI against
oracle/pendulum_oracle.py and against the golden produced by the reference's verbatim train() on PendulumConfig.

Tolerances: the policy is evaluated in fp32 on the device and in fp64 by the oracle; an episode is 200 steps of a
feedback loop, so per-step differences of ~1e-7 grow along the trajectory.  Observed |dR|/|R| <= ~1e-5 on returns of
magnitude ~1e3; the bound used is 2e-4 (ranks may still flip between near-tied members: updates are compared through
the layered protocol, ranks taken from the device's fitness)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import nes_oracle as orc
from oracle import pendulum_oracle as po

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'train_closed_pend.npz')
RTOL = 2e-4


def dev(x, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(x)).to(dtype).cuda()


@pytest.mark.parametrize('H,n,reps,horizon', [(64, 24, 10, 200), (32, 9, 3, 50), (96, 5, 1, 120), (128, 6, 4, 200)])
def test_rollout_fitness_matches_oracle(H, n, reps, horizon):
    from distributedes_b200 import ops
    theta = orc.synthetic_theta(3, H, 1, seed=H)
    seed, gen, off = 21, 3, 5
    totals = torch.zeros(7, dtype=torch.float64, device='cuda')
    eps_out = torch.empty(n * reps, dtype=torch.float32, device='cuda')
    fit = ops.rollout_eval(dev(theta), hidden=H, horizon=horizon, repetitions=reps, sigma=0.1, clip=2.0, seed=seed,
                           generation=gen, member_offset=off, n_local=n, totals_out=totals, episodes_out=eps_out)
    ref, (osum, osq, cnt) = po.closed_fitness(theta, H, 0.1, seed, gen, off, n, reps, None, horizon)
    got = fit.cpu().numpy().astype(np.float64)
    assert np.max(np.abs(got - ref) / np.abs(ref)) < RTOL
    t = totals.cpu().numpy()
    assert t[6] == cnt == n * reps * horizon
    assert np.allclose(t[:3], osum, rtol=1e-4, atol=1e-3 * cnt ** 0.5) and np.allclose(t[3:6], osq, rtol=1e-4)
    # per-episode returns average to the fitness
    assert np.allclose(eps_out.cpu().numpy().reshape(n, reps).mean(1), got, rtol=1e-6)


def test_rollout_with_normaliser_statistics_and_test_episodes():
    from distributedes_b200 import ops
    H, n, reps = 64, 12, 10
    theta = orc.synthetic_theta(3, H, 1, seed=2)
    stats = (np.array([-0.2, 0.01, 0.3], np.float32), np.array([0.5, 0.4, 20.0], np.float32), np.float32(32000))
    st = dev(np.concatenate([stats[0], stats[1], [stats[2]]]))
    fit = ops.rollout_eval(dev(theta), hidden=H, repetitions=reps, sigma=0.1, clip=2.0, seed=4, generation=1,
                           member_offset=0, n_local=n, obs_stats=st)
    ref, _ = po.closed_fitness(theta, H, 0.1, 4, 1, 0, n, reps, stats)
    assert np.max(np.abs(fit.cpu().numpy() - ref) / np.abs(ref)) < RTOL
    # identity while n == 0 (utils.py:48-49)
    st0 = dev(np.concatenate([stats[0], stats[1], [0.0]]))
    fit0 = ops.rollout_eval(dev(theta), hidden=H, repetitions=reps, sigma=0.1, clip=2.0, seed=4, generation=1,
                            member_offset=0, n_local=n, obs_stats=st0)
    fit_none = ops.rollout_eval(dev(theta), hidden=H, repetitions=reps, sigma=0.1, clip=2.0, seed=4, generation=1,
                                member_offset=0, n_local=n)
    assert torch.equal(fit0, fit_none) and not torch.equal(fit0, fit)
    # test() episodes: unperturbed theta, the test reset stream
    ep = torch.empty(reps, dtype=torch.float32, device='cuda')
    ops.rollout_eval(dev(theta), hidden=H, repetitions=reps, sigma=0.1, clip=2.0, seed=4, generation=1, member_offset=0,
                     n_local=1, noiseless=True, obs_stats=st, episodes_out=ep)
    ref_t = po.test_returns(theta, H, 4, 1, reps, stats)
    assert np.max(np.abs(ep.cpu().numpy() - ref_t) / np.abs(ref_t)) < RTOL


def test_rollout_is_shard_invariant_and_deterministic():
    """Members are addressed globally: evaluating [0,20) in one launch or as [0,7)+[7,20) gives identical bits."""
    from distributedes_b200 import ops
    theta = dev(orc.synthetic_theta(3, 64, 1, seed=9))
    kw = dict(hidden=64, repetitions=10, sigma=0.1, clip=2.0, seed=8, generation=2)
    whole = ops.rollout_eval(theta, member_offset=0, n_local=20, **kw)
    again = ops.rollout_eval(theta, member_offset=0, n_local=20, **kw)
    a = ops.rollout_eval(theta, member_offset=0, n_local=7, **kw)
    b = ops.rollout_eval(theta, member_offset=7, n_local=13, **kw)
    assert torch.equal(whole, again) and torch.equal(whole, torch.cat([a, b]))


def test_rollout_action_noise_matches_oracle():
    from distributedes_b200 import ops
    H, n, reps = 32, 6, 2
    theta = orc.synthetic_theta(3, H, 1, seed=1)
    fit = ops.rollout_eval(dev(theta), hidden=H, horizon=60, repetitions=reps, sigma=0.1, clip=2.0, action_noise_std=0.3,
                           seed=17, generation=0, member_offset=2, n_local=n)
    eps = orc.noise(17, 0, 2, n, orc.param_count(3, H, 1))
    ret, _, _, _ = po.rollouts(orc.perturb(theta, 0.1, eps), H, 17, 0, np.arange(2, 2 + n), reps, None, 60, 2.0, 0.3)
    ref = ret.mean(1)
    assert np.max(np.abs(fit.cpu().numpy() - ref) / np.abs(ref)) < 5e-4      # MUFU normals (2^-21 abs) feed the loop


def test_rollout_rejects_bad_arguments():
    from distributedes_b200 import ops
    theta = dev(orc.synthetic_theta(3, 64, 1))
    with pytest.raises(RuntimeError, match='multiple of 32'):
        ops.rollout_eval(dev(orc.synthetic_theta(3, 48, 1)), hidden=48, sigma=0.1, clip=2.0, seed=0, n_local=2)
    with pytest.raises(RuntimeError, match='repetitions'):
        ops.rollout_eval(theta, hidden=64, repetitions=11, sigma=0.1, clip=2.0, seed=0, n_local=2)
    with pytest.raises(RuntimeError, match='unknown environment'):
        ops.rollout_eval(theta, env=5, hidden=64, sigma=0.1, clip=2.0, seed=0, n_local=2)
    with pytest.raises(RuntimeError, match='workspace'):
        ops.rollout_eval(theta, hidden=64, sigma=0.1, clip=2.0, seed=0, n_local=4,
                         totals_out=torch.zeros(7, dtype=torch.float64, device='cuda'),
                         workspace=torch.empty(3, dtype=torch.float64, device='cuda'))


def test_train_on_closed_loop_pendulum_matches_reference_golden():
    """natural_es.train(ClosedLoopPendulumConfig) = BASELINE configs[0] on the device, against the reference's own
    train() on PendulumConfig (golden): test rewards, normaliser statistics, gradient (layered on the device's
    fitness when ranks flip), parameters."""
    from distributedes_b200 import natural_es
    from distributedes_b200.config import ClosedLoopPendulumConfig
    g = np.load(GOLD)
    H, N, reps, seed, gens = int(g['H']), int(g['N']), int(g['reps']), int(g['seed']), int(g['gens'])
    cfg = ClosedLoopPendulumConfig(hidden_size=H)
    cfg.initial_weight = g['theta0'].copy()
    cfg.pop_size, cfg.sigma, cfg.learning_rate, cfg.seed = N, float(g['sigma']), float(g['lr']), seed
    cfg.repetitions = cfg.test_repetitions = reps
    cfg.max_steps = (gens + 1) * N * reps * 200 - 1
    eng = natural_es.build_engine(cfg)
    fits, stats = [], []
    real_rank, real_apply = eng.rank_and_reduce, eng.apply

    def spy_rank():
        fits.append(eng.fitness_all.cpu().numpy().astype(np.float64))
        return real_rank()

    def spy_apply():
        real_apply()
        stats.append(eng.obs_stats.cpu().numpy().copy())
    eng.rank_and_reduce, eng.apply = spy_rank, spy_apply
    rewards, steps, _ = natural_es.train(cfg, engine=eng)
    assert steps == list(g['train_steps'])
    assert np.allclose(rewards, g['test_rewards'], rtol=RTOL)
    theta, opt, P = g['theta0'].copy(), orc.Adam(), g['theta0'].size
    for gen in range(gens):
        assert np.allclose(stats[gen], g['stats'][gen], rtol=5e-4, atol=5e-5)
        s = orc.fitness_shift(fits[gen])
        grad = orc.nes_gradient(orc.noise(seed, gen, 0, N, P), s, float(g['sigma']))
        theta, _ = orc.nes_update(theta, grad, opt, float(g['wd']), float(g['lr']))
    # parameters after `gens` generations: chain on the device's own fitness (layered), then against the golden
    assert np.max(np.abs(eng.theta_numpy() - theta)) <= 1e-5 * np.max(np.abs(theta - g['theta0']))
    if np.max(np.abs(theta - g['theta'][-1])) <= 2e-6:        # no rank flip happened: equals the reference end to end
        assert np.max(np.abs(eng.theta_numpy() - g['theta'][-1])) <= 1e-5 * np.max(np.abs(g['theta'][-1] - g['theta0']))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_gpu_closed_loop_equals_one_gpu(tmp_path):
    import subprocess
    import sys
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'mp_rollout_worker.py')
    out = str(tmp_path)
    subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
                    '127.0.0.1', '--master-port', '29741', script, out], check=True, timeout=300)
    r0, r1 = np.load(os.path.join(out, 'rank0.npz')), np.load(os.path.join(out, 'rank1.npz'))
    for k in ('theta', 'stats', 'fit'):
        assert np.array_equal(r0[k], r1[k]), k
    from distributedes_b200.engine import RolloutEngine
    eng = RolloutEngine(hidden=64, pop_size=37, theta0=orc.synthetic_theta(3, 64, 1), sigma=0.1, learning_rate=0.1, seed=3)
    eng.generation()
    assert np.array_equal(eng.fitness_all.cpu().numpy(), r0['fit0'])
    assert np.allclose(eng.obs_stats.cpu().numpy(), r0['stats0'], rtol=1e-6)
