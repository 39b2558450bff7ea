"""The CUDA path against the REFERENCE's own outputs (tests/golden/*.npz, written by oracle/make_golden.py from
/root/reference) — directly, not through the oracle: Evaluator.eval fitness (utils.py:116-139) and natural_es.train()
run verbatim with the observation normaliser off (natural_es.py:34-99)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import nes_oracle as orc          # tape generator only (synthetic_tape): same RandomState stream as the stub env

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def relnorm(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize('tag', ['pend', 'b64'])
def test_fitness_matches_reference_evaluator(golden_dir, tag):
    """des_nes_eval (fp32 path) == what the reference's Evaluator.eval returned for the same perturbed members: the
    fixture holds -cost of natural_es.py:31-32 for members 0..N-1 of generation 0."""
    from distributedes_b200 import ops
    g = np.load(os.path.join(golden_dir, 'eval_%s.npz' % tag))
    d0, H, A, T = (int(v) for v in g['dims'])
    N, seed, sigma, clip = int(g['N']), int(g['seed']), float(g['sigma']), float(g['clip'])
    obs, target = orc.synthetic_tape(T, d0, A)
    got = ops.nes_eval(torch.from_numpy(g['theta']).to(DEV), torch.from_numpy(obs).to(DEV), torch.from_numpy(target).to(DEV),
                       hidden=H, sigma=sigma, clip=clip, seed=seed, generation=0, member_offset=0, n_local=N,
                       precision='fp32').cpu().numpy().astype(np.float64)
    # the reference runs the forward in fp32 torch on weights perturbed in fp64->fp32; ours regenerates eps with MUFU
    # approximations (|d eps| <= 4e-6): 2e-5 relative on the fitness (measured ~2e-6)
    assert np.max(np.abs(got - g['fitness']) / np.abs(g['fitness'])) < 2e-5
    assert int(g['steps'][0]) == T


@pytest.mark.parametrize('tag', ['pend', 'b64'])
def test_generations_match_reference_train(golden_dir, tag):
    """NESEngine against natural_es.train() run verbatim (observation normaliser off, tests/golden/train_*.npz):
    test rewards (natural_es.py:54), gradient after weight decay (:91-93), Adam step and parameters (:94-96), three
    generations, straight from theta0 — populations of 16 / 24 members, where no rank flips."""
    from distributedes_b200.engine import NESEngine
    g = np.load(os.path.join(golden_dir, 'train_%s.npz' % tag))
    d0, H, A, T = (int(v) for v in g['dims'])
    N, seed = int(g['N']), int(g['seed'])
    obs, target = orc.synthetic_tape(T, d0, A)
    eng = NESEngine(state_dim=d0, hidden=H, action_dim=A, pop_size=N, theta0=g['theta0'], obs=obs, target=target,
                    sigma=float(g['sigma']), learning_rate=float(g['lr']), weight_decay=float(g['wd']), clip=float(g['clip']),
                    seed=seed, precision='fp32', device=DEV)
    for gen in range(int(g['gens'])):
        rew = eng.noiseless_fitness()
        assert abs(rew - g['test_rewards'][gen]) < 2e-5 * abs(g['test_rewards'][gen])
        eng.generation()
        grad = eng.partial.cpu().numpy().astype(np.float64) / N / float(g['sigma']) * (1 - float(g['wd']))
        assert relnorm(grad, g['grad_after_wd'][gen]) <= 2e-5, gen
        if gen >= 1:            # from the second Adam step on the update is well conditioned (the first is ~sign(g))
            upd = eng.update.cpu().numpy()
            assert relnorm(upd, g['update'][gen]) <= 2e-5, gen
        assert np.max(np.abs(eng.theta_numpy() - g['theta'][gen])) <= 2e-5
    assert np.array_equal(np.asarray(g['train_steps'][:2]), [0, N * T])


def test_host_normaliser_surface_feeds_the_device_path():
    """utils.StaticNormalizer / SharedStats with NON-empty statistics: Evaluator.eval normalises the tape on the device
    with the offline statistics (utils.py:48-51,131) and accumulates the online ones (utils.py:68-73)."""
    from distributedes_b200.config import BipedalWalkerConfig
    from distributedes_b200.utils import Evaluator, SharedStats, StaticNormalizer
    cfg = BipedalWalkerConfig(hidden_size=64, tape_len=64)
    env = cfg.env_fn()
    norm = StaticNormalizer(cfg.state_dim)
    warm = SharedStats(cfg.state_dim)
    for o in env.obs[:40]:
        warm.feed(o)
    norm.offline_stats.load(warm)
    ev = Evaluator(cfg, norm)
    cost, steps = ev.eval(cfg.initial_weight)
    stats = orc.ObsStats(cfg.state_dim)
    stats.m, stats.v, stats.n = warm.m.copy(), warm.v.copy(), np.float32(warm.n[0])
    nobs = np.stack([stats.normalize(o) for o in env.obs]).astype(np.float32)
    ref = orc.tape_fitness(orc.forward(cfg.initial_weight, nobs, 24, 64, 4), env.target, 1.0)
    assert steps == 64 and abs(-cost - ref) < 5e-5 * abs(ref)
    assert norm.online_stats.n[0] == 64 * cfg.repetitions
    assert np.allclose(norm.online_stats.m, env.obs.mean(0), atol=1e-5)
