"""The oracle (oracle/nes_oracle.py) against fixtures produced by the reference's own code
(oracle/make_golden.py, run where /root/reference exists).  CPU only."""
import os

import numpy as np
import pytest

from oracle import nes_oracle as orc


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32 at 7 rounds (the noise contract, include/des_b200.h) and at 10 rounds
    ctrs = [((0, 0, 0, 0), (0, 0)), ((0xffffffff,) * 4, (0xffffffff,) * 2),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0))]
    kat = {7: [(0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48), (0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662),
               (0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a)],
           10: [(0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd),
                (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)]}
    assert orc.PHILOX_ROUNDS == 7
    for rounds, wants in kat.items():
        for (ctr, key), want in zip(ctrs, wants):
            got = orc.philox4x32(*[np.asarray([c]) for c in ctr], *key, rounds=rounds)
            assert tuple(int(g[0]) for g in got) == want
    got = orc.philox4x32(*[np.asarray([c]) for c in ctrs[2][0]], *ctrs[2][1])        # the default is the contract
    assert tuple(int(g[0]) for g in got) == kat[7][2]


def test_noise_is_standard_normal_and_counter_based():
    eps = orc.noise(seed=42, gen=3, member_offset=0, n_members=64, P=4481)
    assert eps.shape == (64, 4481)
    assert abs(eps.mean()) < 0.01 and abs(eps.std() - 1) < 0.01
    # pure function of (seed, gen, member, j): a shard regenerates exactly its rows
    sub = orc.noise(seed=42, gen=3, member_offset=17, n_members=5, P=4481)
    assert np.array_equal(sub, eps[17:22])
    assert not np.array_equal(orc.noise(42, 4, 0, 1, 4481), eps[:1])
    assert not np.array_equal(orc.noise(43, 3, 0, 1, 4481), eps[:1])
    # P not a multiple of 4: prefix property
    assert np.array_equal(orc.noise(42, 3, 0, 2, 4479), eps[:2, :4479])


def test_unit_uniform_edges():
    f = orc.u32_to_one_two(np.asarray([0, 1, 0x7FFFFF, 0x800000, 0xFFFFFFFF], dtype=np.uint32))
    assert f[0] == 1.0 and f[1] == 1.0 + 2.0 ** -23 and f[2] == 2.0 - 2.0 ** -23
    assert f[3] == 1.0 and f[4] == f[2]                       # only the low 23 bits count
    z0, z1 = orc.box_muller(np.asarray([0, 0x7FFFFF], dtype=np.uint32), np.asarray([0, 0x400000], dtype=np.uint32))
    assert np.all(np.isfinite(z0)) and np.all(np.isfinite(z1))
    assert abs(np.hypot(z0[0], z1[0]) - np.sqrt(2 * 24 * np.log(2))) < 1e-12      # u1 = 2^-24: the 5.77 sigma tail
    assert np.hypot(z0[1], z1[1]) < 4e-4                                          # u1 = 1 - 2^-24


def test_fitness_shift_matches_reference(golden_dir):
    g = load(golden_dir, 'fitness_shift.npz')
    for i in range(6):
        x, y = g['x%d' % i], g['y%d' % i]
        assert np.array_equal(orc.fitness_shift(x), y)        # fp64, bit exact
    # pinned behaviours the reference leaves implicit
    s = orc.fitness_shift([1, 1, 1, 0])                       # ties -> by index
    assert np.allclose(s, [-1 / 6, 1 / 6, 1 / 2, -1 / 2])
    s = orc.fitness_shift(np.arange(5.0))
    assert s[0] == -0.5 and s[-1] == 0.5 and abs(s.sum()) < 1e-15


def test_adam_matches_reference(golden_dir):
    g = load(golden_dir, 'adam.npz')
    opt = orc.Adam()
    for t in range(len(g['g'])):
        assert np.array_equal(opt.update(g['g'][t]), g['step'][t])
    assert np.array_equal(opt.m, g['m']) and np.array_equal(opt.v, g['v'])
    # first step is ~sign(g)
    first = orc.Adam().update(g['g'][0])
    assert np.allclose(first, np.sign(g["g"][0]), atol=1e-3)


@pytest.mark.parametrize('tag', ['pend', 'b64', 'b256'])
def test_forward_and_flat_layout_match_reference(golden_dir, tag):
    g = load(golden_dir, 'forward.npz')
    d0, H, A, T = (int(v) for v in g[tag + '_dims'])
    flat = g[tag + '_flat']
    W1, b1, W2, b2, W3, b3 = orc.unflatten(flat, d0, H, A)
    for ours, name in ((W1, 'fc1w'), (b1, 'fc1b'), (W2, 'fc2w'), (b2, 'fc2b'), (W3, 'fc3w'), (b3, 'fc3b')):
        assert np.array_equal(ours, g[tag + '_' + name])
    act = orc.forward(flat, g[tag + '_obs'], d0, H, A)
    ref = g[tag + '_act'].astype(np.float64)                  # reference forward is fp32 torch
    assert np.max(np.abs(act - ref)) <= 2e-6 * max(1.0, np.max(np.abs(ref)))


@pytest.mark.parametrize('tag', ['pend', 'b64'])
def test_member_fitness_matches_reference_evaluator(golden_dir, tag):
    g = load(golden_dir, 'eval_%s.npz' % tag)
    d0, H, A, T = (int(v) for v in g['dims'])
    obs, target = orc.synthetic_tape(T, d0, A)
    fit = orc.evaluate_population(g['theta'], obs, target, float(g['sigma']), float(g['clip']),
                                  int(g['seed']), 0, 0, int(g['N']), d0, H, A)
    assert np.all(g['steps'] == T)
    ref = g['fitness']
    assert np.max(np.abs(fit - ref) / np.abs(ref)) < 2e-6     # reference forward is fp32
    assert np.array_equal(orc.ranks_stable(fit), orc.ranks_stable(ref))


@pytest.mark.parametrize('tag', ['pend', 'b64'])
def test_generations_match_reference_train_verbatim(golden_dir, tag):
    """natural_es.train() run verbatim for 3 generations vs the oracle's nes_generation chain."""
    g = load(golden_dir, 'train_%s.npz' % tag)
    d0, H, A, T = (int(v) for v in g['dims'])
    N, seed, sigma, lr, wd, clip = int(g['N']), int(g['seed']), float(g['sigma']), float(g['lr']), \
        float(g['wd']), float(g['clip'])
    obs, target = orc.synthetic_tape(T, d0, A)
    theta = g['theta0']
    opt = orc.Adam()
    # test() at the top of each loop iteration (natural_es.py:54) = noiseless fitness of theta_g
    assert abs(orc.tape_fitness(orc.forward(theta, obs, d0, H, A), target, clip) - g['test_rewards'][0]) \
        < 2e-6 * abs(g['test_rewards'][0])
    for gen in range(int(g['gens'])):
        out = orc.nes_generation(theta, opt, obs, target, sigma=sigma, clip=clip, seed=seed, gen=gen, N=N,
                                 d0=d0, H=H, A=A, weight_decay=wd, learning_rate=lr)
        ref_g = g['grad_after_wd'][gen]
        ours_g = out['gradient'] * (1 - wd)
        assert np.linalg.norm(ours_g - ref_g) <= 1e-9 * np.linalg.norm(ref_g), gen
        assert np.linalg.norm(out['update'] - g['update'][gen]) <= 1e-6 * np.linalg.norm(g['update'][gen])
        assert np.max(np.abs(out['theta'] - g['theta'][gen])) <= 1e-6
        theta = out['theta']
        rew = orc.tape_fitness(orc.forward(theta, obs, d0, H, A), target, clip)
        assert abs(rew - g['test_rewards'][gen + 1]) < 5e-6 * abs(g['test_rewards'][gen + 1])


def _normalised_chain(g, exact_feed):
    """natural_es.train() with the reference's observation normaliser left ON: obs of generation g are normalised
    with the statistics merged from generations < g (utils.py:48-51, natural_es.py:85-89)."""
    d0, H, A, T = (int(v) for v in g['dims'])
    N, seed, sigma, lr, wd, clip = int(g['N']), int(g['seed']), float(g['sigma']), float(g['lr']), float(g['wd']), \
        float(g['clip'])
    obs, target = orc.synthetic_tape(T, d0, A)
    theta, opt, stats = g['theta0'], orc.Adam(), orc.ObsStats(d0)
    errs = []
    for gen in range(int(g['gens'])):
        obs_n = np.stack([stats.normalize(o) for o in obs])
        rew = orc.tape_fitness(orc.forward(theta, obs_n, d0, H, A), target, clip)        # test(), natural_es.py:54
        assert abs(rew - g['test_rewards'][gen]) < 5e-6 * abs(g['test_rewards'][gen]), gen
        out = orc.nes_generation(theta, opt, obs_n, target, sigma=sigma, clip=clip, seed=seed, gen=gen, N=N, d0=d0, H=H,
                                 A=A, weight_decay=wd, learning_rate=lr)
        errs.append(np.linalg.norm(out['gradient'] * (1 - wd) - g['grad_after_wd'][gen]) / np.linalg.norm(g['grad_after_wd'][gen]))
        theta = out['theta']
        if exact_feed:                       # the single worker feeds every raw observation of every member, in order
            online = orc.ObsStats(d0)
            for _ in range(N):
                for o in obs:
                    online.feed(o)
            stats.merge(online)
        else:
            stats.merge_tape(obs, N * T)
    return errs, theta


@pytest.mark.parametrize('tag', ['pend', 'b64'])
def test_generations_with_reference_normaliser_on(golden_dir, tag):
    g = load(golden_dir, 'train_norm_%s.npz' % tag)
    errs, theta = _normalised_chain(g, exact_feed=True)
    assert max(errs) <= 1e-6, errs          # sequential fp32 Welford replicated: only fp32-forward noise remains
    assert np.max(np.abs(theta - g['theta'][-1])) <= 2e-6
    # closed form used on the GPU (tape mean / variance, fp64): same result to the fp32 noise of the online update
    errs2, theta2 = _normalised_chain(g, exact_feed=False)
    assert max(errs2) <= 1e-4, errs2
    # and the normaliser matters: the no-normaliser fixture differs from generation 1 on
    g0 = load(golden_dir, 'train_%s.npz' % tag)
    assert np.allclose(g['grad_after_wd'][0], g0['grad_after_wd'][0])
    assert not np.allclose(g['grad_after_wd'][1], g0['grad_after_wd'][1], rtol=1e-3)


def test_cma_restatement_is_a_working_cma_es():
    """The CMA oracle is unpinned (no pycma, no fixtures in the reference): at least it must behave like CMA-ES —
    weights/constants as in the tutorial's Table 1 and monotone progress on the sphere."""
    from oracle import cma_oracle as cma
    k = cma.cma_constants(1024, 256)
    assert k['mu'] == 128 and abs(k['w'].sum() - 1) < 1e-12 and np.all(np.diff(k['w'][:128]) < 0)
    assert 60 < k['mu_eff'] < 80 and 0 < k['c1'] < k['cmu'] < 1 and k['c1'] + k['cmu'] <= 1
    rs = np.random.RandomState(0)
    st = cma.CMAState(rs.randn(32), 1.0, 16)
    f0 = cma.sphere(st.m[None])[0]
    for _ in range(60):
        X = st.ask(rs.randn(16, 32))
        st.tell(X, cma.sphere(X))
        assert np.allclose(st.C, st.C.T) and np.all(st.D > 0)
    assert cma.sphere(st.m[None])[0] < 0.05 * f0
