#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own code (imported from /root/reference).

TEST INFRASTRUCTURE.  Runs only in the build container (where /root/reference exists); the
fixtures it writes are committed so the GPU box — which has no /root/reference — can check the
oracle and the CUDA path against the reference's outputs.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

What comes from where:
  * fitness_shift, Adam                  -> /root/reference/utils.py:142-166, called directly
  * StandardFCNet forward / flat codec   -> /root/reference/model.py:7-39, called directly
  * per-member fitness                   -> /root/reference/utils.py:108-139 Evaluator.eval over the
                                            stub gym tape env (oracle/gym_stub)
  * closed-loop Pendulum generations     -> natural_es.train() VERBATIM on the reference's PendulumConfig over the stub
                                            gym's restated Pendulum-v0 (golden_train_closed), normaliser on
  * one..three full generations          -> /root/reference/natural_es.py:34-99 train() run VERBATIM
                                            (1 worker; np.random.randn replaced by the Philox noise so
                                            member identity is reproducible; config.opt replaced by a
                                            recording subclass of the reference Adam; once with
                                            SharedStats.merge disabled = observation normaliser off,
                                            SURVEY §8d, and once with the normaliser left on)
The only non-reference ingredient is the noise stream (oracle.nes_oracle.noise): the reference has
no reproducible RNG (natural_es.py:23 seeds from OS entropy).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = '/root/reference'
sys.path.insert(0, os.path.join(HERE, 'gym_stub'))
sys.path.insert(0, REF)
sys.path.insert(0, REPO)

import numpy as np
import torch

torch.set_num_threads(1)

import utils as ref_utils            # noqa: E402  /root/reference/utils.py
import model as ref_model            # noqa: E402  /root/reference/model.py
import config as ref_config          # noqa: E402  /root/reference/config.py
import natural_es as ref_nes         # noqa: E402  /root/reference/natural_es.py
from oracle import nes_oracle as orc  # noqa: E402

OUT = os.path.join(REPO, 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)
ref_utils.logger.setLevel('WARNING')


def golden_fitness_shift():
    rs = np.random.RandomState(7)
    out = {}
    for i, n in enumerate([2, 3, 16, 257, 4096]):
        x = rs.randn(n).astype(np.float32)
        assert len(np.unique(x)) == n          # tie-free: the reference's argsort is unstable on ties
        out['x%d' % i] = x
        out['y%d' % i] = ref_utils.fitness_shift(x)
    # list input, as natural_es.py:90 passes a python list
    out['x5'] = np.asarray([3.0, -1.0, 2.5, 0.0, 10.0], dtype=np.float32)
    out['y5'] = ref_utils.fitness_shift([3.0, -1.0, 2.5, 0.0, 10.0])
    np.savez(os.path.join(OUT, 'fitness_shift.npz'), **out)


def golden_adam():
    rs = np.random.RandomState(11)
    P, steps = 37, 6
    g = rs.randn(steps, P) * np.logspace(-3, 1, P)[None, :]
    opt = ref_utils.Adam()
    outs = np.stack([opt.update(g[t]) for t in range(steps)])
    np.savez(os.path.join(OUT, 'adam.npz'), g=g, step=outs, m=opt.m, v=opt.v,
             beta1_t=opt.beta1_t, beta2_t=opt.beta2_t)


def golden_forward():
    out = {}
    rs = np.random.RandomState(3)
    for tag, (d0, H, A, T) in {'pend': (3, 64, 1, 8), 'b64': (24, 64, 4, 8), 'b256': (24, 256, 4, 4)}.items():
        net = ref_model.StandardFCNet(d0, A, H)
        P = orc.param_count(d0, H, A)
        flat = (rs.randn(P) * 0.2).astype(np.float32)
        net.set_weight(flat.astype(np.float64))          # fp64 in, stored fp32 (model.py:23)
        assert np.array_equal(net.get_weight(), flat)    # codec round trip
        obs = rs.randn(T, d0).astype(np.float32)
        act = net(obs).data.numpy()
        out[tag + '_dims'] = np.asarray([d0, H, A, T])
        out[tag + '_flat'] = flat
        out[tag + '_obs'] = obs
        out[tag + '_act'] = act
        # named-parameter view, to pin the flat layout independently of our unflatten()
        out[tag + '_fc1w'] = net.fc1.weight.data.numpy()
        out[tag + '_fc1b'] = net.fc1.bias.data.numpy()
        out[tag + '_fc2w'] = net.fc2.weight.data.numpy()
        out[tag + '_fc2b'] = net.fc2.bias.data.numpy()
        out[tag + '_fc3w'] = net.fc3.weight.data.numpy()
        out[tag + '_fc3b'] = net.fc3.bias.data.numpy()
    np.savez(os.path.join(OUT, 'forward.npz'), **out)


class TapeConfig(ref_config.BasicConfig):
    def __init__(self, d0, A, T, hidden, clip):
        self.task = 'SynthTape-d%d-a%d-T%d-v0' % (d0, A, T)
        self.action_clip = lambda a: np.clip(a, -clip, clip)
        self.target = 10000
        torch.manual_seed(0)
        ref_config.BasicConfig.__init__(self, hidden)


class RecordingAdam(ref_utils.Adam):
    def __init__(self):
        ref_utils.Adam.__init__(self)
        self.rec_g, self.rec_step = [], []

    def update(self, g):
        self.rec_g.append(np.array(g, dtype=np.float64))
        step = ref_utils.Adam.update(self, g)
        self.rec_step.append(np.array(step, dtype=np.float64))
        return step


def golden_eval(tag, d0, H, A, T, clip, N, seed, sigma):
    """Per-member fitness from the reference Evaluator (utils.py:116-124), noise from the oracle."""
    cfg = TapeConfig(d0, A, T, H, clip)
    cfg.repetitions = 1
    norm = ref_utils.StaticNormalizer(cfg.state_dim)     # offline n == 0 -> identity (utils.py:48-49)
    ev = ref_utils.Evaluator(cfg, norm)
    theta = cfg.initial_weight.astype(np.float32)
    P = len(theta)
    eps = orc.noise(seed, 0, 0, N, P)
    fit = np.empty(N)
    steps = np.empty(N, dtype=np.int64)
    for i in range(N):
        disturbed = np.copy(theta)                        # natural_es.py:28
        disturbed += sigma * eps[i]                       # natural_es.py:30
        cost, st = ev.eval(disturbed)                     # natural_es.py:31
        fit[i] = -cost                                    # natural_es.py:32
        steps[i] = st
    np.savez(os.path.join(OUT, 'eval_%s.npz' % tag), dims=np.asarray([d0, H, A, T]), clip=clip, N=N,
             seed=seed, sigma=sigma, theta=theta, fitness=fit, steps=steps)


def golden_train_verbatim(tag, d0, H, A, T, clip, N, seed, sigma, lr, gens, normalizer=False):
    """natural_es.train() verbatim for `gens` generations (see module docstring for the three hooks)."""
    cfg = TapeConfig(d0, A, T, H, clip)
    cfg.repetitions = 1
    cfg.test_repetitions = 2
    cfg.num_workers = 1
    cfg.pop_size = N
    cfg.sigma = sigma
    cfg.learning_rate = lr
    cfg.opt = RecordingAdam()
    # train() checks `total_steps > max_steps` after collecting each generation and breaks BEFORE the
    # update (natural_es.py:82-84): gens+1 collections give `gens` updates.
    cfg.max_steps = (gens + 1) * N * T - 1
    P = len(cfg.initial_weight)
    theta0 = cfg.initial_weight.astype(np.float32)

    counter = {'k': 0}
    real_randn = np.random.randn

    def philox_randn(*shape):
        n = shape[0]
        if n == P:                                        # natural_es.py:29
            g, member = divmod(counter['k'], N)
            counter['k'] += 1
            return orc.noise(seed, g, member, 1, P)[0]
        return np.zeros(n)                                # utils.py:133, multiplied by action_noise_std=0

    real_merge = ref_utils.SharedStats.merge
    np.random.randn = philox_randn
    if not normalizer:
        ref_utils.SharedStats.merge = lambda self, B: None    # observation normaliser off
    try:
        rewards, steps, _ = ref_nes.train(cfg)
    finally:
        np.random.randn = real_randn
        ref_utils.SharedStats.merge = real_merge
    assert len(cfg.opt.rec_g) == gens, (len(cfg.opt.rec_g), gens)
    # replay natural_es.py:95-96 with torch to obtain theta after each generation
    param = torch.FloatTensor(torch.from_numpy(theta0.copy()))
    thetas, updates = [], []
    for st in cfg.opt.rec_step:
        gradient = torch.FloatTensor(st)                                  # :95
        upd = cfg.learning_rate * gradient
        param.add_(upd)                                                   # :96
        updates.append(upd.numpy().copy())
        thetas.append(param.numpy().copy())
    np.savez(os.path.join(OUT, 'train_%s%s.npz' % ('norm_' if normalizer else '', tag)), dims=np.asarray([d0, H, A, T]), clip=clip, N=N,
             seed=seed, sigma=sigma, lr=lr, wd=cfg.weight_decay, gens=gens, theta0=theta0,
             grad_after_wd=np.stack(cfg.opt.rec_g), adam_step=np.stack(cfg.opt.rec_step),
             update=np.stack(updates), theta=np.stack(thetas),
             test_rewards=np.asarray(rewards, dtype=np.float64), train_steps=np.asarray(steps))


def golden_train_closed(tag, H, N, reps, seed, sigma, lr, gens):
    """natural_es.train() verbatim on the reference's own PendulumConfig (config.py:26-31) over the stub gym's restated
    Pendulum-v0, observation normaliser ON (the reference's real workload).  Hooks: Philox noise for np.random.randn
    (as above), recording Adam, and the stub's reset hook so episode k of the single worker starts from the counter-RNG
    state of (generation, member, repetition) and the master's test() episodes from the test stream."""
    import gym
    from oracle import pendulum_oracle as po
    torch.manual_seed(0)
    first = gym._pendulum_instances[0]            # instance numbers: first = config probe, +1 = worker, +2+g = test(g)
    cfg = ref_config.PendulumConfig(hidden_size=H)
    cfg.repetitions = reps
    cfg.test_repetitions = reps
    cfg.num_workers = 1
    cfg.pop_size = N
    cfg.sigma = sigma
    cfg.learning_rate = lr
    cfg.opt = RecordingAdam()
    cfg.max_steps = (gens + 1) * N * reps * po.HORIZON - 1
    P = len(cfg.initial_weight)
    theta0 = cfg.initial_weight.astype(np.float32)
    counter = {'k': 0}
    real_randn = np.random.randn

    def philox_randn(*shape):
        n = shape[0]
        if n == P:
            g, member = divmod(counter['k'], N)
            counter['k'] += 1
            return orc.noise(seed, g, member, 1, P)[0]
        return np.zeros(n)

    def reset_hook(instance, episode):
        if instance == first + 1:                                        # the worker's environment
            g, rest = divmod(episode, N * reps)
            member, rep = divmod(rest, reps)
        else:                                                            # a test() environment
            g, member, rep = instance - (first + 2), po.TEST_MEMBER, episode
        th, thd = po.reset_states(seed, g, [member], reps)
        return th[0, rep], thd[0, rep]

    stats_log = []
    real_merge = ref_utils.SharedStats.merge

    def logging_merge(self, B):
        real_merge(self, B)
        stats_log.append(np.concatenate([self.m.numpy(), self.v.numpy(), self.n.numpy()]).copy())

    np.random.randn = philox_randn
    gym.pendulum_reset_hook = reset_hook
    ref_utils.SharedStats.merge = logging_merge
    try:
        rewards, steps, _ = ref_nes.train(cfg)
    finally:
        np.random.randn = real_randn
        gym.pendulum_reset_hook = None
        ref_utils.SharedStats.merge = real_merge
    assert len(cfg.opt.rec_g) == gens
    param = torch.FloatTensor(torch.from_numpy(theta0.copy()))
    thetas = []
    for st in cfg.opt.rec_step:
        param.add_(cfg.learning_rate * torch.FloatTensor(st))
        thetas.append(param.numpy().copy())
    np.savez(os.path.join(OUT, 'train_closed_%s.npz' % tag), H=H, N=N, reps=reps, seed=seed, sigma=sigma, lr=lr,
             wd=cfg.weight_decay, gens=gens, theta0=theta0, grad_after_wd=np.stack(cfg.opt.rec_g),
             adam_step=np.stack(cfg.opt.rec_step), theta=np.stack(thetas), stats=np.stack(stats_log),
             test_rewards=np.asarray(rewards, dtype=np.float64), train_steps=np.asarray(steps))


if __name__ == '__main__':
    golden_fitness_shift()
    golden_adam()
    golden_forward()
    golden_eval('pend', 3, 64, 1, 32, 2.0, 16, seed=5, sigma=0.1)
    golden_eval('b64', 24, 64, 4, 16, 1.0, 24, seed=6, sigma=0.1)
    golden_train_verbatim('pend', 3, 64, 1, 32, 2.0, 16, seed=5, sigma=0.1, lr=0.1, gens=3)
    golden_train_verbatim('b64', 24, 64, 4, 16, 1.0, 24, seed=6, sigma=0.1, lr=0.1, gens=3)
    # the same with the reference's observation normaliser left ON (SharedStats.merge untouched)
    golden_train_verbatim('pend', 3, 64, 1, 32, 2.0, 16, seed=5, sigma=0.1, lr=0.1, gens=3, normalizer=True)
    golden_train_verbatim('b64', 24, 64, 4, 16, 1.0, 24, seed=6, sigma=0.1, lr=0.1, gens=3, normalizer=True)
    # BASELINE configs[0]: Pendulum-v0, 2x64 MLP, population 16, 10 repetitions of 200 steps, closed loop
    golden_train_closed('pend', 64, 16, 10, seed=7, sigma=0.1, lr=0.1, gens=2)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
