"""The synthetic observation-tape environment of SURVEY.md §8d (host-side description only).

The reference steps a gym env per member per step (utils.py:126-139).  For the benchmark workload the
env is a fixed tape: observations X[T, d0] shared by all members, targets a*[T, A], reward
r_t = -||clip(a_t) - a*_t||^2, return = sum_t r_t.  The arrays here are what gets copied to the GPU;
stepping happens inside des_nes_eval.
"""
from __future__ import annotations

import numpy as np


class _Box:
    def __init__(self, shape):
        self.shape = shape


class TapeEnv:
    def __init__(self, state_dim, action_dim, tape_len, seed=1234):
        rs = np.random.RandomState(seed)
        self.obs = rs.randn(tape_len, state_dim).astype(np.float32)
        self.target = np.tanh(rs.randn(tape_len, action_dim)).astype(np.float32)
        self.observation_space = _Box((state_dim,))
        self.action_space = _Box((action_dim,))
        self.tape_len = int(tape_len)


class DeviceEnv:
    """Shape descriptor of an environment that is stepped INSIDE des_rollout_eval (closed loop, per-member
    observations): there is no host-side step().  'Pendulum-v0': config.py:26-31."""
    SPECS = {'Pendulum-v0': dict(state_dim=3, action_dim=1, horizon=200, clip=2.0)}

    def __init__(self, task):
        if task not in self.SPECS:
            raise ValueError('no device environment %r (available: %s)' % (task, sorted(self.SPECS)))
        spec = self.SPECS[task]
        self.task = task
        self.observation_space = _Box((spec['state_dim'],))
        self.action_space = _Box((spec['action_dim'],))
        self.horizon, self.clip = spec['horizon'], spec['clip']

    def reset(self):
        raise RuntimeError('%s is stepped on the GPU by des_rollout_eval; it has no host-side reset/step' % self.task)

    step = reset
