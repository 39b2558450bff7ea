"""Minimal stand-in for the ``gym`` package so the reference's config.py / natural_es.py import and run
verbatim in this container (``gym`` is not installed; config.py:1 imports it).  TEST INFRASTRUCTURE ONLY —
used by oracle/make_golden.py and oracle/ref_cpu_baseline.py, never by the product package.

``gym.make('SynthTape-d<d0>-a<A>-T<T>-v0')`` returns the synthetic observation-tape environment of
SURVEY.md §8d: a fixed tape X[T,d0], targets a*[T,A]; reward r_t = -||a_t - a*_t||^2 for the (already
clipped, utils.py:134) action the agent passes; the episode ends after T steps.
"""
import re
import numpy as np


class _Box:
    def __init__(self, shape):
        self.shape = shape


class SynthTapeEnv:
    def __init__(self, d0, A, T, seed=1234):
        rs = np.random.RandomState(seed)
        self.obs = rs.randn(T, d0).astype(np.float32)
        self.target = np.tanh(rs.randn(T, A)).astype(np.float32)
        self.T = T
        self.observation_space = _Box((d0,))
        self.action_space = _Box((A,))
        self.t = 0

    def reset(self):
        self.t = 0
        return self.obs[0]

    def step(self, action):
        a = np.asarray(action, dtype=np.float64).reshape(-1)
        d = a - self.target[self.t].astype(np.float64)
        reward = -float(np.dot(d, d))
        self.t += 1
        done = self.t >= self.T
        obs = self.obs[self.t] if not done else np.zeros_like(self.obs[0])
        return obs, reward, done, {}


def make(task):
    m = re.fullmatch(r'SynthTape-d(\d+)-a(\d+)-T(\d+)-v0', task)
    if m is None:
        raise ValueError('gym stub only knows SynthTape-d<d0>-a<A>-T<T>-v0, got %r' % (task,))
    return SynthTapeEnv(int(m.group(1)), int(m.group(2)), int(m.group(3)))
