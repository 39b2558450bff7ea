"""Minimal stand-in for the ``gym`` package so the reference's config.py / natural_es.py import and run
verbatim in this container (``gym`` is not installed; config.py:1 imports it).  TEST INFRASTRUCTURE ONLY —
used by oracle/make_golden.py and oracle/ref_cpu_baseline.py, never by the product package.

``gym.make('Pendulum-v0')`` returns the restated Pendulum (see PendulumEnv).
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


class PendulumEnv:
    """'Pendulum-v0' of OpenAI gym (classic_control/pendulum.py + the 200-step TimeLimit wrapper), restated from its
    published dynamics: state (theta, theta_dot) in float64, torque clipped to +-2, speed to +-8, dt 0.05, g 10.

    ``reset_hook(instance_index, episode_index) -> (theta, theta_dot)`` (module attribute ``pendulum_reset_hook``) lets
    oracle/make_golden.py feed the counter-RNG reset states; without it reset() draws uniform(-[pi,1], [pi,1])."""
    max_speed, max_torque, dt, horizon = 8.0, 2.0, 0.05, 200

    def __init__(self, instance):
        self.instance = instance
        self.episode = 0
        self.np_random = np.random.RandomState(instance)
        self.observation_space = _Box((3,))
        self.action_space = _Box((1,))

    def _obs(self):
        th, thdot = self.state
        return np.array([np.cos(th), np.sin(th), thdot])

    def reset(self):
        if pendulum_reset_hook is not None:
            self.state = np.asarray(pendulum_reset_hook(self.instance, self.episode), dtype=np.float64)
        else:
            high = np.array([np.pi, 1.0])
            self.state = self.np_random.uniform(low=-high, high=high)
        self.episode += 1
        self.t = 0
        return self._obs()

    def step(self, u):
        th, thdot = self.state
        u = np.clip(u, -self.max_torque, self.max_torque)[0]
        norm = ((th + np.pi) % (2 * np.pi)) - np.pi
        costs = norm ** 2 + 0.1 * thdot ** 2 + 0.001 * (u ** 2)
        newthdot = thdot + (-3 * 10.0 / 2 * np.sin(th + np.pi) + 3.0 * u) * self.dt
        newth = th + newthdot * self.dt
        newthdot = np.clip(newthdot, -self.max_speed, self.max_speed)
        self.state = np.array([newth, newthdot])
        self.t += 1
        return self._obs(), -costs, self.t >= self.horizon, {}


pendulum_reset_hook = None
_pendulum_instances = [0]


def make(task):
    if task == 'Pendulum-v0':
        _pendulum_instances[0] += 1
        return PendulumEnv(_pendulum_instances[0] - 1)
    m = re.fullmatch(r'SynthTape-d(\d+)-a(\d+)-T(\d+)-v0', task)
    if m is None:
        raise ValueError('gym stub only knows Pendulum-v0 and SynthTape-d<d0>-a<A>-T<T>-v0, got %r' % (task,))
    return SynthTapeEnv(int(m.group(1)), int(m.group(2)), int(m.group(3)))
