"""Configuration surface of the reference (config.py:5-53): same attribute names, the tape env in place
of gym.  `sigma` and `learning_rate` are set by all_tasks() in the reference (natural_es.py:143-144);
they get those values as defaults here.  New attributes: tape_len (T), seed, precision, device.
"""
from __future__ import annotations

import numpy as np

from .envs import DeviceEnv, TapeEnv
from .model import StandardFCNet
from .utils import Adam


class BasicConfig:
    def __init__(self, hidden_size):
        if not hasattr(self, 'env_fn'):
            self.env_fn = lambda: TapeEnv(self.state_dim_, self.action_dim_, self.tape_len)
        self.repetitions = 1          # the tape is deterministic: one episode per evaluation (reference: 10)
        self.test_repetitions = 1
        env = self.env_fn()
        self.action_dim = env.action_space.shape[0]
        self.state_dim = env.observation_space.shape[0]
        self.hidden_size = hidden_size
        self.model_fn = lambda: StandardFCNet(self.state_dim, self.action_dim, self.hidden_size, seed=0)
        model = self.model_fn()
        self.initial_weight = model.get_weight()
        self.reward_to_fitness = lambda r: r
        self.pop_size = 30
        self.num_workers = 1          # = GPUs (one process per GPU); set by the launcher
        self.max_steps = 0
        self.opt = Adam()
        self.weight_decay = 0.005
        self.action_noise_std = 0
        self.tag = ''
        self.sigma = 0.1              # natural_es.py:143
        self.learning_rate = 0.1      # natural_es.py:144
        self.seed = 0
        self.precision = 'fp32'
        self.max_generations = 0      # 0 = unbounded (stop on max_steps like the reference)
        self.normalize_obs = False    # True = the reference's StaticNormalizer/SharedStats behaviour (utils.py:37-106)


class SynthTapeConfig(BasicConfig):
    def __init__(self, hidden_size=64, state_dim=24, action_dim=4, tape_len=256, clip=1.0):
        self.task = 'SynthTape-d%d-a%d-T%d-v0' % (state_dim, action_dim, tape_len)
        self.state_dim_, self.action_dim_, self.tape_len = state_dim, action_dim, tape_len
        self.clip = float(clip)
        self.action_clip = lambda a: np.clip(a, -clip, clip)
        self.target = 10000
        BasicConfig.__init__(self, hidden_size)


class PendulumConfig(SynthTapeConfig):
    """Pendulum-v0 shapes (config.py:26-31): obs 3, action 1, clip 2, 200-step episodes."""

    def __init__(self, hidden_size=16, tape_len=200):
        SynthTapeConfig.__init__(self, hidden_size, 3, 1, tape_len, 2.0)


class ClosedLoopPendulumConfig(BasicConfig):
    """The reference's PendulumConfig (config.py:26-31) with the environment stepped on the device: 10 repetitions of
    200-step episodes per member (config.py:8-9), per-member observations, observation normaliser on."""

    def __init__(self, hidden_size=64):
        # limits of des_rollout_eval (csrc/des_envs.cu): checked here, not at the first generation.  The reference's own
        # default hidden_size=16 (config.py:27) is below the kernel's 32-unit granularity.
        if hidden_size % 32 != 0 or not (32 <= hidden_size <= 128):
            raise ValueError('ClosedLoopPendulumConfig: hidden_size must be 32, 64, 96 or 128 on the device path (got %r)'
                             % (hidden_size,))
        self.task = 'Pendulum-v0'
        self.clip = 2.0
        self.action_clip = lambda a: np.clip(a, -2, 2)
        self.target = 10000
        self.tape_len = DeviceEnv.SPECS[self.task]['horizon']
        self.env_fn = lambda: DeviceEnv(self.task)
        BasicConfig.__init__(self, hidden_size)
        self.closed_loop = True
        self.repetitions = 10         # config.py:8
        self.test_repetitions = 10    # config.py:9
        self.normalize_obs = True


class BipedalWalkerConfig(SynthTapeConfig):
    """BipedalWalker-v2 shapes (config.py:34-39): obs 24, action 4, clip 1."""

    def __init__(self, hidden_size=16, tape_len=256):
        SynthTapeConfig.__init__(self, hidden_size, 24, 4, tape_len, 1.0)
