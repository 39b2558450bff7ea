"""Host-side mirror of the reference's utils.py surface for the hot path: fitness_shift (utils.py:142-148),
Adam (utils.py:150-166), Evaluator (utils.py:108-139), StaticNormalizer/SharedStats (utils.py:37-106: these host
classes only carry the empty/identity state; the live statistics are device-resident in engine.NESEngine with
normalize_obs=True, kernels des_obs_stats_merge / des_obs_normalize).  All arithmetic runs in
libdes_b200.so on the GPU; these classes only adapt argument/return conventions.
"""
from __future__ import annotations

import logging

import numpy as np
import torch

logging.basicConfig(format='%(asctime)s - %(name)s - %(levelname)s: %(message)s')
logger = logging.getLogger('MAIN')
logger.setLevel(logging.DEBUG)


def _device(device=None):
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError('distributedes_b200 needs a CUDA device: there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def fitness_shift(x, device=None):
    """utils.py:142-148: centered ranks in [-0.5, 0.5].  Accepts a list / ndarray / tensor, returns a float64
    ndarray like the reference (values are the kernel's fp32 results; ties rank by index)."""
    from . import ops
    if isinstance(x, torch.Tensor) and x.is_cuda:
        f = x.detach().to(torch.float32).reshape(-1).contiguous()
    else:
        f = torch.as_tensor(np.asarray(x, dtype=np.float32).reshape(-1)).to(_device(device))
    return ops.centered_rank(f).cpu().numpy().astype(np.float64)


class Adam:
    """utils.py:150-166.  State (m, v as fp64, beta^t products) lives on the GPU; update(g) returns the
    bias-corrected step direction m_hat / (sqrt(v_hat) + epsilon) as the kernel computes it."""

    def __init__(self, beta1=0.9, beta2=0.999, epsilon=1e-08):
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon
        self._st = None

    def _ensure(self, P, device):
        from . import ops
        if self._st is None:
            self._m = torch.zeros(P, dtype=torch.float64, device=device)
            self._v = torch.zeros(P, dtype=torch.float64, device=device)
            self._st = ops.new_state(device, 0)
            self._zero = torch.zeros(P, dtype=torch.float32, device=device)
            self._out = torch.empty(P, dtype=torch.float32, device=device)

    def update(self, g, device=None):
        from . import ops
        g32 = torch.as_tensor(np.asarray(g, dtype=np.float32).reshape(-1)).to(_device(device))
        self._ensure(g32.numel(), g32.device)
        self._zero.zero_()
        # N=1, sigma=1, wd=0, lr=1: des_nes_apply reduces to the Adam step of utils.py:159-166
        ops.nes_apply(self._zero, self._m, self._v, g32, 1, self._st, sigma=1.0, learning_rate=1.0, weight_decay=0.0,
                      beta1=self.beta1, beta2=self.beta2, epsilon=self.epsilon, update_out=self._out)
        ops.state_advance(self._st, self.beta1, self.beta2)
        return self._out.cpu().numpy().astype(np.float64)


class SharedStats:
    """utils.py:59-106: running mean / variance / count of the observations, fp32 like the reference's torch tensors.
    This is the master-side bookkeeping object (2*d0+1 numbers); the per-generation statistics of whole tapes or
    rollouts are produced on the device (des_obs_stats_merge / des_obs_stats_merge_totals) and land here through
    load_state_dict / merge."""

    def __init__(self, o_size):
        self.m = np.zeros(o_size, dtype=np.float32)
        self.v = np.zeros(o_size, dtype=np.float32)
        self.n = np.zeros(1, dtype=np.float32)

    def feed(self, o):
        """utils.py:68-73, one observation (fp32 arithmetic in the reference's order)."""
        o = np.asarray(o, dtype=np.float32).reshape(self.m.shape)
        n = self.n[0]
        new_m = self.m * (n / (n + np.float32(1))) + o / (n + np.float32(1))
        self.v[:] = self.v * (n / (n + np.float32(1))) + (o - self.m) * (o - new_m) / (n + np.float32(1))
        self.m[:] = new_m
        self.n += np.float32(1)

    def zero(self):
        self.m[:] = 0
        self.v[:] = 0
        self.n[:] = 0

    def load(self, stats):
        self.m[:], self.v[:], self.n[:] = stats.m, stats.v, stats.n

    def merge(self, B):
        """utils.py:85-96 (Chan merge), fp32 in the reference's order; merging empty statistics is a no-op (the reference
        divides 0/0 there and poisons the statistics — not replicated)."""
        n_A, n_B = self.n[0], B.n[0]
        if n_B == 0:
            return
        n = n_A + n_B
        delta = B.m - self.m
        m = self.m + delta * n_B / n
        v = self.v * n_A + B.v * n_B + delta * delta * n_A * n_B / n
        v = v / n
        self.m[:] = m
        self.v[:] = v
        self.n += B.n

    def state_dict(self):
        return {'m': self.m, 'v': self.v, 'n': self.n}

    def load_state_dict(self, saved):
        self.m, self.v, self.n = (np.array(saved[k], dtype=np.float32) for k in ('m', 'v', 'n'))

    def as_device_tensor(self, device):
        """[m | v | n] fp32 on the device: the layout des_obs_normalize / des_obs_stats_merge take."""
        return torch.from_numpy(np.concatenate([self.m, self.v, self.n]).astype(np.float32)).to(device)


class StaticNormalizer:
    """utils.py:37-57: feeds the online statistics and applies the offline ones, (o - m)/sqrt(v + 1e-6), or passes the
    observation through while the offline statistics are empty (utils.py:48-49).  Single observations are handled here
    (host bookkeeping, as in the reference's master process); whole tapes go through normalize_tape (device)."""

    def __init__(self, o_size):
        self.offline_stats = SharedStats(o_size)
        self.online_stats = SharedStats(o_size)

    def __call__(self, o_):
        scalar = np.isscalar(o_)
        o = np.asarray([o_] if scalar else o_, dtype=np.float32)
        self.online_stats.feed(o.reshape(-1))
        if self.offline_stats.n[0] == 0:
            return o_
        std = (self.offline_stats.v + np.float32(1e-6)) ** np.float32(.5)
        out = ((o.reshape(-1) - self.offline_stats.m) / std).astype(np.float32)
        return float(out[0]) if scalar else out.reshape(np.shape(o_))

    def normalize_tape(self, obs_dev, device):
        """The whole observation tape on the device: des_obs_normalize with the offline statistics, and the tape's
        statistics Chan-merged into the online ones by des_obs_stats_merge (what feeding every row would accumulate,
        up to fp32 rounding order)."""
        from . import ops
        T = int(obs_dev.shape[0])
        online = self.online_stats.as_device_tensor(device)
        ops.obs_stats_merge(online, obs_dev, T)
        st = online.cpu().numpy()
        d0 = self.online_stats.m.size
        self.online_stats.load_state_dict({'m': st[:d0], 'v': st[d0:2 * d0], 'n': st[2 * d0:]})
        if self.offline_stats.n[0] == 0:
            return obs_dev
        return ops.obs_normalize(obs_dev, self.offline_stats.as_device_tensor(device))


class Evaluator:
    """utils.py:108-139.  eval(solution) -> (cost = -mean return, steps) for one flat weight vector on the
    tape env: one noiseless des_nes_eval call (sigma = 0, one member)."""

    def __init__(self, config, state_normalizer, device=None):
        self.config = config
        self.model = config.model_fn()
        self.repetitions = config.repetitions
        self.env = config.env_fn()
        self.state_normalizer = state_normalizer
        self.device = _device(device)
        self._obs = torch.from_numpy(self.env.obs).to(self.device)
        self._target = torch.from_numpy(self.env.target).to(self.device)

    def eval(self, solution):
        self.model.set_weight(solution)
        rewards, steps = [], []
        for _ in range(self.repetitions):
            reward, step = self.single_run()
            rewards.append(reward)
            steps.append(step)
        return -np.mean(rewards), np.sum(steps)

    def single_run(self):
        from . import ops
        obs = self.state_normalizer.normalize_tape(self._obs, self.device)       # utils.py:131, whole tape at once
        theta = torch.from_numpy(self.model.get_weight()).to(self.device)
        fit = ops.nes_eval(theta, obs, self._target, hidden=self.config.hidden_size, sigma=0.0,
                           clip=self.config.clip, seed=0, generation=0, member_offset=0, n_local=1, precision='fp32')
        return float(fit[0]), self.env.tape_len
