"""Flat-weight codec and the policy network surface of the reference (model.py:7-39).

`StandardFCNet` keeps the reference's constructor and get_weight/set_weight contract (flat layout
[fc1.weight | fc1.bias | fc2.weight | fc2.bias | fc3.weight | fc3.bias], model.py:17-24,30-32) but holds
no torch modules: the weights are one flat fp32 vector, which is what the CUDA kernels consume.  The
forward itself runs inside des_nes_eval on the GPU.
"""
from __future__ import annotations

import numpy as np


def param_count(state_dim, hidden_size, action_dim):
    return state_dim * hidden_size + hidden_size + hidden_size * hidden_size + hidden_size + \
        hidden_size * action_dim + action_dim


class BaseModel:
    def get_weight(self):
        """model.py:8-13 — flat fp32 copy."""
        return self.flat.copy()

    def set_weight(self, solution):
        """model.py:15-25 — accepts fp64, stores fp32, asserts the vector is fully consumed."""
        solution = np.asarray(solution).reshape(-1)
        assert solution.size == self.flat.size      # model.py:25
        self.flat = solution.astype(np.float32)


class StandardFCNet(BaseModel):
    def __init__(self, state_dim, action_dim, hidden_size, seed=None):
        self.state_dim, self.action_dim, self.hidden_size = int(state_dim), int(action_dim), int(hidden_size)
        rs = np.random.RandomState(seed)
        parts = []
        # nn.Linear default init (model.py:30-32): U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias
        for fan_out, fan_in in ((hidden_size, state_dim), (hidden_size, hidden_size), (action_dim, hidden_size)):
            b = 1.0 / np.sqrt(fan_in)
            parts.append(rs.uniform(-b, b, size=fan_out * fan_in))
            parts.append(rs.uniform(-b, b, size=fan_out))
        self.flat = np.concatenate(parts).astype(np.float32)

    def parameters(self):
        """Views in nn.Module.parameters() order (fc1.weight, fc1.bias, ...)."""
        d0, H, A = self.state_dim, self.hidden_size, self.action_dim
        out, o = [], 0
        for shape in ((H, d0), (H,), (H, H), (H,), (A, H), (A,)):
            n = int(np.prod(shape))
            out.append(self.flat[o:o + n].reshape(shape))
            o += n
        return out
