"""Fitness of a member must not depend on the sharding: full launch vs two half launches, bit for bit (one GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import nes_oracle as orc
from distributedes_b200 import ops
d0, A, T = 24, 4, 256
for H in (64, 128, 256):
    for N in (1000, 1001, 4096):
        obs, target = orc.synthetic_tape(T, d0, A)
        th = torch.from_numpy(orc.synthetic_theta(d0, H, A)).cuda(); o = torch.from_numpy(obs).cuda(); t = torch.from_numpy(target).cuda()
        def run(off, n):
            return ops.nes_eval(th, o, t, hidden=H, sigma=0.1, clip=1.0, seed=21, generation=0, member_offset=off, n_local=n, precision='f16x3').cpu().numpy()
        full = run(0, N); full2 = run(0, N)
        h = N // 2
        a = run(0, h); b = run(h, N - h)
        sh = np.concatenate([a, b])
        bad = np.nonzero(full != sh)[0]
        print(H, N, 'deterministic', np.array_equal(full, full2), 'shard-equal', np.array_equal(full, sh), 'first mismatches', bad[:8], 'max rel', float(np.max(np.abs(full - sh) / np.abs(full))))
