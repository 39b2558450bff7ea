"""Tiny end-to-end run for compute-sanitizer (memcheck): every kernel once, small shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from distributedes_b200 import ops
from distributedes_b200.engine import NESEngine
from distributedes_b200.envs import TapeEnv
from distributedes_b200.model import StandardFCNet
for (H, T, prec, N) in [(64, 256, 'f16x3', 300), (256, 256, 'f16x3', 200), (256, 256, 'f16', 200), (64, 128, 'f16', 50), (20, 70, 'fp32', 40), (256, 512, 'f16x3', 160)]:
    d0, A = (24, 4) if H != 20 else (5, 3)
    env = TapeEnv(d0, A, T)
    eng = NESEngine(state_dim=d0, hidden=H, action_dim=A, pop_size=N, theta0=StandardFCNet(d0, A, H, seed=0).get_weight(),
                    obs=env.obs, target=env.target, sigma=0.1, learning_rate=0.1, clip=1.0, seed=1, precision=prec,
                    device='cuda:0', normalize_obs=True)
    for _ in range(2):
        eng.generation()
    torch.cuda.synchronize()
    print(H, T, prec, float(eng.fitness_all.mean()))
f = torch.randn(20000, device='cuda'); ops.centered_rank(f)
Y = torch.randn(40, 300, device='cuda'); w = torch.rand(40, device='cuda'); dC = ops.cma_rank_mu(Y, w)
ops.cma_cov_apply(torch.eye(300, device='cuda'), dC, torch.randn(300, device='cuda'), decay=0.9, c1=0.01, cmu=0.02)
from distributedes_b200.engine import RolloutEngine
for H, reps in [(64, 10), (32, 3), (128, 7)]:
    reng = RolloutEngine(hidden=H, pop_size=9, theta0=StandardFCNet(3, 1, H, seed=0).get_weight(), sigma=0.1, learning_rate=0.1,
                         repetitions=reps, horizon=25, seed=2, action_noise_std=0.1)
    for _ in range(2):
        reng.generation()
    print('rollout', H, reps, float(reng.fitness_all.mean()), reng.test_returns().mean())
ops.noise_fill(3, 1001, 1, 2); torch.cuda.synchronize(); print('ok')
