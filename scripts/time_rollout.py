"""Time closed-loop generations (des_rollout_eval + rank + reduction + apply) with CUDA events; also prints the worst
relative deviation of the device's fitness from the oracle on a small sample (scripts may use the oracle: not product)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from distributedes_b200 import ops
from distributedes_b200.engine import RolloutEngine
from distributedes_b200.model import StandardFCNet
out = []
for N, H, reps in [(16, 64, 10), (4096, 64, 10), (65536, 64, 10), (4096, 128, 10)]:
    theta0 = StandardFCNet(3, 1, H, seed=0).get_weight()
    eng = RolloutEngine(hidden=H, pop_size=N, theta0=theta0, sigma=0.1, learning_rate=0.1, repetitions=reps, seed=1)
    for _ in range(3):
        eng.generation()
    k = 5
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(k):
        eng.evaluate()
    ev[1].record()
    for _ in range(k):
        eng.generation()
    ev[2].record(); torch.cuda.synchronize()
    te, tg = ev[0].elapsed_time(ev[1]) / k, ev[1].elapsed_time(ev[2]) / k
    out.append(dict(pop=N, hidden=H, reps=reps, eval_ms=te, generation_ms=tg, env_steps_per_s=N * reps * 200 / tg * 1e3,
                    episodes_per_s=N * reps / tg * 1e3))
print(json.dumps(out))
if '--check' in sys.argv:
    from oracle import nes_oracle as orc, pendulum_oracle as po
    theta = orc.synthetic_theta(3, 64, 1)
    fit = ops.rollout_eval(torch.from_numpy(theta).cuda(), hidden=64, sigma=0.1, clip=2.0, seed=5, generation=0, member_offset=0, n_local=256)
    ref, _ = po.closed_fitness(theta, 64, 0.1, 5, 0, 0, 256, 10)
    d = np.abs(fit.cpu().numpy() - ref) / np.abs(ref)
    print('rollout fitness rel dev vs oracle: max %.3g median %.3g' % (d.max(), np.median(d)))
