"""Time des_nes_grad_partial (CUDA events).  python scripts/time_grad.py [pop] [hidden] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import numpy as np
import torch
from oracle import nes_oracle as orc
from distributedes_b200 import ops
pop = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
H = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
P = orc.param_count(24, H, 4)
shaped = torch.from_numpy((np.random.RandomState(1).permutation(pop) / (pop - 1) - 0.5).astype(np.float32)).cuda()
for _ in range(2):
    g = ops.nes_grad_partial(shaped, P, seed=1, generation=1, member_offset=0)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for i in range(reps):
    g = ops.nes_grad_partial(shaped, P, seed=1, generation=1, member_offset=0)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
print(json.dumps({'pop': pop, 'P': P, 'grad_ms': [round(x, 3) for x in ms]}))
