"""The reference's own natural_es.train() run VERBATIM on the host cores and timed.  TEST/BENCH INFRASTRUCTURE.

    python -m oracle.ref_cpu_baseline --d0 24 --hidden 256 --action-dim 4 --tape-len 256 --pop 256 --gens 2

Imports the unmodified reference from oracle/_ref (oracle/build_ref.py), puts oracle/gym_stub first on sys.path (the
reference does `import gym`, config.py:1; gym is not installed) and runs natural_es.train(config) — its own worker
processes (fork), queues, busy-wait master, per-step batch-1 torch forward, fitness_shift, Adam — on the synthetic
observation-tape environment of SURVEY §8d (`SynthTape-d<d0>-a<A>-T<T>-v0` of the stub: the same tape and reward the
GPU bench uses).  num_workers = usable cores - 1 (the master spins on a core, natural_es.py:68-69), one torch thread per
process.  Seconds per generation come from the `training_timestamps` train() returns (natural_es.py:58,99): they
include the master's test() episodes (natural_es.py:54) as SURVEY §8d notes.  Nothing is patched: the noise is the
reference's own np.random.randn, so only the timing is used.

Prints one JSON line: policy-evals/s = pop / seconds_per_generation.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, '_ref')


def usable_cores():
    from .cpu_baseline import usable_cores as u
    return u()


def run(d0, H, A, T, pop, gens, workers=None, test_repetitions=1, skip=0):
    if not os.path.exists(os.path.join(REFDIR, 'natural_es.py')):
        raise RuntimeError('oracle/_ref is empty: run `python oracle/build_ref.py` where /root/reference exists')
    sys.path.insert(0, os.path.join(HERE, 'gym_stub'))
    sys.path.insert(0, REFDIR)
    import numpy as np
    import torch
    torch.set_num_threads(1)
    os.environ['OMP_NUM_THREADS'] = '1'
    os.environ['MKL_NUM_THREADS'] = '1'
    import config as ref_config          # oracle/_ref/config.py
    import natural_es as ref_nes         # oracle/_ref/natural_es.py
    import utils as ref_utils            # oracle/_ref/utils.py
    ref_utils.logger.setLevel('WARNING')

    class TapeConfig(ref_config.BasicConfig):
        def __init__(self):
            self.task = 'SynthTape-d%d-a%d-T%d-v0' % (d0, A, T)
            self.action_clip = lambda a: np.clip(a, -1, 1)
            self.target = 10000
            torch.manual_seed(0)
            ref_config.BasicConfig.__init__(self, H)

    cfg = TapeConfig()
    cores = usable_cores()
    cfg.num_workers = int(workers) if workers else max(1, cores - 1)
    cfg.repetitions = 1                  # one episode of T steps per member: the tape is deterministic
    cfg.test_repetitions = test_repetitions
    cfg.pop_size = pop
    cfg.sigma = 0.1
    cfg.learning_rate = 0.1
    # train() stops BEFORE the update once total_steps > max_steps (natural_es.py:82-84): gens+1 collections
    cfg.max_steps = (gens + 1) * pop * T - 1
    t0 = time.time()
    rewards, steps, stamps = ref_nes.train(cfg)
    wall = time.time() - t0
    stamps = np.asarray(stamps, dtype=np.float64)
    # stamps[k] is taken at the top of generation k (after its test()): differences = one full generation each
    per_gen = np.diff(stamps)[int(skip):]          # the first `skip` generations are warm-up
    sec = float(np.mean(per_gen)) if len(per_gen) else wall
    return {'seconds_per_generation': sec, 'evals_per_sec': pop / sec, 'generations_timed': int(len(per_gen)),
            'pop': pop, 'cores': cores, 'workers': cfg.num_workers, 'wall_seconds': wall,
            'sample': 'reference natural_es.train() verbatim: pop %d, %d workers + spinning master, T=%d steps per member, '
                      '%d full generation(s) timed from training_timestamps (includes the master\'s test())'
                      % (pop, cfg.num_workers, T, len(per_gen))}


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--d0', type=int, default=24)
    ap.add_argument('--hidden', type=int, default=256)
    ap.add_argument('--action-dim', type=int, default=4)
    ap.add_argument('--tape-len', type=int, default=256)
    ap.add_argument('--pop', type=int, default=256)
    ap.add_argument('--gens', type=int, default=2)
    ap.add_argument('--workers', type=int, default=0)
    ap.add_argument('--skip', type=int, default=0)
    a = ap.parse_args()
    print(json.dumps(run(a.d0, a.hidden, a.action_dim, a.tape_len, a.pop, a.gens, a.workers or None, skip=a.skip)))
