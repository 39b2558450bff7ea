"""CPU timing leg of bench.py (`cpu_baseline` and `--impl reference`).  TEST/BENCH INFRASTRUCTURE.

The reference's CPU path for one generation is: every worker process draws eps, perturbs theta, rolls the
policy over the episode, returns (eps, fitness) (natural_es.py:21-32); the master rank-shapes and forms
mean(eps * s)/sigma, Adam, step (natural_es.py:90-96).  /root/reference cannot travel to the GPU box and is
Python-only (nothing to compile into oracle/_ref), so this is the oracle PORT (kind "port") of that path,
arranged the way the reference arranges it — one worker process per core, members handed out in chunks —
but with the per-member forward done as three BLAS matmuls over the whole tape instead of the reference's
per-step batch-1 torch calls (which cost ~220 us of Python dispatch per step).  It is
therefore a considerably FASTER CPU baseline than the reference itself; BASELINE.md holds the verbatim
reference's measured numbers.
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import nes_oracle as orc

_G = {}


def usable_cores():
    """Host threads this process may really use: min(affinity mask, cgroup cpu quota)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for path in ('/sys/fs/cgroup/cpu.max', '/sys/fs/cgroup/cpu/cpu.cfs_quota_us'):
        try:
            txt = open(path).read().split()
            if path.endswith('cpu.max'):
                if txt[0] != 'max':
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                if q > 0:
                    period = int(open('/sys/fs/cgroup/cpu/cpu.cfs_period_us').read())
                    n = min(n, max(1, int(q / period + 0.5)))
            break
        except Exception:
            continue
    return max(1, n)


def _init(theta, obs, target, d0, H, A, sigma, clip, seed):
    os.environ['OMP_NUM_THREADS'] = '1'
    _G.update(theta=theta, obs=obs, target=target, dims=(d0, H, A), sigma=sigma, clip=clip, seed=seed)
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass


def _eval_chunk(args):
    """Members [start, start+n) of generation gen: returns (fitness[n], time spent)."""
    gen, start, n = args
    d0, H, A = _G['dims']
    P = orc.param_count(d0, H, A)
    t0 = time.perf_counter()
    obs32 = _G['obs']
    fit = np.empty(n)
    for i in range(n):
        eps = orc.noise(_G['seed'], gen, start + i, 1, P)[0]                      # natural_es.py:29
        th = orc.perturb(_G['theta'], _G['sigma'], eps)                           # :28-30
        W1, b1, W2, b2, W3, b3 = orc.unflatten(th, d0, H, A)
        h = np.tanh(obs32 @ W1.T + b1)                                            # model.py:36 (fp32, like torch CPU)
        h = np.tanh(h @ W2.T + b2)                                                # model.py:37
        a = h @ W3.T + b3                                                         # model.py:38
        fit[i] = orc.tape_fitness(a, _G['target'], _G['clip'])                   # utils.py:134-137
    return fit, time.perf_counter() - t0


def _grad_chunk(args):
    """sum_i s_i eps_i for members [start, start+n) (natural_es.py:91, eps regenerated as on the GPU)."""
    gen, start, shaped = args
    d0, H, A = _G['dims']
    P = orc.param_count(d0, H, A)
    return np.asarray(shaped, dtype=np.float64) @ orc.noise(_G['seed'], gen, start, len(shaped), P)


class CpuGeneration:
    """One NES generation over a bounded SAMPLE of `sample` members on `procs` worker processes."""

    def __init__(self, d0, H, A, T, sample, sigma=0.1, clip=1.0, seed=0, lr=0.1, wd=0.005, procs=None):
        import multiprocessing as mp
        self.dims = (d0, H, A)
        self.T, self.sample = T, int(sample)
        self.sigma, self.lr, self.wd, self.seed = sigma, lr, wd, seed
        self.procs = procs or usable_cores()
        self.theta = orc.synthetic_theta(d0, H, A)
        obs, target = orc.synthetic_tape(T, d0, A)
        self.opt = orc.Adam()
        self.gen = 0
        # 'spawn', not 'fork': the bench process may already hold CUDA/OpenMP threads, and forking those deadlocks BLAS
        ctx = mp.get_context('spawn')
        self.pool = ctx.Pool(self.procs, initializer=_init,
                             initargs=(self.theta, obs, target, d0, H, A, sigma, clip, seed))
        per = max(1, self.sample // (self.procs * 2))
        self.chunks = [(s, min(per, self.sample - s)) for s in range(0, self.sample, per)]

    def step(self):
        """evaluate -> rank -> gradient -> Adam/step over the sample; returns seconds."""
        t0 = time.perf_counter()
        res = self.pool.map(_eval_chunk, [(self.gen, s, n) for s, n in self.chunks])
        fitness = np.concatenate([r[0] for r in res])
        shaped = orc.fitness_shift(fitness)                                       # natural_es.py:90
        parts = self.pool.map(_grad_chunk, [(self.gen, s, shaped[s:s + n]) for s, n in self.chunks])
        g = np.sum(parts, axis=0) / self.sample / self.sigma                      # :91-92
        self.theta, _ = orc.nes_update(self.theta, g, self.opt, self.wd, self.lr)  # :93-96
        self.gen += 1
        # workers keep the theta they were forked with: the timing does not depend on its value
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def calibrated_sample(d0, H, A, T, pop, target_seconds=5.0, procs=None):
    """Pick the per-step sample so one CPU step takes about target_seconds: time a tiny step first."""
    procs = procs or usable_cores()
    probe = max(16 * procs, 64)
    g = CpuGeneration(d0, H, A, T, probe, procs=procs)
    try:
        g.step()                       # warm the pool
        sec = g.step()
    finally:
        g.close()
    rate = probe / max(sec, 1e-6)
    return int(max(probe, min(pop, rate * target_seconds)))


def time_cpu_generation(d0, H, A, T, sample, steps=1, warmup=0, procs=None):
    g = CpuGeneration(d0, H, A, T, sample, procs=procs)
    try:
        for _ in range(warmup):
            g.step()
        times = [g.step() for _ in range(steps)]
    finally:
        g.close()
    return dict(seconds_per_step=float(np.mean(times)), evals_per_sec=sample / float(np.mean(times)),
                cores=g.procs, sample='%d of the population members per step (all T=%d observations each)' % (sample, T))


def main(argv=None):
    """CLI used by bench.py so the CPU leg runs in its own process (no torch / CUDA state):
    python -m oracle.cpu_baseline --d0 24 --hidden 256 --action-dim 4 --tape-len 256 --pop 65536 --steps 2 --warmup 1"""
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument('--d0', type=int, default=24)
    ap.add_argument('--hidden', type=int, default=256)
    ap.add_argument('--action-dim', type=int, default=4)
    ap.add_argument('--tape-len', type=int, default=256)
    ap.add_argument('--pop', type=int, default=65536)
    ap.add_argument('--sample', type=int, default=0)
    ap.add_argument('--target-seconds', type=float, default=5.0)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--warmup', type=int, default=1)
    a = ap.parse_args(argv)
    sample = a.sample or calibrated_sample(a.d0, a.hidden, a.action_dim, a.tape_len, a.pop, a.target_seconds)
    r = time_cpu_generation(a.d0, a.hidden, a.action_dim, a.tape_len, sample, steps=a.steps, warmup=a.warmup)
    r['sample_members'] = sample
    print(json.dumps(r), flush=True)


if __name__ == '__main__':
    main()
