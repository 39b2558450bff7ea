#!/usr/bin/env python
"""Benchmark of the NES generation hot path (BASELINE.json: generations/s and policy-evals/s, pop 64k).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--pop 65536] [--hidden 256] [--tape-len 256] [--precision fp32|f16|f16x3]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, NCCL)

A "step" is one NES generation (natural_es.py:62-96) over synthetic inputs: sample eps for the whole
population, batched policy forward over population x tape, centered ranks, fitness x noise reduction,
(1-wd)/Adam/step.  The population is fixed as GPUs are added (strong scaling, as BASELINE.json quotes
the metric "at pop 64k, 1/2/4/8 B200").  Prints ONE JSON line on rank 0.

  value       policy-evals/s with everything resident in HBM (generations/s = value / pop)
  e2e         same metric through the host-buffer API (tape H2D, theta+fitness D2H inside the timed region)
  roofline    the dominant kernel (fused sample+forward+fitness) against the materialised-noise HBM contract
              of SURVEY §8d: 8*n_local*P algorithmic bytes per launch (write eps once + read it in the forward)
  cpu_baseline  the oracle port of the same generation on the host cores, bounded sample (N=1, rank 0)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

EMIT = print


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--pop', type=int, default=65536)
    ap.add_argument('--hidden', type=int, default=256)
    ap.add_argument('--state-dim', type=int, default=24)
    ap.add_argument('--action-dim', type=int, default=4)
    ap.add_argument('--tape-len', type=int, default=256)
    ap.add_argument('--precision', default=os.environ.get('DES_BENCH_PRECISION', 'f16x3'),
                    help='policy-forward arithmetic: f16x3 (tensor cores, fp32-grade, default), f16 (tensor cores, fp16 operands), fp32 (CUDA cores)')
    ap.add_argument('--no-other-modes', action='store_true', help='skip the short extra measurement of the other tensor-core mode')
    ap.add_argument('--cpu-sample', type=int, default=0, help='members per CPU-baseline step (0 = auto)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true')
    return ap.parse_args()


def workload_name(a):
    return 'nes_synth_tape d0=%d H=%d A=%d T=%d pop=%d (SURVEY 8d cfg4 shape; strong scaling)' % (
        a.state_dim, a.hidden, a.action_dim, a.tape_len, a.pop)


def peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), float(d.get('bf16_tflops', 0.0)), 'measured'
    return 6650.0, 1590.0, 'fallback'


# --------------------------------------------------------------------------------------------------------
# reference arm: the oracle port of the generation on the host cores
# --------------------------------------------------------------------------------------------------------
def cpu_generation_subprocess(a, steps, warmup, target_seconds):
    """Run the oracle port of the generation in its own process (no torch / CUDA state in it)."""
    cmd = [sys.executable, '-m', 'oracle.cpu_baseline', '--d0', str(a.state_dim), '--hidden', str(a.hidden),
           '--action-dim', str(a.action_dim), '--tape-len', str(a.tape_len), '--pop', str(a.pop), '--steps', str(steps),
           '--warmup', str(warmup), '--sample', str(a.cpu_sample), '--target-seconds', str(target_seconds)]
    r = subprocess.run(cmd, cwd=REPO, capture_output=True, text=True, timeout=900)
    if r.returncode != 0:
        raise RuntimeError('cpu baseline failed: %s' % r.stderr[-2000:])
    return json.loads(r.stdout.strip().splitlines()[-1])


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    r = cpu_generation_subprocess(a, a.steps, a.warmup, 150.0 / max(1, a.steps + a.warmup + 2))
    value = r['evals_per_sec']
    line = {
        'impl': 'reference', 'metric': 'nes_policy_evals_per_sec', 'value': value, 'unit': 'policy-evals/s',
        'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': r['seconds_per_step'] * 1e3,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64/f32 (numpy)',
        'data': 'synthetic', 'generations_per_sec': value / a.pop,
        'config': {'workload': workload_name(a), 'sample_members_per_step': r['sample_members']},
        'cpu_baseline': {'value': value, 'unit': 'policy-evals/s', 'cores': r['cores'], 'kind': 'port',
                         'sample': r['sample']},
        'e2e': {'value': value, 'unit': 'policy-evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    EMIT(json.dumps(line))


# --------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '100'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(', ') for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower() == 'active':
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(np.max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm)}


def run_ours(a):
    import torch
    import torch.distributed as dist
    from distributedes_b200.envs import TapeEnv                 # synthetic tape, RandomState(1234)
    from distributedes_b200.model import StandardFCNet          # nn.Linear-style init, RandomState(0)
    from distributedes_b200.engine import NESEngine
    from distributedes_b200 import _lib

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU arm)')
    _lib.load()
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    if world != a.gpus and rank == 0:
        print('bench.py: --gpus %d but WORLD_SIZE=%d; using WORLD_SIZE' % (a.gpus, world), file=sys.stderr)

    d0, H, A, T, N = a.state_dim, a.hidden, a.action_dim, a.tape_len, a.pop
    env = TapeEnv(d0, A, T)
    obs, target = env.obs, env.target
    theta0 = StandardFCNet(d0, A, H, seed=0).get_weight()
    eng = NESEngine(state_dim=d0, hidden=H, action_dim=A, pop_size=N, theta0=theta0, obs=obs, target=target,
                    sigma=0.1, learning_rate=0.1, weight_decay=0.005, clip=1.0, seed=0, precision=a.precision,
                    device=dev, use_graph=not a.no_graph)
    P = eng.P
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, per_step_hook=None):
        """K steps, each bracketed by CUDA events on the launching stream; L2 flushed between steps
        (outside the events).  Returns (sum of step ms, list)."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for s, e in ev:
            flush.zero_()
            s.record()
            fn()
            e.record()
        barrier()
        ms = [s.elapsed_time(e) for s, e in ev]
        return float(np.sum(ms)), ms

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident generations ----
    for _ in range(max(a.warmup, 3)):
        eng.generation()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    total_ms, _ = timed(eng.generation, a.steps)
    clocks = sampler.stop() if sampler else None
    total_ms = max_over_ranks(total_ms)
    ms_per_step = total_ms / a.steps
    value = N / (ms_per_step * 1e-3)

    # ---- dominant kernel alone (fused sample+forward+fitness) ----
    def eval_only():
        eng.k.nes_eval(eng.theta, eng.obs, eng.target, hidden=H, sigma=eng.sigma, clip=eng.clip, seed=eng.seed,
                       state=eng.state, member_offset=eng.offset, n_local=eng.n_local, precision=eng.precision,
                       out=eng.fitness_all[eng.offset:eng.offset + eng.n_local], workspace=eng.eval_ws)
    for _ in range(2):
        eval_only()
    ev_ms, _ = timed(eval_only, a.steps)
    ev_ms = max_over_ranks(ev_ms) / a.steps
    hbm_peak, bf16_peak, peak_kind = peaks()
    alg_bytes = 8.0 * eng.n_local * P                                  # write eps once + read it in the forward
    achieved = alg_bytes / (ev_ms * 1e-3) / 1e9
    fwd_flops = 2.0 * eng.n_local * T * (d0 * H + H * H + H * A)
    traffic = None
    tp = os.path.join(REPO, 'profiles', 'roofline_traffic.json')
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get('%s_H%d_pop%d' % (a.precision, H, eng.n_local))
        except Exception:
            traffic = None
    roofline = {'kernel': 'des_nes_eval[%s]' % a.precision, 'bound': 'hbm', 'achieved': achieved, 'peak': hbm_peak,
                'unit': 'GB/s', 'frac': achieved / hbm_peak, 'traffic': traffic, 'peak_kind': 'of ' + peak_kind,
                'kernel_ms': ev_ms, 'kernel_share_of_step': ev_ms / ms_per_step,
                'algorithmic_bytes_per_launch': alg_bytes,
                'note': 'materialised-noise contract (SURVEY 8d): eps is regenerated, not stored, so effective GB/s '
                        'may exceed the HBM peak; tensor view alongside',
                'tensor': {'achieved_tflops': fwd_flops / (ev_ms * 1e-3) / 1e12, 'peak_bf16_tflops': bf16_peak,
                           'frac': fwd_flops / (ev_ms * 1e-3) / 1e12 / bf16_peak if bf16_peak else None}}

    # ---- end to end through the host-buffer API ----
    obs_h = torch.from_numpy(obs).pin_memory()
    tgt_h = torch.from_numpy(target).pin_memory()
    theta_h = torch.empty(P, dtype=torch.float32).pin_memory()
    fit_h = torch.empty(N, dtype=torch.float32).pin_memory()

    def e2e_step():
        eng.generation_host(obs_h, tgt_h, theta_out_host=theta_h, fitness_out_host=fit_h)
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / a.steps
    e2e = {'value': N / e2e_s, 'unit': 'policy-evals/s', 'ms_per_step': e2e_s * 1e3,
           'h2d_bytes_per_step': int(obs_h.numel() * 4 + tgt_h.numel() * 4),
           'd2h_bytes_per_step': int(theta_h.numel() * 4 + fit_h.numel() * 4),
           'api': 'NESEngine.generation_host (pinned host tape in, theta + fitness out, synchronous)'}

    # ---- the other tensor-core mode, device-resident, for context (not the headline) ----
    other = None
    if not a.no_other_modes and a.precision in ('f16', 'f16x3'):
        oprec = 'f16' if a.precision == 'f16x3' else 'f16x3'
        try:
            eng2 = NESEngine(state_dim=d0, hidden=H, action_dim=A, pop_size=N, theta0=theta0, obs=obs, target=target,
                             sigma=0.1, learning_rate=0.1, weight_decay=0.005, clip=1.0, seed=0, precision=oprec,
                             device=dev, use_graph=not a.no_graph)
            for _ in range(3):
                eng2.generation()
            o_ms, _ = timed(eng2.generation, max(3, a.steps // 2))
            o_ms = max_over_ranks(o_ms) / max(3, a.steps // 2)
            other = {oprec: {'ms_per_step': o_ms, 'value': N / (o_ms * 1e-3), 'unit': 'policy-evals/s',
                             'note': 'fp16-rounded operands (11 significant bits, like TF32): fitness within 4e-3 of the oracle'
                             if oprec == 'f16' else 'hi/lo split operands: fitness within 3e-5 of the oracle'}}
            del eng2
        except Exception as e:
            other = {oprec: {'error': str(e)[:200]}}

    # ---- closed-loop rollouts, for context: Pendulum-v0 stepped on the device, 10 episodes of 200
    # steps per member, observation normaliser on.  Not the headline; single GPU only.
    closed = None
    if world == 1 and not a.no_other_modes:
        try:
            from distributedes_b200.engine import RolloutEngine
            cN, cH = min(N, 65536), 64
            ceng = RolloutEngine(hidden=cH, pop_size=cN, theta0=StandardFCNet(3, 1, cH, seed=0).get_weight(), sigma=0.1,
                                 learning_rate=0.1, seed=0, device=dev)
            for _ in range(3):
                ceng.generation()
            c_ms, _ = timed(ceng.generation, max(3, a.steps // 2))
            c_ms /= max(3, a.steps // 2)
            closed = {'workload': 'Pendulum-v0 closed loop: pop %d, 2x%d MLP, 10 episodes x 200 steps per member' % (cN, cH),
                      'ms_per_step': c_ms, 'env_steps_per_sec': cN * 10 * 200 / (c_ms * 1e-3),
                      'policy_evals_per_sec': cN / (c_ms * 1e-3),
                      'fp32_tflops': 2.0 * (3 * cH + cH * cH + cH) * cN * 2000 / (c_ms * 1e-3) / 1e12}
            del ceng
        except Exception as e:
            closed = {'error': str(e)[:200]}

    # ---- CPU baseline (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            r = cpu_generation_subprocess(a, 2, 1, 5.0)
            cpu = {'value': r['evals_per_sec'], 'unit': 'policy-evals/s', 'cores': r['cores'], 'kind': 'port',
                   'sample': r['sample']}
        except Exception as e:      # the CPU leg must never take the GPU line down with it
            cpu = {'value': None, 'unit': 'policy-evals/s', 'cores': None, 'kind': 'port', 'sample': 'failed: %s' % e}

    if rank == 0:
        # eval, rank (2 kernels; 5 on the bucketed path for N > 8192), grad_chunk, grad_reduce, apply, state_advance
        launches_per_step = 5 + (5 if N > 8192 else 2)
        line = {
            'metric': 'nes_policy_evals_per_sec', 'value': value, 'unit': 'policy-evals/s', 'n_gpus': world,
            'steps': a.steps, 'warmup': max(a.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': {'fp32': 'f32', 'f16': 'f16 (fp16 operands, f32 accumulate)',
                                                           'f16x3': 'f16x3 (split-fp16 operands ~ f32, f32 accumulate)'}[a.precision],
            'data': 'synthetic', 'generations_per_sec': 1e3 / ms_per_step,
            'forwards_per_sec': value * T,
            'config': {'workload': workload_name(a), 'precision': a.precision, 'param_count': P,
                       'members_per_gpu': eng.n_local, 'cuda_graph': bool(eng._use_graph),
                       'l2': 'flushed: 256 MiB memset between steps, outside the per-step CUDA events',
                       'parallelism': 'population sharded over %d GPU(s); all-reduce fitness[N] + all-reduce partial[P]' % world},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches_per_step * a.steps,
            'roofline': roofline, 'cpu_baseline': cpu, 'other_modes': other, 'closed_loop': closed,
        }
        EMIT(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _guard_stdout():
    """Keep stdout clean for the ONE JSON line: anything libraries print on fd 1 (NCCL writes its version banner
    there) goes to stderr; the JSON line is written to the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    real = os.fdopen(saved, 'w')

    def emit(line):
        real.write(line + '\n')
        real.flush()
    return emit


if __name__ == '__main__':
    args = parse()
    EMIT = _guard_stdout()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)
