"""BASELINE configs[4] on G GPUs (torchrun): rank-mu partial of lambda/G members at n = 4096, all-reduce of the [n, n]
partial over NCCL, covariance update — CUDA events on the launching stream, max over ranks.  One JSON line from rank 0."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from distributedes_b200 import ops
from distributedes_b200.engine import shard_bounds

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
n, lam = 4096, 1024
off, nl = shard_bounds(lam, world, rank)
Y = ops.noise_fill(nl, n, 0, 0, member_offset=off, stream_tag=1, device=dev)
w = torch.rand(nl, device=dev) / lam
C = torch.eye(n, device=dev); pc = torch.randn(n, device=dev); dC = torch.empty(n, n, device=dev)


tiles = torch.empty(ops.cma_packed_elems(n), device=dev)


def step(collective=True, packed=False):
    if packed:          # upper-triangular tiles: half the all-reduce payload (what cma_es.CMAEvolutionStrategy uses)
        ops.cma_rank_mu_packed(Y, w, out=tiles)
        if world > 1 and collective:
            dist.all_reduce(tiles)
        ops.cma_cov_apply_packed(C, tiles, pc, decay=0.99, c1=1e-4, cmu=1e-3)
        return
    ops.cma_rank_mu(Y, w, out=dC)
    if world > 1 and collective:
        dist.all_reduce(dC)
    ops.cma_cov_apply(C, dC, pc, decay=0.99, c1=1e-4, cmu=1e-3)


def timed(fn, k=20):
    for _ in range(3):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / k], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


full = timed(step)
compute = timed(lambda: step(False))
ar = timed(lambda: dist.all_reduce(dC)) if world > 1 else 0.0
full_p = timed(lambda: step(True, True))
compute_p = timed(lambda: step(False, True))
ar_p = timed(lambda: dist.all_reduce(tiles)) if world > 1 else 0.0
if rank == 0:
    print(json.dumps(dict(workload='CMA rank-mu update n=4096 lambda=1024', n_gpus=world, members_per_gpu=nl,
                          ms_per_update=full, updates_per_sec=1e3 / full, compute_only_ms=compute, allreduce_only_ms=ar,
                          allreduce_bytes=4 * n * n, packed=dict(ms_per_update=full_p, updates_per_sec=1e3 / full_p, compute_only_ms=compute_p, allreduce_only_ms=ar_p, allreduce_bytes=4 * tiles.numel()), fp32_tflops_counted_2lambda_n2=2.0 * lam * n * n / full / 1e9)))
if world > 1:
    dist.destroy_process_group()
