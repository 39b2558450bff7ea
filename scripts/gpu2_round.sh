#!/bin/bash
# 2-GPU validation: NCCL parity tests (NES, closed loop, CMA), torchrun bench, CMA rank-mu timing; short timeouts.
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_rollout.py -m gpu -q -x 2>&1 | tail -8
G=${G:-2}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29521 scripts/time_cma_dist.py 2>gpurun_out/cma_dist_$G.err | tee gpurun_out/cma_dist_$G.json
for p in f16x3; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 5 --warmup 3 --precision $p --pop ${POP:-65536} 2>gpurun_out/bench_n$G.err | tee gpurun_out/bench_n${G}_$p.json | cut -c1-600
done
