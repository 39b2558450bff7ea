#!/bin/bash
# 2-GPU validation: NCCL parity test + torchrun bench, short timeouts.
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -8
for p in f16x3; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --precision $p --pop ${POP:-65536} 2>&1 | tail -3 | cut -c1-2500
done
