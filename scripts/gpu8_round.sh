#!/bin/bash
# N-GPU scaling check (torchrun, NCCL), short timeouts.  usage: gpu8_round.sh <ngpus>
N=${1:-8}
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_scale_n$N.json 2> gpurun_out/bench_scale_n$N.err; echo "rc=$?"
tail -c 600 gpurun_out/bench_scale_n$N.json; tail -3 gpurun_out/bench_scale_n$N.err | cut -c1-300
