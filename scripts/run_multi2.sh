mkdir -p gpurun_out
nvidia-smi -L | head -3
echo "== multi-gpu tests"; timeout 900 python -m pytest tests/test_gpu_multi.py -q --tb=short 2>&1 | tail -25
echo "== bench 2 gpus (peer comm)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 --no-other-modes > gpurun_out/bench_n2_peer.json 2> gpurun_out/bench_n2_peer.err; echo rc=$?; tail -c 3000 gpurun_out/bench_n2_peer.json; tail -5 gpurun_out/bench_n2_peer.err
echo "== bench 2 gpus (nccl)"; DES_COMM=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 --no-other-modes --no-configs > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err; echo rc=$?; python -c "
import json; d=json.load(open('gpurun_out/bench_n2_nccl.json')); print(d['ms_per_step'], d['value'], d['parity'])"
