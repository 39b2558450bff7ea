mkdir -p gpurun_out
echo "== multi-gpu tests"; timeout 900 python -m pytest tests/test_gpu_multi.py -q --tb=short 2>&1 | tail -5
for comm in peer nccl; do
echo "== bench 2 gpus pop 16384 ($comm): per-rank load of the 8-GPU run"; DES_COMM=$comm timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2963$((RANDOM%9)) bench.py --gpus 2 --pop 16384 --steps 20 --warmup 5 --no-other-modes --no-configs > gpurun_out/bench_n2_pop16k_$comm.json 2> gpurun_out/bench_n2_pop16k_$comm.err; echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_n2_pop16k_$comm.json')); print(round(d['ms_per_step'],4), 'eval', round(d['roofline']['kernel_ms'],4), d['config']['cuda_graph'], d['parity']['ok_all_ranks'])"
done
echo "== bench 2 gpus pop 65536 (peer)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 10 --warmup 3 --no-other-modes --no-configs > gpurun_out/bench_n2_peer2.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_n2_peer2.json')); print(round(d['ms_per_step'],4), 'eval', round(d['roofline']['kernel_ms'],4), d['parity']['ok_all_ranks'])"
