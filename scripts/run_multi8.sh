mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "== bench 8 gpus (peer comm)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 10 --warmup 3 --no-other-modes > gpurun_out/bench_n8_peer.json 2> gpurun_out/bench_n8_peer.err; echo rc=$?; tail -3 gpurun_out/bench_n8_peer.err
echo "== bench 8 gpus (nccl)"; DES_COMM=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 8 --steps 10 --warmup 3 --no-other-modes --no-configs > gpurun_out/bench_n8_nccl.json 2> gpurun_out/bench_n8_nccl.err; echo rc=$?
echo "== bench 4 gpus (peer comm)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29623 bench.py --gpus 4 --steps 10 --warmup 3 --no-other-modes --no-configs > gpurun_out/bench_n4_peer.json 2> gpurun_out/bench_n4_peer.err; echo rc=$?
python - <<'PY'
import json
for f in ['gpurun_out/bench_n8_peer.json','gpurun_out/bench_n8_nccl.json','gpurun_out/bench_n4_peer.json']:
    try:
        d=json.load(open(f)); print(f, round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['config']['cuda_graph'], d['parity'].get('ok_all_ranks'), d['roofline']['kernel_ms'])
        for c in d.get('configs',[]): print('   ', c.get('workload','')[:70], c.get('ms_per_step'), c.get('rank_mu_update_ms'), c.get('error'))
    except Exception as e: print(f, e)
PY
