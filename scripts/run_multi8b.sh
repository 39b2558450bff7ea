mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() { # N tag extra-env extra-flags
  N=$1; TAG=$2; shift 2
  echo "== bench $N gpus ($TAG)"
  if [ "$N" = 1 ]; then
    env $ENVX timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 "$@" > gpurun_out/r2_bench_n${N}_$TAG.json 2> gpurun_out/r2_bench_n${N}_$TAG.err; echo rc=$?
  else
    env $ENVX timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N+RANDOM%50)) bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/r2_bench_n${N}_$TAG.json 2> gpurun_out/r2_bench_n${N}_$TAG.err; echo rc=$?
  fi
  tail -2 gpurun_out/r2_bench_n${N}_$TAG.err
}
ENVX="DES_COMM=peer" run 8 peer --no-other-modes
ENVX="DES_COMM=nccl" run 8 nccl --no-other-modes --no-configs
ENVX="DES_COMM=peer" run 4 peer --no-other-modes --no-configs
ENVX="DES_COMM=peer" run 2 peer --no-other-modes --no-configs
ENVX="DES_COMM=peer" run 1 samebox --no-other-modes --no-configs
echo "== multi-gpu tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short 2>&1 | tail -3
python - <<'PY'
import json
for f in ['r2_bench_n8_peer','r2_bench_n8_nccl','r2_bench_n4_peer','r2_bench_n2_peer','r2_bench_n1_samebox']:
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1]); print(f, round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['config']['cuda_graph'], (d.get('parity') or {}).get('ok_all_ranks'), d['roofline']['kernel_ms'])
        for c in d.get('configs',[]): print('   ', c.get('workload','')[:70], c.get('ms_per_step'), c.get('rank_mu_update_ms'), c.get('error'))
    except Exception as e: print(f, e)
PY
