# Scaling lines of one 8-GPU box: bench.py at N = 8, 4, 2, 1 (peer-memory exchange), headline workload only.
mkdir -p gpurun_out
for N in 8 4 2 1; do
  echo "== bench $N gpus"
  if [ "$N" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-other-modes --no-configs > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err; echo rc=$?
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) bench.py --gpus $N --steps 10 --warmup 3 --no-other-modes --no-configs > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err; echo rc=$?
  fi
done
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.loads([l for l in open('gpurun_out/r2_scale_n%d.json'%n) if l.startswith('{')][-1]); print(n, round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['roofline']['kernel_ms'], (d.get('parity') or {}).get('ok_all_ranks'), d['clocks'])
    except Exception as e: print(n, e)
PY
