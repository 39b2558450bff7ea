mkdir -p gpurun_out
T="timeout 120 python scripts/time_eval.py 65536 256 f16x3 4 0"
for v in bvh bal g8 g8b; do echo "== $v"; DES_LIB_PATH=$PWD/distributedes_b200/libdes_b200_$v.so $T; done
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo rc=$?; tail -3 gpurun_out/bench_r2_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n1.json'))
print({k:d[k] for k in ['value','ms_per_step','e2e','parity','clocks']})
for c in d['configs']: print(json.dumps({k:c.get(k) for k in ['workload','ms_per_step','value','generation_ms','rank_mu_update_ms','updates_per_sec','error']}))
PY
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-600
