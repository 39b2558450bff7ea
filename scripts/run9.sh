mkdir -p gpurun_out
T="timeout 120 python scripts/time_eval.py 65536 256 f16x3 4 0"
echo "== base(unbal)"; timeout 120 python scripts/time_eval.py 65536 256 f16x3 4 600
for v in prcp g8 g8p; do echo "== $v"; DES_LIB_PATH=$PWD/distributedes_b200/libdes_b200_$v.so $T; done
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_host_surface.py -q --tb=short 2>&1 | grep -E "^E|assert|passed|failed" | head -30
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo rc=$?; tail -3 gpurun_out/bench_r2_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n1.json'))
print({k:d[k] for k in ['value','ms_per_step','e2e','parity','clocks']})
print(json.dumps(d['roofline'])[:900])
for c in d['configs']: print(json.dumps({k:c.get(k) for k in ['workload','ms_per_step','value','generation_ms','rank_mu_update_ms','updates_per_sec','error']}), (c.get('cpu_baseline') or {}).get('value'), (c.get('cpu_baseline') or {}).get('kind'))
print(d['other_modes'], d['closed_loop'])
PY
