# Final 1-GPU check of a build: GPU tests, smoke, default bench line.  bash scripts/final_check.sh
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 --tb=short 2>&1 | grep -E "^E  |assert|passed|failed|^FAILED" | head -30
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench n=1"; timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo rc=$?; tail -3 gpurun_out/r2_bench_n1.err
