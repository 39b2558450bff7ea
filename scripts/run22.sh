mkdir -p gpurun_out
echo "== H=256"; timeout 120 python scripts/time_eval.py 65536 256 f16x3 4 600
echo "== H=128"; timeout 120 python scripts/time_eval.py 333 128 f16x3 2 333
echo "== H=64"; timeout 120 python scripts/time_eval.py 4096 64 f16x3 5 600
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 --tb=short 2>&1 | grep -E "^E  |assert|passed|failed|^FAILED" | head -30
