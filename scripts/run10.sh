mkdir -p gpurun_out
echo "== default"; timeout 120 python scripts/time_eval.py 65536 256 f16x3 4 600
echo "== quad"; DES_LIB_PATH=$PWD/distributedes_b200/libdes_b200_quad.so timeout 120 python scripts/time_eval.py 65536 256 f16x3 4 0
echo "== H=128"; timeout 120 python scripts/time_eval.py 16384 128 f16x3 3 300
echo "== cma"; timeout 300 python scripts/time_cma.py 2>&1 | tail -1
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=6 --tb=short 2>&1 | grep -E "^E  |assert|passed|failed|^FAILED" | head -30
echo "== ncu eval"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_pair_kernel -s 2 -c 1 -o gpurun_out/prof_r2_eval_pair -f python scripts/time_eval.py 65536 256 f16x3 1 0 > gpurun_out/prof_r2_eval_pair.log 2>&1; echo rc=$?
