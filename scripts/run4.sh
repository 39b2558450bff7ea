mkdir -p gpurun_out
echo "== v4 small"; timeout 120 python scripts/time_eval.py 300 256 f16x3 2 300; echo rc=$?
timeout 120 python scripts/time_eval.py 300 128 f16x3 2 300; echo rc=$?
echo "== v4 full"; timeout 180 python scripts/time_eval.py 65536 256 f16x3 5 600; echo rc=$?
echo "== ts3"; DES_LIB_PATH=$PWD/distributedes_b200/libdes_b200_ts3.so timeout 180 python scripts/time_eval.py 65536 256 f16x3 5 0; echo rc=$?
echo "== e112"; DES_LIB_PATH=$PWD/distributedes_b200/libdes_b200_e112.so timeout 180 python scripts/time_eval.py 65536 256 f16x3 5 0; echo rc=$?
echo "== ncu"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_pair_kernel -s 2 -c 1 -o gpurun_out/prof_r2_pair_v4 -f python scripts/time_eval.py 65536 256 f16x3 1 0 > gpurun_out/prof_r2_pair_v4.log 2>&1; echo rc=$?
echo "== gpu tests"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
