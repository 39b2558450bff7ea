mkdir -p gpurun_out
echo "== v6 small"; timeout 120 python scripts/time_eval.py 300 256 f16x3 2 300; echo rc=$?
timeout 120 python scripts/time_eval.py 300 128 f16x3 2 300; echo rc=$?
echo "== v6 full"; timeout 180 python scripts/time_eval.py 65536 256 f16x3 5 600; echo rc=$?
echo "== backoff"; DES_LIB_PATH=$PWD/distributedes_b200/libdes_b200_bo.so timeout 180 python scripts/time_eval.py 65536 256 f16x3 5 0; echo rc=$?
echo "== trace"; DES_PAIR_TRACE=1 DES_LIB_PATH=$PWD/distributedes_b200/libdes_b200_trace.so timeout 180 python scripts/time_eval.py 65536 256 f16x3 1 0 2>&1 | grep -E "TRACE" | grep -E "m17|m18"
echo "== cma"; timeout 300 python scripts/time_cma.py 2>&1 | tail -2
echo "== cma tests"; timeout 900 python -m pytest tests/test_gpu_cma.py tests/test_gpu_host_surface.py tests/test_gpu_goldens.py -q --maxfail=6 --tb=short 2>&1 | tail -30
