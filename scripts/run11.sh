mkdir -p gpurun_out
echo "== static ring default"; timeout 120 python scripts/time_eval.py 65536 256 f16x3 4 600
timeout 120 python scripts/time_eval.py 16384 128 f16x3 3 300
echo "== g64"; DES_LIB_PATH=$PWD/distributedes_b200/libdes_b200_g64.so timeout 120 python scripts/time_eval.py 65536 256 f16x3 4 0
echo "== trace"; DES_PAIR_TRACE=1 DES_LIB_PATH=$PWD/distributedes_b200/libdes_b200_trace.so timeout 180 python scripts/time_eval.py 65536 256 f16x3 1 0 2>&1 | grep -E "TRACE" | grep -E "m17|m18" | head -6
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py tests/test_gpu_host_surface.py tests/test_gpu_cma.py -q --maxfail=6 --tb=short 2>&1 | grep -E "^E  |assert|passed|failed|^FAILED" | head -30
