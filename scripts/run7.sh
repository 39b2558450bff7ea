mkdir -p gpurun_out
T="timeout 120 python scripts/time_eval.py 65536 256 f16x3 4 0"
echo "== base"; $T
for s in 4 6; do echo "== base slots=$s"; DES_PAIR_SLOTS=$s $T; done
for v in bo prcp unbal unbalbo; do echo "== $v"; DES_LIB_PATH=$PWD/distributedes_b200/libdes_b200_$v.so $T; done
echo "== bo slots=5"; DES_PAIR_SLOTS=5 DES_LIB_PATH=$PWD/distributedes_b200/libdes_b200_bo.so $T
echo "== trace"; DES_PAIR_TRACE=1 DES_LIB_PATH=$PWD/distributedes_b200/libdes_b200_trace.so timeout 180 python scripts/time_eval.py 65536 256 f16x3 1 0 2>&1 | grep -E "TRACE" | grep -E "m17|m18"
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_host_surface.py tests/test_gpu_ops.py -q --maxfail=6 --tb=short 2>&1 | tail -8
