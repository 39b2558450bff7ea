mkdir -p gpurun_out
T="timeout 120 python scripts/time_eval.py 65536 256 f16x3 4 0"
echo "== base correctness"; timeout 120 python scripts/time_eval.py 65536 256 f16x3 4 600
timeout 120 python scripts/time_eval.py 300 128 f16x3 2 300
for v in g8 unbal bo2 prcp; do echo "== $v"; DES_LIB_PATH=$PWD/distributedes_b200/libdes_b200_$v.so $T; done
echo "== trace"; DES_PAIR_TRACE=1 DES_LIB_PATH=$PWD/distributedes_b200/libdes_b200_trace.so timeout 180 python scripts/time_eval.py 65536 256 f16x3 1 0 2>&1 | grep -E "TRACE" | grep -E "m17|m18"
echo "== old kernel shapes"; timeout 120 python scripts/time_eval.py 4096 64 f16x3 5 300; timeout 120 python scripts/time_eval.py 65536 256 f16 5 300; DES_TC_PAIR_V2=0 timeout 120 python scripts/time_eval.py 65536 256 f16x3 3 300
echo "== cma"; timeout 300 python scripts/time_cma.py 2>&1 | tail -1
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=6 --tb=short 2>&1 | tail -12
