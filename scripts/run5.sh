mkdir -p gpurun_out
echo "== trace"; DES_PAIR_TRACE=1 DES_LIB_PATH=$PWD/distributedes_b200/libdes_b200_trace.so timeout 180 python scripts/time_eval.py 65536 256 f16x3 1 0 2>&1 | grep -E "TRACE|pop" | tail -30
echo "== old kernel (remapped warps) f16x3 H=64 pop 4096, f16 H=256"
timeout 120 python scripts/time_eval.py 4096 64 f16x3 5 300; timeout 120 python scripts/time_eval.py 65536 256 f16 5 300
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=6 --tb=short 2>&1 | tail -60
