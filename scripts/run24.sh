mkdir -p gpurun_out
for v in traceA traceB; do
echo "== $v"; DES_PAIR_TRACE=1 DES_LIB_PATH=$PWD/distributedes_b200/libdes_b200_$v.so timeout 180 python scripts/time_eval.py 65536 256 f16x3 2 0 2>&1 | grep -E "TRACE|pop" | grep -E "m17|m18|m19|pop" | head -12
done
