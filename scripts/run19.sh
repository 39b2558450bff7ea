mkdir -p gpurun_out
echo "== tests (ops, goldens)"; timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_goldens.py -m gpu -q --tb=short 2>&1 | tail -6
echo "== base"; timeout 120 python scripts/time_eval.py 65536 256 f16x3 4 0
for v in bo100 bo400 quad quadbo; do echo "== $v"; DES_LIB_PATH=distributedes_b200/libdes_b200_$v.so timeout 120 python scripts/time_eval.py 65536 256 f16x3 4 0; done
echo "== grad"; timeout 120 python scripts/time_grad.py 65536 256 5; timeout 120 python scripts/time_grad.py 4096 64 10
echo "== pop4096 launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_h64_pop4096b.csv python scripts/profile_gen.py 4096 64 f16x3 3 > gpurun_out/prof_launch64.log 2>&1; echo rc=$?
