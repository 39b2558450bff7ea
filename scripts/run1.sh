mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
echo "== small correctness v2"; timeout 120 python scripts/time_eval.py 20 256 f16x3 2 20; echo rc=$?
timeout 120 python scripts/time_eval.py 300 128 f16x3 2 300; echo rc=$?
echo "== v2 full"; timeout 180 python scripts/time_eval.py 65536 256 f16x3 5 600; echo rc=$?
echo "== old"; DES_TC_PAIR_V2=0 timeout 180 python scripts/time_eval.py 65536 256 f16x3 5 600; echo rc=$?
echo "== grad"; timeout 120 python scripts/time_grad.py 65536 256 5; echo rc=$?
echo "== pytest subset"; timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py -q -x 2>&1 | tail -8
