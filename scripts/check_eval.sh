# quick check of the forward kernels after a change: timing + parity at the three hidden widths, then the parity tests
echo "== H=256"; timeout 120 python scripts/time_eval.py 65536 256 f16x3 4 600
echo "== H=128"; timeout 120 python scripts/time_eval.py 333 128 f16x3 2 333; timeout 120 python scripts/time_eval.py 16384 128 f16x3 3 0
echo "== H=64"; timeout 120 python scripts/time_eval.py 301 64 f16x3 2 301; timeout 120 python scripts/time_eval.py 4096 64 f16x3 5 600
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_goldens.py tests/test_gpu_fullsize.py tests/test_gpu_host_surface.py -m gpu -q --tb=short 2>&1 | tail -4
