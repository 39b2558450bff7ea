mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_pair_kernel -s 2 -c 1 -o gpurun_out/prof_r2_pair -f python scripts/time_eval.py 65536 256 f16x3 1 0 > gpurun_out/prof_r2_pair.log 2>&1; echo rc=$?
tail -3 gpurun_out/prof_r2_pair.log
ls -la gpurun_out/*.ncu-rep
