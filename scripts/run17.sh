mkdir -p gpurun_out
echo "== shard launch list (N 65536, n_local 8192)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_shard8.csv python scripts/profile_shard.py 65536 8192 256 3 > gpurun_out/prof_shard.log 2>&1; echo rc=$?
echo "== full generation launch list (pop 65536)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_h256_f16x3.csv python scripts/profile_gen.py 65536 256 f16x3 3 > gpurun_out/prof_launch.log 2>&1; echo rc=$?
echo "== pop 4096 H64 launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_h64_pop4096.csv python scripts/profile_gen.py 4096 64 f16x3 3 > gpurun_out/prof_launch64.log 2>&1; echo rc=$?
echo "== eval timings"; timeout 120 python scripts/time_eval.py 65536 256 f16x3 4 0; timeout 120 python scripts/time_eval.py 8192 256 f16x3 10 0;  timeout 120 python scripts/time_eval.py 4096 64 f16x3 10 0
timeout 120 python scripts/time_grad.py 2>&1 | tail -3
timeout 120 python scripts/time_grad.py 8192 256 10 2>&1 | tail -1
