mkdir -p gpurun_out
echo "== rank tests"; timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py -m gpu -q -k "rank" --tb=short 2>&1 | tail -15
echo "== rank timing (default)"; timeout 120 python scripts/time_rank.py
echo "== rank timing (bucket path from 2048)"; DES_RANK_BUCKET_MIN=2048 timeout 120 python scripts/time_rank.py
echo "== ncu full: pair kernel pop 16384"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_pair_kernel -s 1 -c 1 -o gpurun_out/prof_r2_pair_final -f python scripts/profile_gen.py 16384 256 f16x3 2 > gpurun_out/prof_full_pair.log 2>&1; echo rc=$?
echo "== shard launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_shard8b.csv python scripts/profile_shard.py 65536 8192 256 3 > gpurun_out/prof_shard.log 2>&1; echo rc=$?
