mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 --tb=short 2>&1 | grep -E "^E  |assert|passed|failed|^FAILED" | head -30
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench n=1"; timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo rc=$?; tail -3 gpurun_out/r2_bench_n1.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo rc=$?; cat gpurun_out/r2_bench_ref.json | cut -c1-600
echo "== launch lists"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_h256_f16x3_pop65536.csv python scripts/profile_gen.py 65536 256 f16x3 3 > gpurun_out/prof_launch.log 2>&1; echo rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_h64_pop4096.csv python scripts/profile_gen.py 4096 64 f16x3 3 > gpurun_out/prof_launch64.log 2>&1; echo rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_shard_8192_of_65536.csv python scripts/profile_shard.py 65536 8192 256 3 > gpurun_out/prof_shard.log 2>&1; echo rc=$?
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_pair_kernel -s 1 -c 1 -o gpurun_out/prof_r2_final_eval_pair -f python scripts/profile_gen.py 65536 256 f16x3 2 > gpurun_out/prof_full1.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grad_chunk_kernel -s 1 -c 1 -o gpurun_out/prof_r2_final_grad -f python scripts/profile_gen.py 65536 256 f16x3 2 > gpurun_out/prof_full2.log 2>&1; echo rc=$?
echo "== cma timing"; timeout 300 python scripts/time_cma.py
echo "== rank timing"; timeout 120 python scripts/time_rank.py
