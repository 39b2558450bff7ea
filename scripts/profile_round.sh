#!/bin/bash
# ncu: launch list + full capture of the top kernels.  args: pop hidden precision tag
POP=${1:-16384}; H=${2:-256}; PREC=${3:-f16}; TAG=${4:-r1}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}_h${H}_${PREC}.csv python scripts/profile_gen.py $POP $H $PREC 3 > gpurun_out/prof_launch.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:eval_(pair|tc)_kernel" -s 1 -c 1 -o gpurun_out/prof_${TAG}_eval_h${H}_${PREC} -f python scripts/profile_gen.py $POP $H $PREC 2 > gpurun_out/prof_full1.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grad_chunk_kernel -s 1 -c 1 -o gpurun_out/prof_${TAG}_grad_h${H} -f python scripts/profile_gen.py $POP $H $PREC 2 > gpurun_out/prof_full2.log 2>&1; echo rc=$?
ls -la gpurun_out | tail -6
