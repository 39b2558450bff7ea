#!/bin/bash
# One GPU round: probe, parity tests, smoke, bench lines.  Every step has its own timeout.
mkdir -p gpurun_out
if [ -x scripts/probe/umma_probe ]; then timeout 120 scripts/probe/umma_probe 2>&1 | tail -12; fi
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -6
for cfg in $BENCH_CFGS; do
  tag=$(echo "$cfg" | tr -d ' -' | tr ',' '_'); args=$(echo "$cfg" | tr ',' ' ')
  timeout 600 python bench.py --steps 5 --warmup 3 $args > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench[$tag] rc=$?"
  cat gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_$tag.err
done
