#!/bin/bash
# One GPU-box visit: parity tests, smoke, a short bench, then (only if the plain bench exited 0) the ncu launch list.
# Usage under gpurun:  bash tools/gpu_round.sh [quick|full]
mode=${1:-quick}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,driver_version --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; rc=$?
echo "bench exit $rc"; tail -c 3000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
if [ "$mode" = "full" ] && [ $rc -eq 0 ]; then
  timeout 600 python bench.py --steps 2 --warmup 3 --utts 2000 --no-e2e --no-cpu > gpurun_out/plain_small.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --utts 2000 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_mfcc -s 3 -c 2 -o gpurun_out/prof_fused \
      python bench.py --steps 2 --warmup 3 --utts 2000 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
