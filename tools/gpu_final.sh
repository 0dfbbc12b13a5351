#!/bin/bash
# End-of-iteration evidence: parity tests, smoke, full bench (N=1), reference arm, ncu launch list + full profile of K1.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,driver_version --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err; echo "reference arm exit $?"
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; rc=$?; echo "bench exit $rc"
cmd="python bench.py --steps 3 --warmup 3 --utts 2000 --no-e2e --no-cpu"
if [ $rc -eq 0 ]; then
  timeout 600 $cmd > gpurun_out/plain_small.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_|nccl|Kernel" -c 60 --csv --log-file gpurun_out/launches.csv $cmd > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_ -s 3 -c 1 -o gpurun_out/prof_final $cmd > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
