#!/bin/bash
# End-of-iteration evidence on one B200 box: parity tests, smoke, reference arm, full bench (N=1), then (only after the
# plain command exited 0) the ncu launch list and ONE full profile of K1 on a 2000-utterance launch.
# Usage under gpurun:  bash tools/gpu_final.sh [tag]
tag=${1:-r02_final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,driver_version --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference arm exit $?"
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; rc=$?; echo "bench exit $rc"
cmd="python bench.py --steps 3 --warmup 3 --utts 2000 --no-e2e --no-cpu --no-extras"
if [ $rc -eq 0 ]; then
  timeout 600 $cmd > gpurun_out/${tag}_plain_small.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_|nccl|Kernel" -c 80 --csv --log-file gpurun_out/${tag}_launches.csv $cmd > gpurun_out/${tag}_ncu_launches.log 2>&1
  echo "ncu launches exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_ -s 3 -c 1 -o gpurun_out/${tag}_k_fused $cmd > gpurun_out/${tag}_ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
