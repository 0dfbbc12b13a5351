#!/bin/bash
# ncu on the fused kernel only (after a plain run of the same command exits 0). Usage: bash tools/gpu_prof.sh <tag> [extra bench args]
tag=${1:-prof}; shift
mkdir -p gpurun_out
cmd="env AFE_FUSED_SHAPE=${AFE_FUSED_SHAPE:-8x4} python bench.py --steps 2 --warmup 3 --utts 2000 --no-e2e --no-cpu $*"
timeout 600 $cmd > gpurun_out/${tag}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_ -s 3 -c 1 -o gpurun_out/${tag} $cmd > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/${tag}_ncu.log
