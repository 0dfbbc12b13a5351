#!/bin/bash
# Multi-GPU evidence on one box (gpurun --gpus N): the 2-GPU parity tests (NCCL corpus CMVN, time-sharded stream), the copy
# ceiling of the box and the bench line at N ranks. Usage under gpurun --gpus N: bash tools/gpu_multi.sh N [tag]
N=${1:-2}; tag=${2:-r02_n$N}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu -k "two_gpus or nccl or sharded" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/${tag}_pytest.log
run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $run tools/h2d_d2h_ceiling.py > gpurun_out/${tag}_ceiling.json 2> gpurun_out/${tag}_ceiling.err; echo "ceiling exit $?"
timeout 900 $run bench.py --gpus $N --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference arm exit $?"
timeout 1200 $run bench.py --gpus $N > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
tail -c 600 gpurun_out/${tag}_bench.json
