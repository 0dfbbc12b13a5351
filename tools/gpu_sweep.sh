#!/bin/bash
# parity tests, then a short device-only bench per (warps, tile) setting
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
for w in ${WARPS:-8}; do for tile in ${TILES:-352}; do
  AFE_FUSED_WARPS=$w AFE_TILE_FRAMES=$tile timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu $BENCH_ARGS > gpurun_out/sweep_${w}_${tile}.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sweep_${w}_${tile}.log").read().strip().splitlines()[-1])
    print("warps ${w} tile ${tile}: step %.2f ms  K1 %.2f ms  %.0f Mframes/s  tiles %d  corpus %.2f ms" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["value"]/1e6, d["tiles_per_gpu"], d["corpus_cmvn"].get("ms_per_step", -1)))
except Exception as e:
    print("warps ${w} tile ${tile}: failed", e); print(open("gpurun_out/sweep_${w}_${tile}.log").read()[-1500:])
PY
done; done
