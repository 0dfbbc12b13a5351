#!/bin/bash
# parity tests with the default shape, then a short device-only bench for each kernel shape / tile size
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
for shape in 8x4 4x8; do for tile in 264; do
  AFE_FUSED_SHAPE=$shape AFE_TILE_FRAMES=$tile timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/sweep_${shape}_${tile}.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sweep_${shape}_${tile}.log").read().strip().splitlines()[-1])
    print("shape ${shape} tile ${tile}: step %.2f ms  K1 %.2f ms  %.0f Mframes/s  tiles %d" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["value"]/1e6, d["tiles_per_gpu"]))
except Exception as e:
    print("shape ${shape} tile ${tile}: failed", e)
PY
done; done
