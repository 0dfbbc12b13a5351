#!/bin/bash
# A/B timing of several builds of libafe_cuda.so on ONE box (box-to-box clock differences are ~5 %).
# Usage under gpurun: bash tools/gpu_ab.sh <lib_a.so> <lib_b.so> ...   (paths relative to the repo root)
mkdir -p gpurun_out
for rep in 1 2; do for lib in "$@"; do
  AFE_LIB_OVERRIDE=$PWD/$lib timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu > gpurun_out/ab_$(basename $lib).log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$(basename $lib).log").read().strip().splitlines()[-1])
    print("$lib rep $rep: step %.3f ms  %.0f Mframes/s  sm %s MHz" % (d["ms_per_step"], d["value"]/1e6, d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$lib: failed", e); print(open("gpurun_out/ab_$(basename $lib).log").read()[-800:])
PY
done; done
