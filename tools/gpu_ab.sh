#!/bin/bash
# A/B of kernel variants on one B200: parity tests first, then the device-resident bench line per variant.
# Usage under gpurun: bash tools/gpu_ab.sh <tag> ["bench flags of variant A" "bench flags of variant B" ...]
tag=${1:-ab}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu -k "shape or batch or golden or config" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${tag}_pytest.log
i=0
for flags in "$@"; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-extras $flags > gpurun_out/${tag}_bench_$i.json 2> gpurun_out/${tag}_bench_$i.err
  echo "variant $i [$flags] exit $?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_$i.json").read().strip().splitlines()[-1])
    print("   ms_per_step", d["ms_per_step"], "frac", d["roofline"]["frac"], d["roofline"]["kernel"], "parity", (d.get("parity") or {}).get("ok"), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("   no line:", e)
PY
  i=$((i+1))
done
