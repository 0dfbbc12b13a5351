#!/bin/bash
# Same-box A/B of library builds: tools/ab/lib_<NAME>.so (built locally with EXTRA=-D..., not tracked) are copied over
# the package's libafe_cuda.so one after the other and the device-resident bench line is taken for each (timing only:
# experiment builds may compute wrong features). Usage under gpurun: bash tools/gpu_ablibs.sh <tag> "<bench flags>" NAME...
tag=$1; flags=$2; shift 2
mkdir -p gpurun_out
cp asr-featext-opencl_b200/libafe_cuda.so /tmp/lib_orig.so
for rep in 1 2; do
for name in "$@"; do
  cp tools/ab/lib_$name.so asr-featext-opencl_b200/libafe_cuda.so
  timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-extras $flags > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_$name.json").read().strip().splitlines()[-1])
    print("%-12s rep $rep  %.3f ms  frac %.4f  %s  clk %s" % ("$name", d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$name: no line:", e)
PY
done
done
cp /tmp/lib_orig.so asr-featext-opencl_b200/libafe_cuda.so
