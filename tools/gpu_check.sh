#!/bin/bash
# parity tests, then the device-only bench: default kernel and the opt-in warp-specialised kernel, same box
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
for v in default nocluster ws; do
  extra=""; [ $v = ws ] && extra="--ws-kernel"; [ $v = nocluster ] && extra="--no-cluster"
  timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu $extra $BENCH_ARGS > gpurun_out/check_$v.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/check_$v.log").read().strip().splitlines()[-1])
    print("$v: step %.3f ms  K1 %.3f ms  %.0f Mframes/s  tiles %d  corpus %.2f ms  kernel %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["value"]/1e6, d["tiles_per_gpu"], d["corpus_cmvn"].get("ms_per_step", -1), d["roofline"]["kernel"]))
except Exception as e:
    print("$v: failed", e); print(open("gpurun_out/check_$v.log").read()[-1500:])
PY
done
