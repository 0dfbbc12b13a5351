for sk in 0 1 2 4 6 7; do AFE_DEBUG_SKIP=$sk timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/skip_$sk.log 2>&1; python -c "
import json
d=json.loads(open('gpurun_out/skip_$sk.log').read().strip().splitlines()[-1]); print('skip $sk: %.2f ms' % d['ms_per_step'])"; done
