#!/bin/bash
set -u
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --min-seconds 0.2 > gpurun_out/r2_g_bench.json 2>gpurun_out/r2_g_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_g_bench.json').read().strip().splitlines()[-1])
print('train', d['value'], d['ms_per_step'])
print('eval', d['eval']['value'], d['eval'].get('per_bin_forward'), d['eval'].get('e2e')['value'])
for k,v in d['eval']['kernels'].items(): print('  ', k, v)
print('cfg0', d.get('eval_cfg0'))
PY
