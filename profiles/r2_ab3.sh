#!/bin/bash
set -u
for env in "SNNFLOW_FWD_PERSIST=1" "SNNFLOW_FWD_PERSIST=1 SNNFLOW_PRODUCERS=2" ""; do
  echo "== $env"
  env $env python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eval > gpurun_out/r2_ab3.json 2>gpurun_out/r2_ab3.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_ab3.json').read().strip().splitlines()[-1])
print('train', round(d['value'],1), round(d['ms_per_step'],4), d['gpu_launches_per_step'], {k:(v['launches'], v['ms']) for k,v in list(d['kernels'].items())[:6]})
PY
done
