#!/bin/bash
set -u
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for env in "" "SNNFLOW_STREAM_STEP=0" "SNNFLOW_STREAM_STEP=0 SNNFLOW_PRODUCERS=1"; do
  echo "== $env"
  env $env python profiles/run_stream_forward.py | head -2
  env $env python profiles/run_stream_forward.py --kind LIFFireNet | head -3
done
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_p3_bench.json 2>gpurun_out/r2_p3_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_p3_bench.json').read().strip().splitlines()[-1])
print('train', d['value'], d['ms_per_step'], d['gpu_launches_per_step'])
for k,v in list(d['kernels'].items())[:7]: print(' ', k, v['launches'], v['ms'])
print('eval', d['eval']['value'], d['eval'].get('per_bin_forward'), d['eval'].get('e2e')['value'])
print('cfg0', d.get('eval_cfg0')['value'])
PY
