#!/bin/bash
# full GPU suite + smoke + bench line
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_full_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_full_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_full_bench.json 2> gpurun_out/r2_full_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_full_bench.json').read().strip().splitlines()[-1])
print('train', d['value'], d['ms_per_step'], d['gpu_launches_per_step'], 'e2e', d['e2e']['value'])
print('roofline', d['roofline'])
print('cpu', d['cpu_baseline'])
print('eval', d['eval']['value'], d['eval'].get('per_bin_forward'), d['eval'].get('e2e')['value'], d['eval'].get('roofline'))
print('cfg0', d.get('eval_cfg0'))
PY
