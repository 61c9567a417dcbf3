#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_window.py tests/test_gpu_train_step.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eval > gpurun_out/r2_pw_bench.json 2>gpurun_out/r2_pw_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_pw_bench.json').read().strip().splitlines()[-1])
print('train', d['value'], d['ms_per_step'], d['gpu_launches_per_step'])
for k,v in list(d['kernels'].items())[:7]: print(' ', k, v['launches'], v['ms'])
PY
