#!/bin/bash
# tail-kernel round: affected tests, then the bench line
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_window.py tests/test_gpu_train_step.py tests/test_gpu_network.py tests/test_gpu_encode_iwe.py tests/test_gpu_engine.py -m gpu -x -q > gpurun_out/r2_tail_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_tail_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_tail_bench.json 2> gpurun_out/r2_tail_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_tail_bench.json').read().strip().splitlines()[-1])
print('train', d['value'], d['ms_per_step'], d['gpu_launches_per_step'])
for k,v in d['kernels'].items(): print(' ', k, v['launches'], v['ms'])
print('eval', d['eval']['value'], d['eval'].get('per_bin_forward'), d['eval'].get('e2e'))
print('cfg0', d.get('eval_cfg0'))
PY
