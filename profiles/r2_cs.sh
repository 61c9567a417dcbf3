#!/bin/bash
set -u
for env in "" "SNNFLOW_EXP=4" "" "SNNFLOW_EXP=4"; do
  echo "== $env"
  env $env python profiles/run_window_step.py --kind LIFFireFlowNet --res 256 --batch 16 --eval --reps 3 --time 30 2>&1 | tail -2 | head -1
  env $env python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eval 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train', round(d['value'],1), round(d['ms_per_step'],4), d['kernels']['win_fwd_seq'])"
done
