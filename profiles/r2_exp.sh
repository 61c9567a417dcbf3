#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_window.py tests/test_gpu_network.py tests/test_gpu_train_step.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -3
SNNFLOW_WT_TIMING=1 timeout 300 python profiles/run_window_step.py --reps 2 2>&1 | grep "recbwd" | tail -2 | cut -c1-330
SNNFLOW_RB_R=1 SNNFLOW_WT_TIMING=1 timeout 300 python profiles/run_window_step.py --reps 2 2>&1 | grep "recbwd" | tail -2 | cut -c1-330
B="timeout 300 python bench.py --steps 30 --warmup 3 --no-eval --no-cpu-baseline"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2_b6_$tag.json 2> gpurun_out/r2_b6_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_b6_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"],4), "ms/step", {k:(v["launches"],v["ms"]) for k,v in list(d["kernels"].items())[:7]})
except Exception as e: print("$tag", "failed", e)
PY
}
run persist A=1
run persist_r1 SNNFLOW_RB_R=1
run perbin SNNFLOW_RB_PERSIST=0
run persist_b A=1
run persist_r1_b SNNFLOW_RB_R=1
