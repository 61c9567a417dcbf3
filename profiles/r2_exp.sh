#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_window.py tests/test_gpu_train_step.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -3
SNNFLOW_WT_TIMING=1 timeout 300 python profiles/run_window_step.py --reps 2 2>&1 | grep "dgpw" | tail -2 | cut -c1-330
B="timeout 300 python bench.py --steps 30 --warmup 3 --no-eval --no-cpu-baseline"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2_b7_$tag.json 2> gpurun_out/r2_b7_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_b7_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"],4), "ms/step", {k:(v["launches"],v["ms"]) for k,v in list(d["kernels"].items())[:7]})
except Exception as e: print("$tag", "failed", e)
PY
}
run aux A=1
run noaux SNNFLOW_DP_AUX=0
run aux2 A=1
run noaux2 SNNFLOW_DP_AUX=0
