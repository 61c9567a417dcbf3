#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_window.py tests/test_gpu_network.py tests/test_gpu_train_step.py -m gpu -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -3
SNNFLOW_WT_TIMING=1 python profiles/run_window_step.py --reps 2 2>&1 | grep "dgpw" | tail -2 | cut -c1-330
B="python bench.py --steps 30 --warmup 3 --no-eval --no-cpu-baseline"
for i in 1 2; do
$B > gpurun_out/r2_b4_default.json 2> gpurun_out/r2_b4_default.err; echo "default rc=$?"
python - <<'PY'
import json
for n in ("default",):
    try:
        d=json.loads(open(f"gpurun_out/r2_b4_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],4), "ms/step", {k:(v["launches"],v["ms"]) for k,v in list(d["kernels"].items())[:7]})
    except Exception as e: print(n, "failed", e)
PY
done
