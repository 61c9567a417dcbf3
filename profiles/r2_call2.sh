#!/bin/bash
# round 2, GPU call 2: time-fused recurrent forward (per-tile flags) + L2 prefetch hints: tests, A/B bench lines, role timers
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_window.py tests/test_gpu_train_step.py tests/test_gpu_reference_seam.py tests/test_gpu_network.py tests/test_gpu_engine.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2_pytest2.log
B="python bench.py --steps 30 --warmup 3 --no-eval --no-cpu-baseline"
$B > gpurun_out/r2_b2_default.json 2> gpurun_out/r2_b2_default.err; echo "default rc=$?"
SNNFLOW_FWD_PERSIST=0 $B > gpurun_out/r2_b2_nopersist.json 2>/dev/null; echo "nopersist rc=$?"
SNNFLOW_L2_PREFETCH=0 $B > gpurun_out/r2_b2_nopf.json 2>/dev/null; echo "nopf rc=$?"
SNNFLOW_WT_TIMING=1 python profiles/run_window_step.py --reps 2 > gpurun_out/r2_roletimer2.log 2>&1; echo "roletimer rc=$?"
python - <<'PY'
import json
for n in ("default","nopersist","nopf"):
    try:
        d=json.loads(open(f"gpurun_out/r2_b2_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],4), "ms/step", {k:(v["launches"],v["ms"]) for k,v in list(d["kernels"].items())[:7]})
    except Exception as e: print(n, "failed", e)
PY
