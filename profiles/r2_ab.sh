#!/bin/bash
# sustained (>= 0.5 s timed) A/B of the round-2 kernel changes against the round-1 execution plan, same box, same bench
set -u
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-eval --no-cpu-baseline"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2_ab_$tag.json 2> gpurun_out/r2_ab_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_ab_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"],4), "ms/step", round(d["value"],1), "samples/s", d["gpu_launches_per_step"], "launches", d["clocks"]["sm_mhz"], "MHz")
except Exception as e: print("$tag", "failed", e)
PY
}
run r2_default A=1
run r1_plan SNNFLOW_FWD_PERSIST=0 SNNFLOW_RB_FUSE=0 SNNFLOW_RB_PERSIST=0 SNNFLOW_L2_PREFETCH=0
run r2_default_b A=1
run r1_plan_b SNNFLOW_FWD_PERSIST=0 SNNFLOW_RB_FUSE=0 SNNFLOW_RB_PERSIST=0 SNNFLOW_L2_PREFETCH=0
run no_fwd_persist SNNFLOW_FWD_PERSIST=0
run no_rb_persist SNNFLOW_RB_PERSIST=0
