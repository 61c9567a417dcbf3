#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_window.py tests/test_gpu_network.py tests/test_gpu_engine.py tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/r2_col_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_col_tests.log
for env in "" "SNNFLOW_STREAM_STEP=0" "SNNFLOW_COL_TILES=0" "SNNFLOW_COL_TILES=0 SNNFLOW_STREAM_STEP=0"; do
  echo "== $env"
  env $env python profiles/run_stream_forward.py | head -3
  env $env python profiles/run_stream_forward.py --kind LIFFireNet | head -4
  env $env python profiles/run_window_step.py --kind LIFFireFlowNet --res 256 --batch 16 --eval --reps 3 --time 20 2>&1 | tail -2
done
