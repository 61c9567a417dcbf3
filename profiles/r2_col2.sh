#!/bin/bash
set -u
for env in "" "SNNFLOW_L2_PREFETCH=0"; do
  echo "== $env"
  env $env python profiles/run_stream_forward.py | head -3
  env $env python profiles/run_stream_forward.py --kind LIFFireNet | head -4
done
timeout 600 python -m pytest tests/test_gpu_network.py -m gpu -x -q 2>&1 | tail -2
