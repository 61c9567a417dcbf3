#!/bin/bash
# final 1-GPU record of the round: full suite, smoke, bench line (with CPU baselines), reference arm
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_final_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; echo "ref rc=$?"
tail -c 600 gpurun_out/r2_final_ref.json
