#!/bin/bash
# round 2, GPU call 1: parity tests, role timers of the tensor-core pipeline, a bench line, ncu of the encode / IWE kernels
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt 2>&1
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2_pytest1.log
SNNFLOW_WT_TIMING=1 python profiles/run_window_step.py --reps 2 > gpurun_out/r2_roletimer.log 2>&1; echo "roletimer rc=$?"
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench1_ref.json 2> gpurun_out/r2_bench1_ref.err; echo "ref rc=$?"
python profiles/run_encode_iwe.py > gpurun_out/r2_encode_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"encode_|iwe_|flow_gather|wl_" -c 40 -o gpurun_out/prof_r2_encode_iwe python profiles/run_encode_iwe.py > gpurun_out/r2_encode_ncu.log 2>&1
ncu -i gpurun_out/prof_r2_encode_iwe.ncu-rep --page raw --csv > gpurun_out/prof_r2_encode_iwe_raw.csv 2>/dev/null
[ -f gpurun_out/prof_r2_encode_iwe.ncu-rep ] && [ $(stat -c %s gpurun_out/prof_r2_encode_iwe.ncu-rep) -gt 30000000 ] && rm -f gpurun_out/prof_r2_encode_iwe.ncu-rep
ls -la gpurun_out | tail -15
