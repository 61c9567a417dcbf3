#!/bin/bash
# ncu --set full with source correlation of one steady-state wt_dgpw and one wt_recbwd launch of the training window
set -u
mkdir -p gpurun_out
CMD="python profiles/run_window_step.py --reps 2"
$CMD > gpurun_out/r2_src_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wt_dgpw -s 4 -c 1 -o gpurun_out/prof_r2_dgpw $CMD > gpurun_out/r2_ncu_dgpw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wt_recbwd -s 22 -c 1 -o gpurun_out/prof_r2_recbwd $CMD > gpurun_out/r2_ncu_recbwd.log 2>&1
ls -la gpurun_out/*.ncu-rep
