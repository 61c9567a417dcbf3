#!/bin/bash
# ncu --set full of the seven layer launches of one eval window (LIFFireFlowNet 256x256 batch 16, T = 10, no_grad)
set -u
mkdir -p gpurun_out
CMD="python profiles/run_window_step.py --kind LIFFireFlowNet --res 256 --batch 16 --eval --reps 2"
$CMD > gpurun_out/r2_eval_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:"wt_fwd|pred_fwd|pack_planes" -s 9 -c 9 -o gpurun_out/prof_r2_eval $CMD > gpurun_out/r2_ncu_eval.log 2>&1
ncu -i gpurun_out/prof_r2_eval.ncu-rep --page raw --csv > gpurun_out/prof_r2_eval_raw.csv 2>/dev/null
rm -f gpurun_out/prof_r2_eval.ncu-rep
ls -la gpurun_out/prof_r2_eval_raw.csv; tail -2 gpurun_out/r2_ncu_eval.log
