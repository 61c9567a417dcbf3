#!/bin/bash
# Round-2 profile capture on a B200 box (through gpurun, from the repo root):  bash profiles/capture_r2.sh
#   1. ncu metrics pass (time + DRAM bytes of every launch) of the train step / the eval window / the encode + IWE set
#      -> gpurun_out/traffic_{train,eval,micro}.csv  -> profiles/traffic.json, profiles/r2_launches_*.md (make_traffic.py)
#   2. ncu --set full of one whole training window of tensor-core kernels -> gpurun_out/prof_r2_window_raw.csv
#      -> profiles/r2_kernels_window_step.md (summarize.py)
# Every command runs plainly first (must exit 0) before it runs under ncu.
set -u
mkdir -p gpurun_out
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
C1="python bench.py --steps 1 --warmup 1 --min-seconds 0 --no-graph --no-eval --no-cpu-baseline"
$C1 > gpurun_out/r2_t_plain1.log 2>&1 && ncu $M --log-file gpurun_out/traffic_train.csv $C1 > gpurun_out/r2_t_ncu1.log 2>&1; echo "train rc=$?"
C2="python profiles/run_window_step.py --kind LIFFireFlowNet --res 256 --batch 16 --eval --reps 2"
$C2 > gpurun_out/r2_t_plain2.log 2>&1 && ncu $M --log-file gpurun_out/traffic_eval.csv $C2 > gpurun_out/r2_t_ncu2.log 2>&1; echo "eval rc=$?"
C3="python profiles/run_encode_iwe.py"
$C3 > gpurun_out/r2_t_plain3.log 2>&1 && ncu $M --log-file gpurun_out/traffic_micro.csv $C3 > gpurun_out/r2_t_ncu3.log 2>&1; echo "micro rc=$?"
C4="python profiles/run_window_step.py --reps 2"
$C4 > gpurun_out/r2_t_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"wt_|wg_|pw_seq|win_reduce" -s 40 -c 40 -o gpurun_out/prof_r2_window $C4 > gpurun_out/r2_t_ncu4.log 2>&1; echo "window rc=$?"
ncu -i gpurun_out/prof_r2_window.ncu-rep --page raw --csv > gpurun_out/prof_r2_window_raw.csv 2>/dev/null
[ -f gpurun_out/prof_r2_window.ncu-rep ] && [ $(stat -c %s gpurun_out/prof_r2_window.ncu-rep) -gt 40000000 ] && rm -f gpurun_out/prof_r2_window.ncu-rep
ls -la gpurun_out/traffic_*.csv gpurun_out/prof_r2_window*
