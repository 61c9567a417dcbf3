#!/bin/bash
# full bench line + ncu DRAM-traffic passes for the three workload sections (train step, eval window, encode / IWE)
set -u
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
C1="python bench.py --steps 1 --warmup 1 --no-graph --no-eval --no-cpu-baseline"
$C1 > gpurun_out/r2_t_plain1.log 2>&1 && ncu $M --log-file gpurun_out/traffic_train.csv $C1 > gpurun_out/r2_t_ncu1.log 2>&1; echo "train rc=$?"
C2="python profiles/run_window_step.py --kind LIFFireFlowNet --res 256 --batch 16 --eval --reps 2"
$C2 > gpurun_out/r2_t_plain2.log 2>&1 && ncu $M --log-file gpurun_out/traffic_eval.csv $C2 > gpurun_out/r2_t_ncu2.log 2>&1; echo "eval rc=$?"
C3="python profiles/run_encode_iwe.py"
$C3 > gpurun_out/r2_t_plain3.log 2>&1 && ncu $M --log-file gpurun_out/traffic_micro.csv $C3 > gpurun_out/r2_t_ncu3.log 2>&1; echo "micro rc=$?"
ls -la gpurun_out/traffic_*.csv
