#!/bin/bash
# Profile capture on a B200 box (run through gpurun from the repo root):  bash profiles/capture.sh <tag>
#   1. the benchmark command runs plainly (must exit 0),
#   2. ncu launch list of the same command  -> gpurun_out/launches_<tag>.csv   (every launch with its device time),
#   3. ncu --set full of the window-engine kernels of the second profiled step -> gpurun_out/prof_<tag>_{fwd,bwd}.ncu-rep
# Summaries are produced here with profiles/summarize.py and committed under profiles/.
set -u
TAG=${1:-rX}
CMD="python bench.py --steps 1 --warmup 1 --no-graph --no-eval --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
# regex-filtered launch index inside one window step (63 launches): 11 = feed-forward layer forward (all T bins),
# 13.. = recurrent layer forward steps, 25 = pointwise BPTT, 26 = data gradient, 27 = weight gradient, 31.. = recurrent BPTT steps
ncu --set full --clock-control none --import-source on -k regex:"wt_|wg_|pw_seq" -s 74 -c 4 -o gpurun_out/prof_${TAG}_fwd $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"wt_|wg_|pw_seq" -s 88 -c 8 -o gpurun_out/prof_${TAG}_bwd $CMD > gpurun_out/ncu_b_$TAG.log 2>&1
ls -la gpurun_out/
