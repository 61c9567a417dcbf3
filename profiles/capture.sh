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
# --set full of 12 consecutive window-engine launches of the third host-launched step that cover every kernel kind
# (filtered launch order inside a step, 66 launches: 0 fwd_seq, 1-10 fwd_rec, 11-12 fwd_seq, 13-22 fwd_rec, 23-24 fwd_seq,
#  25 pw_seq, 26 wgrad, 27 reduce, 28 dgrad+pw, 29 wgrad, 30 reduce, 31 dgrad, 32-41 rec_bwd, ...): skip 2*66 + 22
ncu --set full --clock-control none --import-source on -k regex:"wt_|wg_|pw_seq|win_reduce" -s 154 -c 12 -o gpurun_out/prof_${TAG}_step $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
ncu -i gpurun_out/prof_${TAG}_step.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_step_raw.csv 2>/dev/null
# gpurun brings back at most 64 MiB: drop the report itself when it is too large (the raw CSV above stays)
[ $(stat -c %s gpurun_out/prof_${TAG}_step.ncu-rep) -gt 45000000 ] && rm -f gpurun_out/prof_${TAG}_step.ncu-rep
ls -la gpurun_out/
