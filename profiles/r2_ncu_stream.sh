#!/bin/bash
# ncu --set full with source correlation of one steady-state streamed (T = 1) forward launch of a feed-forward layer
set -u
mkdir -p gpurun_out
CMD="python profiles/run_stream_forward.py --n 2"
$CMD > gpurun_out/r2_stream_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wt_fwd_kernel -s 30 -c 1 -o gpurun_out/prof_r2_stream_fwd $CMD > gpurun_out/r2_ncu_stream.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/r2_ncu_stream.log
