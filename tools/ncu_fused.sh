#!/bin/bash
# ncu --set full capture of the fused tile kernel for the variants given (after a plain run of the same command)
n=${1:-16}; shift
for v in "$@"; do
  CVB_FUSED=$v python tools/prof_run.py $n 1 > gpurun_out/plain_$v.log 2>&1 &&
  CVB_FUSED=$v ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -f -o gpurun_out/fused_$v python tools/prof_run.py $n 1 > gpurun_out/ncu_$v.log 2>&1
  tail -3 gpurun_out/ncu_$v.log
done
