#!/bin/bash
# profile selected forward kernels once each (run under gpurun; writes gpurun_out/*.ncu-rep)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f -k kernel_entry -s 1 -c 1"
for spec in "$@"; do
  set -- $spec
  $NCU -o gpurun_out/prof_$1_$2 python tools/prof_one.py $1 $2 > gpurun_out/ncu_$1_$2.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
