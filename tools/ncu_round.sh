#!/bin/bash
# Round profile capture (run under gpurun): launch list of one bench step + full captures of
# the four forward kernels.  Outputs under gpurun_out/.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-roofline > gpurun_out/launches_bench.json 2> gpurun_out/launches_bench.err
NCU="ncu --set full --clock-control none --import-source on -f -k kernel_entry -s 1 -c 1"
for spec in "f1 cfg4" "f2 cfg4" "f3 cfg4" "f4 cfg4"; do
  set -- $spec
  $NCU -o gpurun_out/prof_$1_$2 python tools/prof_one.py $1 $2 > gpurun_out/ncu_$1_$2.log 2>&1
done
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
