#!/bin/bash
# Round profile capture (run under gpurun): launch list of one bench step + full captures of
# the forward kernels.  Outputs under gpurun_out/.  Usage: tools/ncu_round.sh [kernels...]
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-roofline > gpurun_out/launches_bench.json 2> gpurun_out/launches_bench.err
NCU="ncu --set full --clock-control none --import-source on -f -k kernel_entry -c 2"
for spec in "${@:-f1 f2 f3 f4}"; do
  set -- $spec
  $NCU -s ${2:-1} -o gpurun_out/prof_$1_cfg4 python tools/prof_one.py $1 cfg4 > gpurun_out/ncu_$1_cfg4.log 2>&1
done
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
