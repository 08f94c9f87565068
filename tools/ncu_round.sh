#!/bin/bash
# Round profile capture (run under gpurun): launch list of one bench step + full captures of
# the forward kernels in the shapes of the step.  Outputs under gpurun_out/.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:kernel_entry --csv \
    --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-roofline --no-selfcheck \
    > gpurun_out/launches_bench.json 2> gpurun_out/launches_bench.err
NCU="ncu --set full --clock-control none --import-source on -f -k regex:kernel_entry"
$NCU -s 2 -c 2 -o gpurun_out/prof_f1_cfg4 python tools/prof_one.py f1 cfg4 > gpurun_out/ncu_f1_cfg4.log 2>&1
for which in f2 f3 f4; do
  $NCU -s 1 -c 1 -o gpurun_out/prof_${which}_cfg4 python tools/prof_one.py $which cfg4 > gpurun_out/ncu_${which}_cfg4.log 2>&1
done
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
