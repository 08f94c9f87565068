#!/bin/bash
# Round-end evidence (run under gpurun, one GPU): bench line, full ncu capture of K2, the GPU
# test suite, launch list of one bench step.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
NCU="ncu --set full --clock-control none --import-source on -f -k regex:kernel_entry"
timeout 150 $NCU -s 1 -c 1 -o gpurun_out/prof_k2_tmem python tools/prof_one.py f2 cfg4 > gpurun_out/ncu_k2_tmem.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_k2_tmem.ncu-rep > gpurun_out/r02_ncu_k2_tmem.txt 2>&1
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest_gpu.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:kernel_entry --csv \
    --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-roofline --no-selfcheck \
    > gpurun_out/launches_bench.json 2> gpurun_out/launches_bench.err
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/r02_launch_list_summary.txt 2>&1
tail -2 gpurun_out/r02c_pytest_gpu.txt
tail -c 300 gpurun_out/r02c_bench_n1.json
