#!/bin/bash
# ncu evidence for the bench command: (1) launch list with device times, (2) one --set full capture of the top kernel.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --coalitions 152"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 20 -c 3 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_list.log gpurun_out/ncu_full.log
