#!/bin/bash
# ncu evidence for the bench command: (1) launch list with device times, (2) --set full captures of the dominant kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --coalitions 152"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
# launches 0..~320 are weight re-layout + the 1-row target-selection forward; the 152-coalition batches follow
$CMD > gpurun_out/ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none -k regex:gemm_tc2_kernel -s 40 -c 12 -o gpurun_out/prof_gemm_pair $CMD > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:"conv0_kernel|attention_tc_kernel|layernorm_vec|posconv_kernel|mask_kernel|head_reduce" -s 8 -c 10 -o gpurun_out/prof_others $CMD > gpurun_out/ncu_others.log 2>&1
tail -n 1 gpurun_out/ncu_list.log gpurun_out/ncu_full.log gpurun_out/ncu_others.log

ls -la gpurun_out/ | tail -12; du -sh gpurun_out
