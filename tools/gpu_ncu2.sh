#!/bin/bash
# ncu --set full for the non-contraction kernels at the real batch size (152 coalitions)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --coalitions 152"
$CMD > gpurun_out/ncu_plain.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:"mask_kernel|conv0_kernel|conv0_stats" -s 3 -c 3 -o gpurun_out/prof_front $CMD > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none -k regex:"head_reduce" -c 2 -o gpurun_out/prof_head $CMD > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none -k regex:"attention_fa|posconv_kernel|layernorm_vec" -s 41 -c 6 -o gpurun_out/prof_mid $CMD > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
