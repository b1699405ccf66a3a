#!/bin/bash
# one attention_fa launch of the C2 step under ncu (full set + source), after the same command ran clean without ncu
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --coalitions 152"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_fa_kernel -s 14 -c 1 -f -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
tail -n 2 gpurun_out/ncu_attn.log
ls -la gpurun_out/prof_attn.ncu-rep
