#!/bin/bash
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --coalitions 152"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_fa_kernel -s 14 -c 1 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
tail -n 2 gpurun_out/ncu_attn.log
