#!/bin/bash
# round-end evidence: full GPU suite, smoke, default bench line, the other workloads, ncu launch list + full captures
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 )
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 )
timeout 900 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.err
for w in C1 C3 C4; do timeout 900 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$w.log 2>&1; done
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --coalitions 152"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none -k regex:gemm_tc2_kernel -s 40 -c 12 -f -o gpurun_out/prof_gemm_pair $CMD > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:"mask_kernel|conv0_mma|conv0_stats" -s 3 -c 3 -f -o gpurun_out/prof_front $CMD > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none -k regex:"attention_fa|posconv_kernel|layernorm_vec" -s 41 -c 6 -f -o gpurun_out/prof_mid $CMD > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
