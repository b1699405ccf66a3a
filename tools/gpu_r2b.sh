#!/bin/bash
# round-2 GPU pass B: carried LayerNorm -- tests, A/B, other workloads, ncu launch list + full captures
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
python -m pytest tests -m gpu -q -rA 2>&1 | grep -vE "Warning|warnings|^$" | tail -170 > $O/r2b_tests.log
python bench.py > $O/r2b_bench_default.json 2> $O/r2b_bench_default.err
python bench.py --no-cpu --steps 3 --unfused-ln > $O/r2b_bench_unfused.json 2>> $O/r2b_ab.err
for w in C1 C3; do python bench.py --no-cpu --steps 3 --workload $w > $O/r2b_bench_$w.json 2>> $O/r2b_ab.err; done
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --coalitions 152"
$CMD > $O/r2b_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 400 --csv --log-file $O/launches.csv $CMD > $O/r2b_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 40 -c 12 -f -o $O/prof_gemm_pair $CMD > $O/r2b_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv0_mma|conv0_stats|ln_stats_finalize|head_reduce" -s 4 -c 4 -f -o $O/prof_front $CMD > $O/r2b_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"attention_fa|posconv_kernel|layernorm_vec" -s 14 -c 4 -f -o $O/prof_mid $CMD > $O/r2b_ncu_c.log 2>&1
CMD3="python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --workload C3 --coalitions 37"
$CMD3 > $O/r2b_ncu_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"attention_fa" -s 26 -c 2 -f -o $O/prof_attn_c3 $CMD3 > $O/r2b_ncu_d.log 2>&1
tail -4 $O/r2b_tests.log; ls -la $O/*.ncu-rep
