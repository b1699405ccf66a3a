#!/bin/bash
# round-2 GPU pass F (final evidence): whole GPU suite, smoke, default bench, C1 (pair-kernel threshold), gradient bench,
# ncu launch list + full captures of the final kernels
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time python -m pytest tests -m gpu -q -rA ) 2>&1 | grep -vE "Warning|warnings|^$" | tail -200 > $O/r2f_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2f_smoke.log 2>&1
python bench.py > $O/r2f_bench_default.json 2> $O/r2f_bench_default.err
for w in C1 C3 C4; do python bench.py --no-cpu --steps 3 --workload $w > $O/r2f_bench_$w.json 2>> $O/r2f.err; done
python tools/bench_grad.py --rows 32 > $O/r2f_gradbench.json 2>> $O/r2f.err
python tools/bench_grad.py --rows 64 --samples 80000 > $O/r2f_gradbench_5s.json 2>> $O/r2f.err
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --coalitions 152"
$CMD > $O/r2f_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 400 --csv --log-file $O/launches.csv $CMD > $O/r2f_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 40 -c 12 -f -o $O/prof_gemm_pair $CMD > $O/r2f_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv0_mma|conv0_stats|head_reduce|layernorm_vec" -s 8 -c 5 -f -o $O/prof_front $CMD > $O/r2f_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"attention_fa|posconv_kernel" -s 26 -c 3 -f -o $O/prof_mid $CMD > $O/r2f_ncu_c.log 2>&1
tail -4 $O/r2f_tests.log; ls -la $O/*.ncu-rep
