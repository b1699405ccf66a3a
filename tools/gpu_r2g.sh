#!/bin/bash
# round-2 GPU pass G (final tree): whole GPU suite + smoke + default bench + reference arm + conformer gradient throughput,
# then ONE ncu --set full capture of the front-end kernels at the bench's batch tile
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time python -m pytest tests -m gpu -q -rA ) 2>&1 | grep -vE "Warning|warnings|^$" | tail -200 > $O/r2g_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2g_smoke.log 2>&1
python bench.py > $O/r2g_bench_default.json 2> $O/r2g_bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2g_bench_reference.json 2> $O/r2g_bench_reference.err
python tools/bench_grad.py --model wav2vec2-conformer-large --samples 80000 --rows 32 > $O/r2g_bench_grad_conformer.json 2> $O/r2g_bench_grad_conformer.err
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --coalitions 152"
# matching launches 0-2 belong to the 1-row target-selection forward; 3-5 are the first 152-coalition tile
$CMD > $O/r2g_ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"conv0_mma|conv0_stats|head_reduce" -s 3 -c 3 -f -o $O/prof_front_g $CMD > $O/r2g_ncu_front.log 2>&1
tail -5 $O/r2g_tests.log; tail -2 $O/r2g_smoke.log; tail -c 600 $O/r2g_bench_grad_conformer.json; ls -la $O/*.ncu-rep
