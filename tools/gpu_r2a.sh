#!/bin/bash
# round-2 GPU pass A: full GPU test-suite, default bench, A/B switches, other workloads
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
python -m pytest tests -m gpu -q -rA 2>&1 | grep -vE "Warning|warnings|^$" | tail -150 > $O/r2_tests.log
python bench.py > $O/r2_bench_default.json 2> $O/r2_bench_default.err
python bench.py --no-cpu --steps 3 --no-graph > $O/r2_bench_nograph.json 2>> $O/r2_ab.err
python bench.py --no-cpu --steps 3 --preln-bf16 > $O/r2_bench_preln.json 2>> $O/r2_ab.err
python bench.py --no-cpu --steps 3 --pdl > $O/r2_bench_pdl.json 2>> $O/r2_ab.err
for w in C1 C3 C4; do python bench.py --no-cpu --steps 3 --workload $w > $O/r2_bench_$w.json 2>> $O/r2_ab.err; done
( cd shap_transformer_asr_b200/csrc && make clean >/dev/null && make -j16 EXTRA=-DW2S_MBAR_POLLCOUNT >/dev/null 2>&1 )
python bench.py --no-cpu --steps 3 > $O/r2_bench_pollcount.json 2>> $O/r2_ab.err
tail -5 $O/r2_tests.log
