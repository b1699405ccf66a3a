#!/bin/bash
# round-2 GPU pass E: whole GPU suite (incl. the gradient path) + smoke + default bench + reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time python -m pytest tests -m gpu -q -rA ) 2>&1 | grep -vE "Warning|warnings|^$" | tail -190 > $O/r2e_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2e_smoke.log 2>&1
python bench.py > $O/r2e_bench_default.json 2> $O/r2e_bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2e_bench_reference.json 2> $O/r2e_bench_reference.err
tail -6 $O/r2e_tests.log; tail -2 $O/r2e_smoke.log
