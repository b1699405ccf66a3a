#!/bin/bash
# round-2 GPU pass O (final tree): whole GPU suite + smoke + default bench + reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time python -m pytest tests -m gpu -q -rA ) 2>&1 | grep -vE "Warning|^$" | tail -220 > $O/r2o_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2o_smoke.log 2>&1
python bench.py > $O/r2o_bench_default.json 2> $O/r2o_bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2o_bench_reference.json 2> $O/r2o_bench_reference.err
grep -E "passed|failed" $O/r2o_tests.log | tail -3; tail -2 $O/r2o_smoke.log; python tools/show_bench.py $O/r2o_bench_default.json | head -3
