#!/bin/bash
# round-2 GPU pass R (final tree): whole GPU suite + smoke
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time python -m pytest tests -m gpu -q -rA ) 2>&1 | grep -vE "Warning|^$" | tail -240 > $O/r2r_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2r_smoke.log 2>&1
grep -E "passed|failed" $O/r2r_tests.log | tail -3; tail -2 $O/r2r_smoke.log
