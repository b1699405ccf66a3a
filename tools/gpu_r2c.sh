#!/bin/bash
# round-2 GPU pass C: carried LayerNorm v2 (packed epilogue math, residual prefetch, parallel finalize) + tensor-core layer-norm conv0
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $O/r2c_tests.log
python bench.py --no-cpu --steps 3 > $O/r2c_bench_default.json 2> $O/r2c.err
python bench.py --no-cpu --steps 3 --unfused-ln > $O/r2c_bench_unfused.json 2>> $O/r2c.err
for w in C1 C3 C4; do python bench.py --no-cpu --steps 3 --workload $w > $O/r2c_bench_$w.json 2>> $O/r2c.err; done
cat $O/r2c_tests.log
