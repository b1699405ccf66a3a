#!/bin/bash
# round-2 GPU pass D: bf16 pre-LayerNorm tensors as default (carried-LayerNorm epilogues reverted) -- tests + bench A/B
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
python -m pytest tests -m gpu -q -rA 2>&1 | grep -vE "Warning|warnings|^$" | tail -170 > $O/r2d_tests.log
python bench.py > $O/r2d_bench_default.json 2> $O/r2d_bench_default.err
python bench.py --no-cpu --steps 3 --preln-fp32 > $O/r2d_bench_fp32preln.json 2>> $O/r2d.err
for w in C1 C3 C4; do python bench.py --no-cpu --steps 3 --workload $w > $O/r2d_bench_$w.json 2>> $O/r2d.err; done
tail -3 $O/r2d_tests.log
