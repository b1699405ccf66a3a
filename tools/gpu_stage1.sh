#!/bin/bash
# First on-hardware pass: each group in its own process so one faulting kernel cannot poison the rest.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; ( timeout "$@" ) > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -5 gpurun_out/$name.log; }
run t1_simt   300 python -m pytest tests/test_gpu_kernels.py -q -s -k "simt or mask or wls"
run t2_tc     300 python -m pytest tests/test_gpu_kernels.py -q -s -k "tcgen05"
run t3_valid  600 python -m pytest tests/test_gpu_parity.py -q -s -k "validate"
run t4_fast   900 python -m pytest tests/test_gpu_parity.py -q -s -k "not validate"
run t5_smoke  300 python -c "import __graft_entry__ as g; g.smoke()"
run t6_bench  900 python bench.py --steps 2 --warmup 3
