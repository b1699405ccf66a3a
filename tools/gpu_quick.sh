#!/bin/bash
# quick regression + bench after a kernel change
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -q -x -s 2>&1 | grep -E "rel err|KernelSHAP|passed|failed|Error|error" | tail -25 )
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_quick.log 2>&1
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_quick.log").read().strip().splitlines()[-1])
    print("VALUE", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]), "gemm TF/s", round(d["roofline"]["achieved"]), "frac", round(d["roofline"]["frac"],3), d["clocks"])
    for k,v in d["kernel_breakdown"].items(): print(f"  {k:12s} {v['ms_per_step']:7.2f} ms  {('%5.0f TF/s'%v['tflops']) if v['tflops'] else ''} {('%5.0f GB/s'%v['gbs']) if v['gbs'] else ''}")
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_quick.log").read()[-3000:])
PY
