#!/bin/bash
# A/B of environment switches on the C2 step: usage gpu_ab.sh "VAR=a" "VAR=b" ...
for v in "$@"; do
  env $v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ab.log 2>&1
  python - "$v" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/ab.log").read().strip().splitlines()[-1])
kb=d["kernel_breakdown"]
print(sys.argv[1], "ms/step", round(d["ms_per_step"],1), "mhz", d["clocks"]["sm_mhz"], " ".join(f"{k}={v['ms_per_step']:.2f}" for k,v in kb.items() if k in ("pos_conv","attention","conv0","ffn1","qkv","out_proj")))
PY
done
