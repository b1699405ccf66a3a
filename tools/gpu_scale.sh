#!/bin/bash
# weak-scaling run on N GPUs of one box (torchrun, one rank per GPU)
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.log 2>&1
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${N}gpu.log").read().strip().splitlines()[-1])
    print("N=$N VALUE", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]), "sec/clip", round(d["config"]["sec_per_explained_clip"],4), d["clocks"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_${N}gpu.log").read()[-3000:])
PY
