#!/bin/bash
# 2-GPU weak-scaling bench + other workloads on one GPU
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1
tail -c 1500 gpurun_out/bench_2gpu.log | cut -c 1-1500
for wl in C1 C3 C4; do
  timeout 900 python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_$wl.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$wl.log").read().strip().splitlines()[-1])
    print("$wl VALUE", round(d["value"],1), "ms/step", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],1), "gemm TF/s", round(d["roofline"]["achieved"]), "whole", round(d["roofline"]["whole_step_tflops"]), d["config"]["batch_tile"])
    for k,v in list(d["kernel_breakdown"].items())[:8]: print(f"  {k:12s} {v['ms_per_step']:8.2f} ms  {('%5.0f TF/s'%v['tflops']) if v['tflops'] else ''}")
except Exception as e:
    print("$wl failed", e); print(open("gpurun_out/bench_$wl.log").read()[-2000:])
PY
done
