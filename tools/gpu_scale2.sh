#!/bin/bash
# N-GPU weak-scaling bench line (N = $1), one rank per GPU over NCCL
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_${N}gpu.log 2>&1
tail -n 1 gpurun_out/bench_${N}gpu.log | cut -c 1-400
