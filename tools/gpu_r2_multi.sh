#!/bin/bash
# round-2 multi-GPU pass: usage  gpurun --gpus N -- 'bash tools/gpu_r2_multi.sh N "C2 C3 C4" [c5]'
cd "$GRAFT_REPO_ROOT" || exit 1
N=$1; WL=$2; C5=$3
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for w in $WL; do
  $TR bench.py --gpus $N --steps 3 --warmup 3 --workload $w > $O/r2_multi_${w}_${N}gpu.json 2> $O/r2_multi_${w}_${N}gpu.err
  tail -c 300 $O/r2_multi_${w}_${N}gpu.err
done
if [ -n "$C5" ]; then
  $TR tools/run_c5.py --clips 64 --save-limit 1 --out /tmp/data_c5 > $O/r2_multi_C5_${N}gpu.json 2> $O/r2_multi_C5_${N}gpu.err
  tail -c 300 $O/r2_multi_C5_${N}gpu.err
  du -sh /tmp/data_c5 | tail -1
fi
ls -la $O | grep r2_multi
