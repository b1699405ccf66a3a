#!/bin/bash
# ncu --set full of the fused attention backward kernel inside the gradient bench (wav2vec2-base, T' = 573, 32 rows per call)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
CMD="python tools/bench_grad.py --rows 32 --steps 1"
$CMD > $O/r2m_plain.json 2> $O/r2m_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:"attention_bwd_kernel|attn_delta|attn_dq_cast" -s 15 -c 3 -f -o $O/prof_attn_bwd $CMD > $O/r2m_ncu.log 2>&1
tail -2 $O/r2m_ncu.log; ls -la $O/prof_attn_bwd.ncu-rep
