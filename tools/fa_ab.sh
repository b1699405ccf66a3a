#!/bin/bash
# A/B of two prebuilt attention variants (libw2s_a.so / libw2s_b.so, built in the dev container) + ncu of one launch each
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
P=shap_transformer_asr_b200
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-graph --coalitions 152"
for v in a b; do
  cp $P/libw2s_$v.so $P/libw2s.so
  $CMD > $O/fa_ab_$v.json 2> $O/fa_ab_$v.err
  python tools/show_bench.py $O/fa_ab_$v.json 2>/dev/null | grep -E "VALUE|attention"
done
# ncu: the 30th attention launch of each variant (warm-up tile), full set with source counters
for v in none; do
  cp $P/libw2s_$v.so $P/libw2s.so
  ncu --set full --clock-control none --import-source on -k regex:attention_fa -s 30 -c 1 -f -o $O/prof_fa_$v $CMD > $O/fa_ncu_$v.log 2>&1
done
cp $P/libw2s_b.so $P/libw2s.so
timeout 400 python -m pytest tests/test_gpu_parity.py -q -x -k "tiny_logits or long_clips or beyond_512 or bench_configuration" 2>&1 | tail -3
