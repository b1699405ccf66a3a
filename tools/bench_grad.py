#!/usr/bin/env python
"""Expected-gradients path: forward + backward passes per second (the unit the reference's recorded run is quoted in:
114 600 passes in 5586 s for the 11.5 s clip, evaluation.ipynb:463,513 -> 20.5 passes/s at batch 1 on their GPU).

    python tools/bench_grad.py [--model wav2vec2-base] [--samples 183600] [--rows 64] [--steps 3]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_grad.py --explain-frames -1
        (the output frames of the explained clip are split over the ranks; one all-gather at the end)
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shap_transformer_asr_b200 import MODELS, Engine, synthetic_clip
from shap_transformer_asr_b200.modelzoo import build_random_init_model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="wav2vec2-base")
    ap.add_argument("--samples", type=int, default=183600)
    ap.add_argument("--rows", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--simt-attention", action="store_true", help="attention backward on the CUDA-core cross-check kernels")
    ap.add_argument("--unfused-attention", action="store_true", help="attention backward as batched contractions + row kernels "
                    "(the form before the fused tcgen05 kernel; still the product path of the relative-position conformer)")
    ap.add_argument("--explain-frames", type=int, default=0, help="also time ExpectedGradientsExplainer.shap_values over the "
                    "first N output frames (-1 = all T' frames: the reference's full job) with 200 samples and 5 backgrounds")
    args = ap.parse_args()
    cfg = MODELS[args.model]
    from shap_transformer_asr_b200 import dist as wdist
    rank, world = wdist.init_from_env()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    eng = Engine(build_random_init_model(cfg, seed=0), cfg, device=local, max_batch=4)
    eng.grad_debug(False, simt_attention=args.simt_attention, unfused_attention=args.unfused_attention)
    L = args.samples
    T = eng.num_frames(L)
    x = torch.from_numpy(np.stack([synthetic_clip(L, seed=s) for s in range(4)])).cuda()
    x = x[torch.arange(args.rows) % 4].contiguous()
    frames = (np.arange(args.rows) * 7) % T
    eng.grad_waveforms(x, frames)                                  # plans, transposed weights
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        g, out = eng.grad_waveforms(x, frames)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    eng.profile(True)
    eng.grad_waveforms(x, frames)
    torch.cuda.synchronize()
    prof = eng.profile_read()
    eng.profile(False)
    tot = sum(v["ms"] for v in prof.values())
    explain = None
    if args.explain_frames:
        import time
        from shap_transformer_asr_b200 import ExpectedGradientsExplainer, make_background
        nf = T if args.explain_frames < 0 else min(T, args.explain_frames)
        ex = ExpectedGradientsExplainer(eng, make_background(L, 5, seed=0), nsamples=200, seed=0, batch=args.rows)
        clip = synthetic_clip(L)
        torch.cuda.synchronize()
        wdist.barrier()
        t0 = time.perf_counter()
        phi = ex.shap_values(clip, np.arange(nf, dtype=np.int32))
        torch.cuda.synchronize()
        wdist.barrier()
        dt = time.perf_counter() - t0
        explain = {"frames": nf, "samples_per_frame": 200, "passes": nf * 200, "seconds": dt, "passes_per_s": nf * 200 / dt,
                   "shap_shape": list(phi.shape), "finite": bool(np.isfinite(phi).all()),
                   "reference_recorded_seconds_all_frames": 5586.0 if L == 183600 else None, "n_gpus": world,
                   "sharding": "output frames split over the ranks, one all-gather of phi [D, L]"}
    top = {k: round(v["ms"], 2) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:14]}
    fwd_names = {"conv0_stats", "conv0", "featproj_ln", "featproj", "pos_pad", "pos_conv", "pos_add", "encoder_ln", "qkv",
                 "attention", "out_proj", "ln1", "ffn1", "ffn_gelu", "ffn2", "ln2", "lm_head"}
    fwd_names |= {f"conv{l}" for l in range(1, 8)} | {f"conv{l}_gelu" for l in range(0, 8)}
    fwd = sum(v["ms"] for k, v in prof.items() if k.replace("grad.", "") in fwd_names)
    if rank != 0:
        return
    print(json.dumps({"metric": "expected_gradient_passes_per_sec", "value": args.rows / (ms / 1e3), "unit": "fwd+bwd passes/s",
                      "model": args.model, "attention_backward": "cuda_core" if args.simt_attention else ("contractions" if args.unfused_attention or (cfg.kind == "conformer" and cfg.position_embeddings_type == "relative") else "fused_tcgen05"), "num_samples": L, "frames": T, "rows_per_call": args.rows, "ms_per_call": ms,
                      "reference_recorded": {"passes_per_s": 114600 / 5586.0, "source": "evaluation.ipynb:463,513 (batch 1, GPU model not recorded)"},
                      "explain": explain, "profiled_ms": tot, "forward_share": fwd / tot, "top_steps_ms": top,
                      "clip_estimate_s": {"passes": T * 200, "seconds": T * 200 / (args.rows / (ms / 1e3))}}), flush=True)


if __name__ == "__main__":
    main()
