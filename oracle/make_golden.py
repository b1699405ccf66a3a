"""Generates tests/golden/*.npz from the REAL third-party reference stack (run in the build container).

    python -m oracle.make_golden

The model forward of the reference path lives in ``transformers`` (requirements.txt:3); this script
imports the installed copy (5.5.0), builds random-init models of the named architectures with fixed
seeds, runs them in eval mode on seeded synthetic waveforms and stores inputs' seeds + outputs as small
fixtures.  Tests then check (a) oracle/w2v2_forward.py and (b) the CUDA path against these vectors
without needing /root/reference or regenerating anything.  Weight tensors are NOT stored (they are
re-created from the seed); a checksum guards against RNG drift.
"""
from __future__ import annotations

import dataclasses
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import callback as CB  # noqa: E402
from oracle import w2v2_forward as W  # noqa: E402
from oracle.kernelshap_ref import KernelExplainerRef  # noqa: E402
from shap_transformer_asr_b200.config import MODELS  # noqa: E402
from shap_transformer_asr_b200.preprocess import synthetic_clip  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TINY = MODELS["wav2vec2-tiny"]
VARIANTS = {
    "tiny_group": TINY,
    "tiny_layer_stable": dataclasses.replace(TINY, feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True),
    "tiny_conformer_rel": dataclasses.replace(TINY, kind="conformer", feat_extract_norm="layer", conv_bias=True,
                                              hidden_act="swish", position_embeddings_type="relative"),
    "tiny_conformer_rotary": dataclasses.replace(TINY, kind="conformer", feat_extract_norm="layer", conv_bias=True,
                                                 hidden_act="swish", position_embeddings_type="rotary"),
}


def weight_checksum(model) -> float:
    return float(sum(p.detach().double().abs().sum() for p in model.state_dict().values() if p.is_floating_point()))


def build(cfg, seed=0):
    return W.randomize_affine(W.build_hf_model(cfg.to_dict(), seed=seed), seed=seed + 1)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    # ---- tiny variants: full logits on seeded noise -------------------------------------------------
    for name, cfg in VARIANTS.items():
        model = build(cfg)
        g = torch.Generator().manual_seed(7)
        x = torch.randn(3, 4000, generator=g)
        with torch.no_grad():
            logits = model(x).logits.numpy()
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), x=x.numpy(), logits=logits,
                            checksum=weight_checksum(model))
        print(name, logits.shape, float(np.abs(logits).mean()))
    # ---- C1: wav2vec2-base, 1 s synthetic clip, 32 segments -----------------------------------------------
    cfg = MODELS["wav2vec2-base"]
    model = build(cfg)
    clip = synthetic_clip(16000)
    bounds = CB.segment_bounds(16000, 32)
    np.random.seed(0)
    Z, kw = KernelExplainerRef(None, 32).sample(256)
    rows = np.concatenate([np.ones((1, 32)), np.zeros((1, 32)), Z[:6], Z[100:106]])
    X = CB.materialize(clip, rows, bounds)
    with torch.no_grad():
        logits = model(torch.from_numpy(X)).logits.numpy()
    frames, tokens = CB.char_targets(logits[0])
    np.savez_compressed(os.path.join(OUT, "c1_base.npz"), rows=rows.astype(np.uint8), logits=logits,
                        frames=frames, tokens=tokens, checksum=weight_checksum(model),
                        clip_head=clip[:64], clip_sum=float(np.abs(clip.astype(np.float64)).sum()))
    print("c1_base", logits.shape, "D =", len(frames))
    # ---- sampler regression pins (self-generated: shap itself is not installable here) ----------------------
    pins = {}
    for M, K in [(32, 256), (100, 2048), (200, 8192), (128, 2048), (12, 300), (8, "auto")]:
        np.random.seed(0)
        Zs, ws = KernelExplainerRef(None, M).sample(K)
        pins[f"M{M}_K{K}"] = np.array([Zs.shape[0], int(hashlib.sha256(Zs.astype(np.uint8).tobytes()).hexdigest()[:12], 16),
                                       float(ws.sum()), float(ws[:2 * M].sum() if Zs.shape[0] > 2 * M else ws.sum())])
    np.savez_compressed(os.path.join(OUT, "sampler_pins.npz"), **pins)
    print("sampler pins", {k: v.tolist() for k, v in pins.items()})


if __name__ == "__main__":
    main()
