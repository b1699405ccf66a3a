"""Shared test helpers: model variants and HF/oracle construction (tests may use oracle/)."""
import dataclasses
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import w2v2_forward as W  # noqa: E402
from shap_transformer_asr_b200.config import MODELS  # noqa: E402

TINY = MODELS["wav2vec2-tiny"]
VARIANTS = {
    "tiny_group": TINY,
    "tiny_layer_stable": dataclasses.replace(TINY, feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True),
    "tiny_conformer_rel": dataclasses.replace(TINY, kind="conformer", feat_extract_norm="layer", conv_bias=True,
                                              hidden_act="swish", position_embeddings_type="relative"),
    "tiny_conformer_rotary": dataclasses.replace(TINY, kind="conformer", feat_extract_norm="layer", conv_bias=True,
                                                 hidden_act="swish", position_embeddings_type="rotary"),
}


def build_model(cfg, seed=0):
    """Same construction as oracle/make_golden.py: seeded random init + perturbed affine terms."""
    return W.randomize_affine(W.build_hf_model(cfg.to_dict(), seed=seed), seed=seed + 1)


def weight_checksum(model) -> float:
    return float(sum(p.detach().double().abs().sum() for p in model.state_dict().values() if p.is_floating_point()))


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
