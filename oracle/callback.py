"""ORACLE (test infrastructure only): CPU restatement of the model-evaluation callbacks.

Restates, with the HF forward from ``oracle/w2v2_forward.py`` underneath:

* ``masker``                       feasability_tests/conformer_test.ipynb:138-141 (fill value 0.0)
* ``ModelWrapper.forward``         shap_calculation.py:31-50              (mode "max": max logit / frame)
* ``predict_function``             feasability_tests/w2v2conformer.py:116-131 (mode "logit": one (t, tok) logit)
* ``lime_predict_fn``              feasability_tests/lime_shap_wav2vec2_comparison.py:60-71 (mode "mean")
* per-character frame selection    visualization.py:313-327
* first-character target           feasability_tests/w2v2conformer.py:93-110
* processor normalisation          HF wav2vec2/feature_extraction_wav2vec2.py:78-97

Mask convention (documented deviation): the notebook's
``numpy.ma`` masker zeroes samples where ``mask`` is TRUE; here, as in shap's own
convention, coalition bit 1 = segment KEPT, bit 0 = segment replaced by the baseline.
"""
from __future__ import annotations

import numpy as np
import torch

from . import w2v2_forward as W

MODES = ("max", "logit", "logprob", "mean")
PAD_ID = 0      # CTC blank == <pad>, shap_calculation.py:221-254 / visualization.py:320
SPACE_ID = 4    # "|"


def normalize_clip(x: np.ndarray) -> np.ndarray:
    """(x - mean) / sqrt(var + 1e-7), HF feature_extraction_wav2vec2.py:78-97, once per clean clip."""
    x = np.asarray(x)
    return ((x - x.mean()) / np.sqrt(x.var() + 1e-7)).astype(np.float32)


def segment_bounds(num_samples: int, num_segments: int) -> np.ndarray:
    """Contiguous near-equal blocks: bounds[i] = floor(i * L / M).  The reference has no
    segmentation precedent; when M divides L this is the equal-block rule."""
    i = np.arange(num_segments + 1, dtype=np.int64)
    return ((i * int(num_samples)) // int(num_segments)).astype(np.int32)


def masker(x: np.ndarray, keep: np.ndarray, baseline: float = 0.0) -> np.ndarray:
    """Sample-wise select (conformer_test.ipynb:138-141 with the keep-convention above)."""
    return np.where(np.asarray(keep, dtype=bool), x, np.float32(baseline)).astype(np.float32)


def materialize(x: np.ndarray, Z: np.ndarray, bounds: np.ndarray, baseline: float = 0.0) -> np.ndarray:
    """Coalition matrix Z[K, M] (1 = keep) -> masked waveforms [K, L]."""
    Z = np.asarray(Z)
    lens = np.diff(bounds)
    keep = np.repeat(Z.astype(bool), lens, axis=1)
    return np.where(keep, x[None, :], np.float32(baseline)).astype(np.float32)


def char_targets(logits: np.ndarray, all_frames_if_empty: bool = True):
    """visualization.py:319-327: frames where a new non-blank, non-'|' token starts, with the
    token fixed to the unmasked argmax (w2v2conformer.py:97-108).  Random-init weights can give
    an empty set; then every frame with its argmax token is used."""
    ids = np.asarray(logits).argmax(-1)
    frames = [i for i, t in enumerate(ids)
              if t != PAD_ID and t != SPACE_ID and (i == 0 or t != ids[i - 1])]
    if not frames and all_frames_if_empty:
        frames = list(range(len(ids)))
    frames = np.asarray(frames, dtype=np.int32)
    return frames, ids[frames].astype(np.int32)


def first_char_target(logits: np.ndarray, special_ids=(0, 1, 2, 3)):
    """w2v2conformer.py:93-110: first frame whose argmax is non-special and not '|';
    falls back to the middle frame."""
    ids = np.asarray(logits).argmax(-1)
    for i, t in enumerate(ids):
        if t not in special_ids and t != SPACE_ID:
            return int(i), int(t)
    mid = len(ids) // 2
    return int(mid), int(ids[mid])


def reduce_logits(logits: torch.Tensor, mode: str, frames=None, tokens=None) -> torch.Tensor:
    if mode == "max":          # shap_calculation.py:50
        return logits.max(dim=-1).values
    if mode == "mean":         # lime_shap_wav2vec2_comparison.py:68-70
        return logits.mean(dim=-1).mean(dim=1, keepdim=True)
    f = torch.as_tensor(np.asarray(frames), dtype=torch.long)
    t = torch.as_tensor(np.asarray(tokens), dtype=torch.long)
    if mode == "logit":        # w2v2conformer.py:40-42
        return logits[:, f, t]
    if mode == "logprob":      # north-star reduction: log_softmax(logits)[t_c, id_c]
        return torch.log_softmax(logits, dim=-1)[:, f, t]
    raise ValueError(mode)


@torch.no_grad()
def evaluate(sd, cfg: dict, X: np.ndarray, mode: str = "logprob", frames=None, tokens=None,
             batch: int = 32, dtype=torch.float32) -> np.ndarray:
    """Waveforms X[n, L] -> reduced outputs [n, D] (the numpy-in/numpy-out callback shape)."""
    X = np.atleast_2d(np.asarray(X))
    outs = []
    for s in range(0, X.shape[0], batch):
        xb = torch.from_numpy(np.ascontiguousarray(X[s:s + batch])).to(dtype)
        outs.append(reduce_logits(W.ctc_logits(sd, cfg, xb), mode, frames, tokens).float())
    return torch.cat(outs).numpy()


@torch.no_grad()
def evaluate_coalitions(sd, cfg, x, Z, bounds, baseline=0.0, **kw) -> np.ndarray:
    """f(Z[n, M]) -> [n, D]: what KernelExplainer(f, zeros[1,M]).shap_values(ones[1,M]) calls."""
    Z = np.atleast_2d(np.asarray(Z))
    batch = kw.pop("batch", 32)
    outs = []
    for s in range(0, Z.shape[0], batch):
        outs.append(evaluate(sd, cfg, materialize(x, Z[s:s + batch], bounds, baseline), batch=batch, **kw))
    return np.concatenate(outs)
