"""Clip pre-processing and segmentation (host side, once per clip).

* ``normalize_clip`` -- the processor's zero-mean / unit-variance step the reference applies once to
  the clean clip before any explanation (shap_calculation.py:117; HF
  wav2vec2/feature_extraction_wav2vec2.py:78-97).  Masking is applied AFTER it, so the 0.0 baseline
  is the clip mean.
* ``segment_bounds`` -- contiguous near-equal blocks, bounds[i] = floor(i L / M).  The reference has
  no segmentation precedent; when M divides L this is the equal-block rule of the BASELINE configs.
* ``synthetic_clip`` -- the seeded synthetic 16 kHz audio used by tests and bench.
"""
from __future__ import annotations

import numpy as np


def normalize_clip(x) -> np.ndarray:
    x = np.asarray(x)
    return ((x - x.mean()) / np.sqrt(x.var() + 1e-7)).astype(np.float32)


def segment_bounds(num_samples: int, num_segments: int) -> np.ndarray:
    i = np.arange(num_segments + 1, dtype=np.int64)
    return ((i * int(num_samples)) // int(num_segments)).astype(np.int32)


def synthetic_clip(num_samples: int, seed: int = 1234, sr: int = 16000) -> np.ndarray:
    """A few drifting sinusoids plus noise, then processor-normalised."""
    rng = np.random.default_rng(seed)
    t = np.arange(num_samples, dtype=np.float64) / sr
    x = 0.3 * rng.standard_normal(num_samples)
    for f0 in (140.0, 410.0, 1270.0, 2900.0):
        amp = 0.5 + 0.5 * np.sin(2 * np.pi * rng.uniform(0.5, 3.0) * t + rng.uniform(0, 6.28))
        x += amp * np.sin(2 * np.pi * f0 * t + rng.uniform(0, 6.28))
    return normalize_clip(x)


def pack_coalitions(Z) -> np.ndarray:
    """Z[K, M] in {0,1} -> uint32 words [K, ceil(M/32)], bit m of row k = Z[k, m] (1 = keep)."""
    Z = np.atleast_2d(np.asarray(Z))
    K, M = Z.shape
    words = (M + 31) // 32
    pad = np.zeros((K, words * 32), dtype=np.uint8)
    pad[:, :M] = Z != 0
    b = np.packbits(pad.reshape(K, words, 4, 8), axis=-1, bitorder="little").reshape(K, words, 4)
    return (b[..., 0].astype(np.uint32) | (b[..., 1].astype(np.uint32) << 8) |
            (b[..., 2].astype(np.uint32) << 16) | (b[..., 3].astype(np.uint32) << 24))
