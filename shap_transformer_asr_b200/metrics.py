"""Downstream consumers of the attributions: eta_raw, ITM, WER, greedy CTC decode.

Host-side numpy: these are cheap segmented reductions over arrays the path has already produced.
"""
from __future__ import annotations

import numpy as np

# the reference's vocabulary (shap_calculation.py:221-254): <pad>=0 is the CTC blank, "|"=4 the word separator
VOCAB = ["<pad>", "<s>", "</s>", "<unk>", "|", "E", "T", "A", "O", "N", "I", "H", "S", "R", "D", "L", "U", "M", "W", "C", "F",
         "G", "Y", "P", "B", "V", "K", "'", "X", "J", "Q", "Z"]


def eta_raw(clean_audio, noise_audio, shap_matrix, sr: int, segment_ms: float = 20, percentile: float = 99.0,
            itm_ratio: float = 0.5) -> float:
    """Raw-audio speech relevance score.

    Follows ``calculate_eta_raw`` of calculate_metric.py:74-149 (``itm_ratio=0.5``: ITM = E_c > 0.5 E_u, :118);
    ``itm_ratio=1.0`` gives the variant of nraw_vs_wer.py:20-62 (ITM = E_c > E_u, :46).  ``shap_matrix`` is
    ``[L, T']`` (``[T', L]`` is transposed as at calculate_metric.py:92-95).
    """
    clean_audio = np.asarray(clean_audio)
    noise_audio = np.asarray(noise_audio)
    shap_matrix = np.asarray(shap_matrix)
    seg = int(sr * (segment_ms / 1000.0))
    if seg == 0:
        raise ValueError("segment_ms is too small, resulting in 0 samples per segment.")
    if shap_matrix.shape[0] != clean_audio.shape[0]:
        if shap_matrix.shape[1] == clean_audio.shape[0]:
            shap_matrix = shap_matrix.T
        else:
            raise ValueError(f"SHAP matrix shape {shap_matrix.shape} is incompatible with audio length {len(clean_audio)}.")
    min_len = min(len(clean_audio), len(noise_audio), shap_matrix.shape[0])
    nseg = min_len // seg
    trunc = nseg * seg
    if nseg == 0:
        return 0.0
    e_c = np.square(clean_audio[:trunc].reshape(nseg, seg)).sum(axis=1)
    e_u = np.square(noise_audio[:trunc].reshape(nseg, seg)).sum(axis=1)
    itm = (e_c > itm_ratio * e_u).astype(int)
    phi_total = np.abs(shap_matrix[:trunc, :]).sum(axis=1)
    bar_phi = phi_total.reshape(nseg, seg).mean(axis=1)
    tau = np.percentile(bar_phi, percentile)
    relevant = (bar_phi > tau).astype(int)
    den = relevant.sum()
    if den == 0:
        return 0.0
    return float((relevant * itm).sum() / den)


def eta_raw_segments(clean_audio, noise_audio, phi, bounds, sr: int, **kw) -> float:
    """``eta_raw`` for attributions that are constant inside a waveform segment (what KernelSHAP over segment coalitions
    produces): ``phi`` is ``[M, D]``, ``bounds`` the ``M + 1`` segment boundaries.  The per-sample total ``sum_d |phi|`` is
    formed per SEGMENT and repeated, instead of reducing the expanded ``[L, T']`` array (130 MB for a 6.4 s clip) -- the
    same numbers summed in the same order, so the result is bit-identical to ``eta_raw`` on the expanded matrix."""
    phi = np.asarray(phi)
    total = np.repeat(np.abs(phi).sum(axis=1), np.diff(np.asarray(bounds)))
    return eta_raw(clean_audio, noise_audio, total[:, None], sr, **kw)


def greedy_ctc_decode(ids, vocab=VOCAB, pad_id: int = 0, space_id: int = 4) -> str:
    """Collapse repeats, drop blanks, map '|' to a space (what processor.batch_decode does for this vocabulary;
    visualization.py:305-309, nraw_vs_wer.py:75-79)."""
    out, prev = [], None
    for t in np.asarray(ids).tolist():
        if t != prev and t != pad_id:
            out.append(" " if t == space_id else vocab[t])
        prev = t
    return " ".join("".join(out).split())


def wer(reference: str, hypothesis: str) -> float:
    """Word error rate (S + D + I) / N over whitespace-split words -- the definition jiwer.wer implements
    (nraw_vs_wer.py:82; jiwer is not installable here)."""
    ref, hyp = reference.split(), hypothesis.split()
    if not ref:
        return 0.0 if not hyp else float(len(hyp))
    d = np.arange(len(hyp) + 1)
    for i, r in enumerate(ref, 1):
        prev, d[0] = d[0], i
        for j, h in enumerate(hyp, 1):
            cur = d[j]
            d[j] = min(d[j] + 1, d[j - 1] + 1, prev + (r != h))
            prev = cur
    return float(d[len(hyp)] / len(ref))
