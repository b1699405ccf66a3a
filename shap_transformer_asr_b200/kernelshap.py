"""KernelSHAP around the device callback: host coalition sampler, sharded evaluation, device solve.

The sampler restates ``shap.KernelExplainer.explain`` (shap is pinned only as ``>=0.40.0`` in the
reference's requirements.txt:4 and is not installed here; algorithm per SURVEY.md Appendix A).  It
draws from the GLOBAL legacy ``np.random`` state in shap's exact call order --
``np.random.choice(len(p), 4 * left, p=p)`` once, then one ``np.random.permutation(M)`` per draw --
so ``np.random.seed(s)`` fixes the coalition index sets exactly as it would for shap.  ``l1_reg`` is
pinned to ``False`` (plain constrained WLS, solved on the device by ``w2s_wls``).
"""
from __future__ import annotations

import itertools
from math import comb
from typing import Optional

import numpy as np
import torch

from . import dist as wdist
from .preprocess import pack_coalitions
from .targets import char_targets


def sample_coalitions(M: int, nsamples="auto", seed: Optional[int] = None):
    """-> (Z uint8 [K, M] with 1 = segment kept, kernel weights float64 [K], info dict)."""
    if seed is not None:
        np.random.seed(seed)
    M = int(M)
    if M < 2:
        raise ValueError("KernelSHAP needs at least 2 features")
    nsamples = 2 * M + 2 ** 11 if nsamples == "auto" else int(nsamples)
    max_samples = 2 ** 30
    if M <= 30:
        max_samples = 2 ** M - 2
        nsamples = min(nsamples, max_samples)
    Z = np.zeros((nsamples, M), dtype=np.uint8)
    kw = np.zeros(nsamples, dtype=np.float64)
    added = 0

    n_sizes = int(np.ceil((M - 1) / 2.0))
    n_paired = int(np.floor((M - 1) / 2.0))
    wv = np.array([(M - 1.0) / (i * (M - i)) for i in range(1, n_sizes + 1)])
    wv[:n_paired] *= 2
    wv /= np.sum(wv)

    # sizes that can be enumerated completely with the sample budget
    n_full = 0
    left = nsamples
    rem = wv.copy()
    for size in range(1, n_sizes + 1):
        nsub = float(comb(M, size))
        if size <= n_paired:
            nsub *= 2
        if left * rem[size - 1] / nsub >= 1.0 - 1e-8:
            n_full += 1
            left -= nsub
            if rem[size - 1] < 1.0:
                rem /= (1 - rem[size - 1])
            w = wv[size - 1] / comb(M, size)
            if size <= n_paired:
                w /= 2.0
            for inds in itertools.combinations(range(M), size):
                Z[added, list(inds)] = 1
                kw[added] = w
                added += 1
                if size <= n_paired:
                    Z[added] = 1 - Z[added - 1]
                    kw[added] = w
                    added += 1
        else:
            break

    n_fixed = added
    samples_left = nsamples - added
    if n_full != n_sizes:
        rem = wv.copy()
        rem[:n_paired] /= 2
        rem = rem[n_full:]
        rem /= np.sum(rem)
        ind_set = np.random.choice(len(rem), 4 * samples_left, p=rem)
        pos = 0
        seen = {}
        # rows are collected as (index in Z, permutation prefix, complement flag) and written in one vectorised pass at
        # the end; the RNG call sequence (one permutation per draw) is exactly shap's
        row = np.zeros(M, dtype=np.uint8)
        new_at, new_rows, comp_at = [], [], []
        while samples_left > 0 and pos < len(ind_set):
            size = int(ind_set[pos]) + n_full + 1
            pos += 1
            row[:] = 0
            row[np.random.permutation(M)[:size]] = 1
            key = row.tobytes()
            at = seen.get(key)
            fresh = at is None
            if fresh:
                seen[key] = added
                samples_left -= 1
                new_at.append(added)
                new_rows.append(key)
                kw[added] = 1.0
                added += 1
            else:
                kw[at] += 1.0
            if samples_left > 0 and size <= n_paired:
                if fresh:
                    samples_left -= 1
                    comp_at.append((added, len(new_rows) - 1))
                    kw[added] = 1.0
                    added += 1
                else:
                    kw[at + 1] += 1.0
        if new_rows:
            R = np.frombuffer(b"".join(new_rows), dtype=np.uint8).reshape(len(new_rows), M)
            Z[np.asarray(new_at)] = R
            if comp_at:
                ca = np.asarray(comp_at)
                Z[ca[:, 0]] = 1 - R[ca[:, 1]]
        weight_left = np.sum(wv[n_full:])
        kw[n_fixed:] *= weight_left / kw[n_fixed:].sum()
    info = dict(nsamples=nsamples, n_fixed=n_fixed, n_full_sizes=n_full, max_samples=max_samples)
    return Z[:added], kw[:added], info


class KernelShapExplainer:
    """Explains one clip: Shapley value of every waveform segment for every per-character output.

    ``engine`` is a :class:`~shap_transformer_asr_b200.engine.Engine`.  With ``torch.distributed``
    initialised, the coalition rows are sharded across ranks (no data-path collective), the
    per-coalition outputs are all-gathered once, and every rank solves the regression."""

    def __init__(self, engine, nsamples="auto", seed: int = 0):
        self.engine = engine
        self.nsamples = nsamples
        self.seed = seed

    def select_targets(self, mode: str = "logprob"):
        eng = self.engine
        eng.set_targets("logits")
        ones = eng.bits_to_device(np.ones((1, eng.num_segments), dtype=np.uint8))
        logits = eng.eval_bits(ones).view(-1, eng.config.vocab_size).cpu().numpy()
        frames, tokens = char_targets(logits)
        eng.set_targets(mode, frames, tokens)
        return frames, tokens, logits

    def explain(self, clip, num_segments: int, mode: str = "logprob", targets=None, baseline: float = 0.0):
        eng = self.engine
        eng.set_clip(clip, num_segments=num_segments, baseline=baseline)
        if mode in ("max", "mean", "logits"):
            frames, tokens = (), ()
            eng.set_targets(mode)
        M = eng.num_segments
        Z = None
        if mode in ("max", "mean", "logits"):
            pass
        elif targets is None:
            # the one-row target-selection forward is asynchronous on the engine's stream: draw the coalitions on the
            # host while it runs, read the logits back afterwards
            eng.set_targets("logits")
            ones = eng.bits_to_device(np.ones((1, M), dtype=np.uint8))
            logits_dev = eng.eval_bits(ones)
            Z, kw, info = sample_coalitions(M, self.nsamples, seed=self.seed)
            logits = logits_dev.view(-1, eng.config.vocab_size).cpu().numpy()
            frames, tokens = char_targets(logits)
            eng.set_targets(mode, frames, tokens)
        else:
            frames, tokens = targets
            eng.set_targets(mode, frames, tokens)
        if Z is None:
            Z, kw, info = sample_coalitions(M, self.nsamples, seed=self.seed)
        K = Z.shape[0]
        # rows 0/1 of the evaluated matrix are the empty and the full coalition (fnull, fx)
        Zall = np.concatenate([np.zeros((1, M), np.uint8), np.ones((1, M), np.uint8), Z])
        rank, world = wdist.rank_world()
        lo, hi = wdist.shard_range(K + 2, rank, world)
        bits_all = eng.bits_to_device(Zall)
        y_local = eng.eval_bits(bits_all[lo:hi]) if hi > lo else torch.empty((0, eng.out_width()), device=eng.device)
        y_all = wdist.all_gather_rows(y_local, K + 2, rank, world)
        fnull = y_all[0].double()
        fx = y_all[1].double()
        w_dev = torch.from_numpy(kw).to(eng.device)
        phi, status = eng.wls(bits_all[2:], w_dev, y_all[2:].contiguous(), fx, fnull, M)
        return dict(phi=phi, fx=fx, fnull=fnull, frames=frames, tokens=tokens, Z=Z, weights=kw, y=y_all[2:],
                    status=status, info=info)


def expand_to_samples(phi: np.ndarray, bounds: np.ndarray, per_sample: bool = False) -> np.ndarray:
    """phi[M, D] -> [1, L, D], the reference's on-disk layout (shap_calculation.py:200-210;
    evaluation.ipynb:503-504).  Each segment's value is repeated over its samples (or divided by the
    segment length when ``per_sample``); visualization.py:354 only uses relative magnitudes."""
    lens = np.diff(bounds)
    v = phi / lens[:, None] if per_sample else phi
    return np.repeat(v, lens, axis=0)[None]
