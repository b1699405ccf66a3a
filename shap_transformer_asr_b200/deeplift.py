"""DeepLIFT attributions with the reference's custom handler rules, on the device gradient path.

``feasability_tests/custom_shap_handlers.py:35-80`` registers handlers for shap's ``DeepExplainer`` (PyTorch backend) so
that it can walk the wav2vec2 / conformer graph (``feasability_tests/test_shap_asr.py:67``, ``w2v2conformer.py:139``):

    LayerNorm, GroupNorm -> ``linear_1d``     (gradient passes through unchanged: the ordinary backward)
    SiLU                 -> ``nonlinear_1d``  (the DeepLIFT rescale rule: multiplier (y - y_ref) / (x - x_ref))
    GLU                  -> a placeholder: ``grad_output * 5e-6`` wherever the input differs from the reference

and every module shap has no handler for (``GELUActivation``, the attention matmuls, ``F.softmax``) is differentiated
normally.  shap evaluates the model on the stacked batch ``[x; reference]``, runs ONE backward pass with those rules and
returns ``phi = mean over background of  grad * (x - reference)``.  ``DeepLiftExplainer`` restates that loop (algorithm from
memory of shap's ``PyTorchDeep``: parity with shap itself is UNPINNED -- shap is not installable here, and the reference's
own attempts crashed (``conformer_test.ipynb:88-117``); what is pinned is the rule-modified gradient, against a torch
implementation of the same rules with backward hooks on the ``transformers`` model, tests/test_gpu_grad.py).

For ``Wav2Vec2ForCTC`` (GELU everywhere) the rules change nothing: the attribution is gradient x (input - reference).
"""
from __future__ import annotations

import numpy as np
import torch


class DeepLiftExplainer:
    """phi[1, L, D] for the ModelWrapper outputs (max logit per frame) of one clip against a background set."""

    def __init__(self, engine, background: np.ndarray, rescale_silu: bool = True, glu_placeholder: bool = False):
        self.engine = engine
        self.background = np.ascontiguousarray(background, dtype=np.float32)
        self.rescale_silu = bool(rescale_silu)
        self.glu_placeholder = bool(glu_placeholder)

    def shap_values(self, x, frames=None) -> np.ndarray:
        eng = self.engine
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(-1))
        L = x.size
        T = eng.num_frames(L)
        frames = np.arange(T, dtype=np.int32) if frames is None else np.asarray(frames, dtype=np.int32)
        D, B = len(frames), self.background.shape[0]
        half = eng.GRAD_TILE_ROWS // 2                      # pairs per device call
        if B > half:
            raise ValueError(f"at most {half} background rows per call are supported")
        bg = torch.from_numpy(self.background).to(eng.device)
        xt = torch.from_numpy(x).to(eng.device)
        delta = xt[None] - bg                                # [B, L]
        phi = torch.zeros((D, L), dtype=torch.float32, device=eng.device)
        per_call = max(1, half // B)                         # output frames per call
        eng.grad_rules(self.rescale_silu, self.glu_placeholder)
        try:
            for d0 in range(0, D, per_call):
                fr = frames[d0:d0 + per_call]
                k = len(fr)
                # rows: [x for every (frame, background) pair | the backgrounds in the same order]
                rows = torch.cat([xt[None].expand(k * B, L), bg.repeat(k, 1)], 0).contiguous()
                row_frames = np.concatenate([np.repeat(fr, B), np.repeat(fr, B)]).astype(np.int32)
                g, _ = eng.grad_waveforms(rows, row_frames)
                contrib = g[:k * B].view(k, B, L) * delta[None]
                phi[d0:d0 + k] = contrib.mean(1)
        finally:
            eng.grad_rules(False, False)
        return phi.t().contiguous()[None].cpu().numpy()
