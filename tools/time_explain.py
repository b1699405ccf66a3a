#!/usr/bin/env python
"""Where the fixed cost of one explained clip goes (C2, one rank of an 8-GPU split: 257 of the 2050 rows), phase by phase
with a device synchronisation after each -- a diagnosis aid, not a bench line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shap_transformer_asr_b200 import MODELS, WORKLOADS, Engine, char_targets, sample_coalitions, synthetic_clip
from shap_transformer_asr_b200.modelzoo import build_random_init_model


def main():
    wl = WORKLOADS["C2"]
    cfg = MODELS[wl.model]
    eng = Engine(build_random_init_model(cfg, seed=0), cfg, max_batch=0)
    clip = synthetic_clip(wl.num_samples)
    M = wl.num_segments
    out = {}

    def phase(name, fn, reps=5):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        torch.cuda.synchronize()
        out[name] = round((time.perf_counter() - t0) / reps * 1e3, 3)
        return r

    phase("set_clip", lambda: eng.set_clip(clip, num_segments=M))
    eng.set_targets("logits")
    ones = eng.bits_to_device(np.ones((1, M), np.uint8)).clone()
    lg = phase("target_forward_1_row", lambda: eng.eval_bits(ones))
    phase("target_forward_d2h", lambda: lg.cpu())
    logits = lg.view(-1, cfg.vocab_size).cpu().numpy()
    phase("sampler_host", lambda: sample_coalitions(M, wl.num_coalitions, seed=0, packed=True), reps=3)
    words, kw, _ = sample_coalitions(M, wl.num_coalitions, seed=0, packed=True)
    frames, tokens = phase("char_targets_host", lambda: char_targets(logits))
    phase("set_targets", lambda: eng.set_targets("logprob", frames, tokens))
    bits = phase("bits_upload", lambda: eng.bits_to_device(words)).clone()
    y = phase("eval_257_rows", lambda: eng.eval_bits(bits[:257]))
    yall = phase("eval_2048_rows", lambda: eng.eval_bits(bits), reps=2)
    w_dev = torch.from_numpy(kw).cuda()
    fx, fnull = yall[1].double(), yall[0].double()
    phase("wls", lambda: eng.wls(bits, w_dev, yall, fx, fnull, M))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
