"""CPU: the oracle restatement against the real `transformers` modules and the committed golden vectors."""
import os

import numpy as np
import pytest
import torch

from helpers import VARIANTS, build_model, weight_checksum
from oracle import callback as CB
from oracle import w2v2_forward as W
from shap_transformer_asr_b200 import preprocess, targets
from shap_transformer_asr_b200.config import MODELS

# golden L -> T' pairs pinned by the reference's saved notebook outputs:
# evaluation.ipynb:460,463; shap_value_test.ipynb:301,342,501; visualize_shap_data.ipynb:228;
# audio_amplification_wav2vec2_test.py:116; test_shap_asr.py:86
REFERENCE_FRAME_COUNTS = [(183600, 573), (199760, 624), (90240, 281), (77040, 240), (16000, 49), (93680, 292)]


@pytest.mark.parametrize("L,T", REFERENCE_FRAME_COUNTS)
def test_frame_count_matches_reference_notebooks(L, T):
    assert MODELS["wav2vec2-base"].num_frames(L) == T


def test_frame_counts_of_baseline_configs():
    assert MODELS["wav2vec2-base"].conv_lengths(80000) == [15999, 7999, 3999, 1999, 999, 499, 249]
    assert MODELS["wav2vec2-large"].num_frames(160000) == 499


@pytest.mark.parametrize("name", list(VARIANTS))
def test_oracle_matches_transformers_live(name):
    cfg = VARIANTS[name]
    model = build_model(cfg)
    x = torch.randn(2, 3000, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        ref = model(x).logits
        out = W.ctc_logits(W.state_dict_of(model), cfg.to_dict(), x)
    assert ref.shape == out.shape
    assert (ref - out).abs().max().item() <= 2e-4


@pytest.mark.parametrize("name", list(VARIANTS))
def test_oracle_matches_golden(name, golden_dir):
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    cfg = VARIANTS[name]
    model = build_model(cfg)
    assert abs(weight_checksum(model) - float(g["checksum"])) <= 1e-6 * float(g["checksum"]), "weight RNG drift"
    with torch.no_grad():
        out = W.ctc_logits(W.state_dict_of(model), cfg.to_dict(), torch.from_numpy(g["x"])).numpy()
    assert np.abs(out - g["logits"]).max() <= 2e-4


def test_oracle_c1_base_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "c1_base.npz"))
    cfg = MODELS["wav2vec2-base"]
    model = build_model(cfg)
    assert abs(weight_checksum(model) - float(g["checksum"])) <= 1e-6 * float(g["checksum"])
    clip = preprocess.synthetic_clip(16000)
    assert np.allclose(clip[:64], g["clip_head"], atol=1e-6)
    bounds = CB.segment_bounds(16000, 32)
    X = CB.materialize(clip, g["rows"][:4], bounds)
    with torch.no_grad():
        out = W.ctc_logits(W.state_dict_of(model), cfg.to_dict(), torch.from_numpy(X)).numpy()
    assert np.abs(out - g["logits"][:4]).max() <= 5e-4
    frames, tokens = CB.char_targets(g["logits"][0])
    assert np.array_equal(frames, g["frames"]) and np.array_equal(tokens, g["tokens"])


def test_reductions_and_callback_shapes():
    cfg = VARIANTS["tiny_group"]
    model = build_model(cfg)
    sd = W.state_dict_of(model)
    x = np.random.default_rng(0).standard_normal((3, 2000)).astype(np.float32)
    with torch.no_grad():
        logits = model(torch.from_numpy(x)).logits
    # ModelWrapper.forward semantics, shap_calculation.py:50
    assert np.allclose(CB.evaluate(sd, cfg.to_dict(), x, "max"), logits.max(-1).values.numpy(), atol=1e-4)
    # lime_predict_fn semantics, lime_shap_wav2vec2_comparison.py:68-70
    assert np.allclose(CB.evaluate(sd, cfg.to_dict(), x, "mean"), logits.mean(-1).numpy().mean(1, keepdims=True), atol=1e-4)
    # predict_function semantics, w2v2conformer.py:40-42
    out = CB.evaluate(sd, cfg.to_dict(), x, "logit", [2], [5])
    assert out.shape == (3, 1) and np.allclose(out[:, 0], logits[:, 2, 5].numpy(), atol=1e-4)
    lp = CB.evaluate(sd, cfg.to_dict(), x, "logprob", [1, 3], [4, 7])
    assert np.allclose(lp, torch.log_softmax(logits, -1)[:, [1, 3], [4, 7]].numpy(), atol=1e-4)


def test_masker_and_materialize_edge_cases():
    x = np.arange(10, dtype=np.float32) + 1
    keep = np.array([1, 1, 0, 0, 1, 0, 1, 1, 1, 0])
    assert np.array_equal(CB.masker(x, keep), x * keep)
    assert np.array_equal(CB.masker(x, keep, baseline=-2.0), np.where(keep, x, -2.0))
    # ragged segmentation: M does not divide L
    b = CB.segment_bounds(10, 3)
    assert b.tolist() == [0, 3, 6, 10] and np.array_equal(b, preprocess.segment_bounds(10, 3))
    Xm = CB.materialize(x, np.array([[1, 0, 1], [0, 0, 0], [1, 1, 1]]), b)
    assert Xm[0].tolist() == [1, 2, 3, 0, 0, 0, 7, 8, 9, 10]
    assert not Xm[1].any() and np.array_equal(Xm[2], x)
    # normalisation: HF feature_extraction_wav2vec2.py:78-97
    n = CB.normalize_clip(x)
    assert abs(n.mean()) < 1e-6 and abs(n.var() - 1) < 1e-5
    assert np.array_equal(n, preprocess.normalize_clip(x))


def test_char_targets_follow_visualization_rule():
    # ids: blank a a | b blank b c c   -> new non-blank, non-'|' tokens start at frames 1, 4, 6, 7
    ids = [0, 5, 5, 4, 6, 0, 6, 7, 7]
    logits = np.full((len(ids), 32), -1.0)
    logits[np.arange(len(ids)), ids] = 1.0
    for fn in (CB.char_targets, targets.char_targets):
        f, t = fn(logits)
        assert f.tolist() == [1, 4, 6, 7] and t.tolist() == [5, 6, 6, 7]
    # degenerate transcript (all blank): every frame with its argmax
    f, t = targets.char_targets(np.tile(np.eye(32)[0], (5, 1)))
    assert f.tolist() == [0, 1, 2, 3, 4] and t.tolist() == [0] * 5
    assert CB.first_char_target(logits) == targets.first_char_target(logits) == (1, 5)


def test_pack_coalitions_roundtrip():
    rng = np.random.default_rng(0)
    for M in (1, 31, 32, 33, 100, 200):
        Z = rng.integers(0, 2, size=(17, M))
        w = preprocess.pack_coalitions(Z)
        assert w.shape == (17, (M + 31) // 32) and w.dtype == np.uint32
        back = (w[:, np.arange(M) // 32] >> (np.arange(M) % 32).astype(np.uint32)) & 1
        assert np.array_equal(back, Z)
