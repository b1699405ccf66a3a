"""CPU: the "next" rows f1 / f2 -- eta_raw against golden vectors produced by the REAL reference functions
(calculate_metric.py:74-149 imported; nraw_vs_wer.py:20-62 extracted), WER / CTC decode, test-set builder, file layout."""
import os

import numpy as np
import pytest

from oracle.make_golden_metrics import case
from shap_transformer_asr_b200 import metrics, sweep
from shap_transformer_asr_b200.kernelshap import expand_to_samples


def test_eta_raw_matches_reference_golden_vectors(golden_dir):
    rows = np.load(os.path.join(golden_dir, "eta_raw.npz"))["rows"]
    assert len(rows) >= 12
    seen_fractional = 0
    for seed, L, T, lo, hi, seg_ms, pct, ref_a, ref_b in rows:
        clean, noise, shap = case(int(seed), int(L), int(T), int(lo), int(hi))
        a = metrics.eta_raw(clean, noise, shap, 16000, segment_ms=seg_ms, percentile=pct, itm_ratio=0.5)
        b = metrics.eta_raw(clean, noise, shap, 16000, segment_ms=seg_ms, percentile=pct, itm_ratio=1.0)
        assert a == pytest.approx(ref_a, abs=1e-12) and b == pytest.approx(ref_b, abs=1e-12)
        seen_fractional += 0 < ref_a < 1
    assert seen_fractional >= 4      # the fixture is not all 0 / 1


def test_eta_raw_edge_cases():
    rng = np.random.default_rng(0)
    clean, noise, shap = rng.standard_normal(3200), rng.standard_normal(3200), rng.standard_normal((3200, 4))
    # [T', L] input is transposed like calculate_metric.py:92-95
    assert metrics.eta_raw(clean, noise, shap.T, 16000) == metrics.eta_raw(clean, noise, shap, 16000)
    with pytest.raises(ValueError):
        metrics.eta_raw(clean, noise, shap, 16000, segment_ms=0.01)          # 0 samples per segment
    with pytest.raises(ValueError):
        metrics.eta_raw(clean, noise, rng.standard_normal((10, 4)), 16000)   # incompatible shape
    assert metrics.eta_raw(clean, noise, np.ones((3200, 2)), 16000) == 0.0   # nothing above the threshold
    assert metrics.eta_raw(clean[:100], noise[:100], shap[:100], 16000) == 0.0   # shorter than one segment


def test_wer_and_ctc_decode():
    assert metrics.wer("the cat sat", "the cat sat") == 0.0
    assert metrics.wer("the cat sat", "the bat sat") == pytest.approx(1 / 3)
    assert metrics.wer("the cat sat", "the sat") == pytest.approx(1 / 3)            # deletion
    assert metrics.wer("the cat sat", "the big cat sat") == pytest.approx(1 / 3)    # insertion
    assert metrics.wer("a b c d", "") == 1.0 and metrics.wer("", "") == 0.0
    # ids: H H <pad> E | | C A A <pad> A T  ->  "HE CAAT"
    ids = [11, 11, 0, 5, 4, 4, 19, 7, 7, 0, 7, 6]
    assert metrics.greedy_ctc_decode(ids) == "HE CAAT"
    assert metrics.greedy_ctc_decode([0, 0, 4, 0]) == ""


def test_test_set_builder_follows_reference_layout():
    ts = sweep.make_test_set(num_clips=2, num_samples=100000, seed=3)
    assert [t["type"] for t in ts] == ["clean", "noisy", "noisy", "noisy"] * 2         # shap_calculation.py:79-105
    assert [t["snr"] for t in ts[:4]] == [float("inf"), 5, 2, 1]
    for t in ts:
        assert len(t["audio"]) >= 100000 and t["noise"].shape == t["audio"].shape      # :75
    clean, noisy = ts[0], ts[1]
    assert not clean["noise"].any()
    snr = 10 * np.log10(np.mean(clean["audio"] ** 2) / np.mean(noisy["noise"] ** 2))
    assert abs(snr - 5) < 0.1                                                            # :55-60
    assert np.allclose(noisy["audio"] - clean["audio"], noisy["noise"])
    again = sweep.make_test_set(num_clips=2, num_samples=100000, seed=3)
    assert all(np.array_equal(a["audio"], b["audio"]) for a, b in zip(ts, again))       # seeded, unlike the reference


def test_saved_layout_is_what_the_reference_consumers_expect(tmp_path):
    # visualization.py:337-344 requires squeeze(shap) of shape (L, T'); calculate_metric.py reads the same files
    phi = np.random.default_rng(0).standard_normal((8, 5))
    bounds = np.array([0, 10, 20, 30, 40, 50, 60, 70, 83])
    arr = expand_to_samples(phi, bounds).astype(np.float32)
    assert arr.shape == (1, 83, 5)
    p = tmp_path / "shap_values_sample_1_clean_inf.npy"
    np.save(p, arr)
    back = np.load(p).squeeze()
    assert back.shape == (83, 5) and np.allclose(back[25], phi[2], atol=1e-6)


def test_eta_raw_on_segment_level_attributions_is_bit_identical():
    """eta_raw_segments (no [L, T'] expansion) against eta_raw on the expanded matrix, both ITM thresholds."""
    from shap_transformer_asr_b200 import eta_raw, eta_raw_segments, expand_to_samples, segment_bounds
    rng = np.random.default_rng(3)
    L, M, D = 102400, 128, 40
    clean = rng.standard_normal(L) * (1 + np.sin(np.arange(L) / 3000.0))
    noise = 0.8 * rng.standard_normal(L)
    phi = rng.standard_normal((M, D)) * rng.random((M, 1))
    b = segment_bounds(L, M)
    full = expand_to_samples(phi, b)[0]
    for ratio in (0.5, 1.0):
        assert eta_raw_segments(clean, noise, phi, b, 16000, itm_ratio=ratio) == eta_raw(clean, noise, full, 16000, itm_ratio=ratio)
