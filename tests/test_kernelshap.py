"""CPU: KernelSHAP restatement (oracle, parity unpinned vs shap itself) -- properties, and the product
sampler against the oracle sampler bit for bit."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle.kernelshap_ref import KernelExplainerRef, brute_force_shapley
from shap_transformer_asr_b200.kernelshap import expand_to_samples, sample_coalitions


@pytest.mark.parametrize("M,K", [(32, 256), (100, 2048), (200, 8192), (128, 2048), (12, 300), (8, "auto"), (5, 100), (3, 50), (2, 10)])
def test_product_sampler_is_bit_identical_to_oracle(M, K):
    np.random.seed(0)
    Zr, wr = KernelExplainerRef(None, M).sample(K)
    Z, w, info = sample_coalitions(M, K, seed=0)
    assert Z.shape == Zr.shape and np.array_equal(Z, Zr.astype(np.uint8))
    assert np.array_equal(w, wr)


def test_sampler_regression_pins(golden_dir):
    pins = np.load(os.path.join(golden_dir, "sampler_pins.npz"))
    for key in pins.files:
        M, K = key[1:].split("_K")
        K = K if K == "auto" else int(K)
        Z, w, _ = sample_coalitions(int(M), K, seed=0)
        p = pins[key]
        assert Z.shape[0] == int(p[0])
        assert int(hashlib.sha256(Z.tobytes()).hexdigest()[:12], 16) == int(p[1])
        assert abs(w.sum() - p[2]) < 1e-12


def test_sampler_structure_matches_survey_appendix_a():
    # only size 1 (+complements) is enumerated at the BASELINE configs,
    # with normalised weight 0.2563 / 0.1951 / 0.1711 for M = 32 / 100 / 200
    for M, K, nfixed, w1 in [(32, 256, 64, 0.2563), (100, 2048, 200, 0.1951), (200, 8192, 400, 0.1711)]:
        Z, w, info = sample_coalitions(M, K, seed=0)
        assert info["n_fixed"] == nfixed and Z.shape[0] == K
        assert abs(w[:nfixed].sum() - w1) < 5e-5
        assert abs(w.sum() - 1.0) < 1e-9
        sizes = Z[:nfixed].sum(1)
        assert set(sizes.tolist()) == {1, M - 1}
        # the fixed rows follow itertools.combinations order, each followed by its complement
        assert np.array_equal(Z[0], np.eye(M, dtype=np.uint8)[0]) and np.array_equal(Z[1], 1 - Z[0])
        assert len({r.tobytes() for r in Z}) == K  # de-duplicated


def test_seed_determinism_and_sensitivity():
    a = sample_coalitions(64, 1000, seed=3)[0]
    b = sample_coalitions(64, 1000, seed=3)[0]
    c = sample_coalitions(64, 1000, seed=4)[0]
    assert np.array_equal(a, b) and not np.array_equal(a, c)


def _game(M, D, seed=0, interactions=True):
    rng = np.random.default_rng(seed)
    lin = rng.standard_normal((M, D))
    pair = rng.standard_normal((M, M, D)) * (0.3 if interactions else 0.0)
    base = rng.standard_normal(D)

    def f(Z):
        Z = np.atleast_2d(np.asarray(Z, dtype=np.float64))
        return base + Z @ lin + np.einsum("ki,kj,ijd->kd", Z, Z, pair)

    return f, lin


def test_full_enumeration_equals_brute_force_shapley():
    M, D = 7, 3
    f, _ = _game(M, D)
    phi, fx, fnull = KernelExplainerRef(f, M).shap_values(nsamples=10 ** 6)
    assert np.abs(phi - brute_force_shapley(f, M)).max() < 1e-9
    assert np.abs(phi.sum(0) - (fx - fnull)).max() < 1e-9


def test_additive_game_is_exact_under_sampling():
    M, D = 40, 4
    f, lin = _game(M, D, interactions=False)
    np.random.seed(1)
    phi, fx, fnull = KernelExplainerRef(f, M).shap_values(nsamples=600)
    assert np.abs(phi - lin).max() < 1e-8


def test_efficiency_holds_for_sampled_nonadditive_game():
    M, D = 30, 2
    f, _ = _game(M, D)
    np.random.seed(2)
    phi, fx, fnull = KernelExplainerRef(f, M).shap_values(nsamples=500)
    assert np.abs(phi.sum(0) - (fx - fnull)).max() < 1e-9


def test_expand_to_samples_layout():
    # reference on-disk layout [1, L, T'] (shap_calculation.py:200-210, evaluation.ipynb:503-504)
    phi = np.arange(6, dtype=np.float64).reshape(3, 2)
    b = np.array([0, 2, 3, 6])
    out = expand_to_samples(phi, b)
    assert out.shape == (1, 6, 2) and out[0, :, 0].tolist() == [0, 0, 2, 4, 4, 4]
    per = expand_to_samples(phi, b, per_sample=True)
    assert np.allclose(per[0].sum(0), phi.sum(0))


class _MockEngine:
    """CPU stand-in with the Engine surface KernelShapExplainer.explain uses: an additive game over M segments
    (output d of a coalition = bias_d + sum of the kept segments' weights), so the Shapley values are the weights."""

    def __init__(self, M, D, vocab=8, frames=6):
        import types
        rng = np.random.default_rng(7)
        self.num_segments = M
        self.device = torch.device("cpu")
        self.config = types.SimpleNamespace(vocab_size=vocab)
        self.W = rng.standard_normal((M, D))
        self.bias = rng.standard_normal(D)
        self.logit_table = rng.standard_normal((frames, vocab))
        self.mode, self.sel = None, None
        self.calls = []

    def set_clip(self, clip, num_segments, baseline=0.0):
        assert num_segments == self.num_segments
        self.calls.append("set_clip")

    def set_targets(self, mode, frames=None, tokens=None):
        self.mode = mode
        self.sel = None if frames is None else (np.asarray(frames), np.asarray(tokens))
        self.calls.append(f"set_targets:{mode}")

    def out_width(self):
        return self.logit_table.size if self.mode == "logits" else self.W.shape[1]

    def bits_to_device(self, Z):
        from shap_transformer_asr_b200.preprocess import pack_coalitions
        Z = np.asarray(Z)
        words = Z if Z.dtype == np.uint32 else pack_coalitions(Z)       # same contract as Engine.bits_to_device
        return torch.from_numpy(np.ascontiguousarray(words).view(np.int32))

    def _unpack(self, bits):
        w = bits.numpy().view(np.uint32)
        return np.stack([(w[:, m // 32] >> (m % 32)) & 1 for m in range(self.num_segments)], 1).astype(np.float64)

    def eval_bits(self, bits):
        self.calls.append(f"eval:{self.mode}:{bits.shape[0]}")
        if self.mode == "logits":
            return torch.from_numpy(np.tile(self.logit_table.reshape(1, -1), (bits.shape[0], 1)).astype(np.float32))
        return torch.from_numpy((self._unpack(bits) @ self.W + self.bias).astype(np.float32))

    def wls(self, bits, w, y, fx, fnull, M):
        ks = KernelExplainerRef(None, M)
        ks.maskMatrix, ks.kernelWeights, ks.nsamplesAdded = self._unpack(bits), w.numpy(), bits.shape[0]
        return torch.from_numpy(ks.solve(y.numpy(), fx.numpy(), fnull.numpy())), torch.zeros(1, dtype=torch.int32)


@pytest.mark.parametrize("targets_given", [False, True])
def test_explain_orchestration_on_a_mock_engine(targets_given):
    """explain(): target selection (overlapped with the host sampler), evaluation of [empty, full, Z], device-solve call:
    on an additive game the attributions are the segment weights, and the sampled rows are the seeded sampler's."""
    from shap_transformer_asr_b200.kernelshap import KernelShapExplainer, sample_coalitions
    M, D = 12, 5
    eng = _MockEngine(M, D)
    tg = (np.arange(D, dtype=np.int32), np.zeros(D, dtype=np.int32)) if targets_given else None
    res = KernelShapExplainer(eng, nsamples=300, seed=3).explain(np.zeros(1000, np.float32), num_segments=M, targets=tg)
    Z, kw, _ = sample_coalitions(M, 300, seed=3)
    assert np.array_equal(res["Z"], Z) and np.array_equal(res["weights"], kw)
    assert np.abs(res["phi"].numpy() - eng.W).max() < 1e-5
    assert np.abs(res["fx"].numpy() - (eng.W.sum(0) + eng.bias)).max() < 1e-5
    assert eng.calls[0] == "set_clip" and eng.calls[-1] == f"eval:logprob:{Z.shape[0] + 2}"
    if not targets_given:
        assert "eval:logits:1" in eng.calls and len(res["frames"]) > 0


def test_native_sampler_continues_numpys_global_generator():
    """The per-draw permutations run natively (csrc/sampler.cu) on np.random's own MT19937 state: after a call the global
    generator must be exactly where shap's pure-numpy loop would have left it, for unseeded continuation too."""
    from shap_transformer_asr_b200.kernelshap import sample_coalitions, unpack_coalitions
    from shap_transformer_asr_b200.preprocess import pack_coalitions
    for M, K in [(7, 40), (33, 200), (100, 2048), (257, 300), (2048, 40)]:
        np.random.seed(123)
        np.random.random(17)                               # generator somewhere in the middle of its state block
        Z1, w1, _ = sample_coalitions(M, K, seed=None)
        Z2, w2, _ = sample_coalitions(M, K, seed=None)      # continues from the state the first call left
        tail = np.random.random(3)
        ref = KernelExplainerRef(lambda z: z.sum(1, keepdims=True), M)
        np.random.seed(123)
        np.random.random(17)
        Za, wa = ref.sample(K)
        ref2 = KernelExplainerRef(lambda z: z.sum(1, keepdims=True), M)
        Zb, wb = ref2.sample(K)
        assert np.array_equal(Z1, Za.astype(np.uint8)) and np.array_equal(w1, wa)
        assert np.array_equal(Z2, Zb.astype(np.uint8)) and np.array_equal(w2, wb)
        assert np.array_equal(tail, np.random.random(3))
        words, w3, _ = sample_coalitions(M, K, seed=5, packed=True)
        Z3, _, _ = sample_coalitions(M, K, seed=5)
        assert np.array_equal(words, pack_coalitions(Z3)) and np.array_equal(unpack_coalitions(words, M), Z3)


def test_expected_gradients_estimator_on_a_mock_engine():
    """ExpectedGradientsExplainer (the reference's GradientExplainer loop, shap_calculation.py:125-162) on a quadratic
    toy model whose input gradients are known in closed form: layout [1, L, D], and for f_j(x) = 0.5 a_j |x|^2 the
    expected-gradients attribution of sample i is E[a_j (bg + alpha (x - bg))_i (x - bg)_i]."""
    import types
    from shap_transformer_asr_b200.expected_gradients import ExpectedGradientsExplainer, make_background

    L, T = 50, 4
    a = np.array([1.0, -2.0, 0.5, 3.0])

    class Mock:
        device = torch.device("cpu")

        def num_frames(self, n):
            return T

        def grad_waveforms(self, xs, frames):          # one target frame per row
            aj = torch.from_numpy(a[np.broadcast_to(np.asarray(frames), (xs.shape[0],))]).float()
            return aj[:, None] * xs, 0.5 * aj * (xs ** 2).sum(1)

    x = np.random.default_rng(0).standard_normal(L).astype(np.float32)
    bg = make_background(L, 5, seed=2)
    ex = ExpectedGradientsExplainer(Mock(), bg, nsamples=4000, seed=1, batch=512)
    phi = ex.shap_values(x)
    assert phi.shape == (1, L, T)
    # E over alpha ~ U(0,1) and the 5 backgrounds of a_j (bg + alpha d) d,  d = x - bg
    d = x[None] - bg
    expect = np.stack([a[j] * ((bg + 0.5 * d) * d).mean(0) for j in range(T)], 1)
    assert np.abs(phi[0] - expect).max() < 0.05 * np.abs(expect).max()
    # completeness: sum_i phi_ij = f_j(x) - E f_j(bg) for this model, up to sampling noise
    gap = phi[0].sum(0) - (0.5 * a * (x ** 2).sum() - 0.5 * a * (bg ** 2).sum(1).mean())
    assert np.abs(gap).max() < 0.05 * np.abs(0.5 * a * (x ** 2).sum()).max()


def test_deeplift_explainer_pairs_rows_and_averages_on_a_mock_engine():
    """DeepLiftExplainer (the reference's shap.DeepExplainer + custom_shap_handlers.py loop) on a mock engine: every device
    call carries [explained | reference] halves of at most GRAD_TILE_ROWS rows, the rules are switched on for the calls and
    off afterwards, and phi = mean over the background of gradient x (x - reference) in the layout [1, L, D]."""
    import torch
    from shap_transformer_asr_b200.deeplift import DeepLiftExplainer
    L, D, B = 40, 7, 5
    a = np.linspace(0.5, 2.0, D).astype(np.float32)
    log = []

    class Mock:
        device = torch.device("cpu")
        GRAD_TILE_ROWS = 32
        rules = (False, False)

        def num_frames(self, n):
            return D

        def grad_rules(self, rescale_silu=False, glu_placeholder=False):
            Mock.rules = (rescale_silu, glu_placeholder)

        def grad_waveforms(self, rows, frames):
            n = rows.shape[0]
            assert Mock.rules == (True, False) and n % 2 == 0 and n <= 32
            h = n // 2
            assert np.array_equal(np.asarray(frames)[:h], np.asarray(frames)[h:])
            log.append(n)
            # a toy "rule-modified gradient": a_j * (x + ref) for the explained half, zero for the reference half
            aj = torch.from_numpy(a[np.asarray(frames)[:h]]).float()
            g = torch.zeros_like(rows)
            g[:h] = aj[:, None] * (rows[:h] + rows[h:])
            return g, torch.zeros(n)

    rng = np.random.default_rng(0)
    x = rng.standard_normal(L).astype(np.float32)
    bg = rng.standard_normal((B, L)).astype(np.float32)
    phi = DeepLiftExplainer(Mock(), bg).shap_values(x)
    assert phi.shape == (1, L, D) and Mock.rules == (False, False)
    want = np.stack([(a[j] * (x[None] + bg) * (x[None] - bg)).mean(0) for j in range(D)], 1)
    assert np.allclose(phi[0], want, rtol=1e-5, atol=1e-6)
    assert log == [30, 30, 10]          # three output frames x five pairs per call, then the last frame
