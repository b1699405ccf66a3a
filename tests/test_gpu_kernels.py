"""GPU: individual kernels through the C ABI against torch fp32 / numpy references."""
import numpy as np
import pytest
import torch

from helpers import rel_err
from oracle import callback as CB
from oracle.kernelshap_ref import KernelExplainerRef

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import shap_transformer_asr_b200 as pkg
    assert torch.cuda.is_available(), "GPU tests need a B200"
    return pkg


GEMM_SHAPES = [
    (128, 256, 64), (128, 256, 512), (300, 512, 1536), (1000, 768, 512), (996, 2304, 768),
    (4096, 3072, 768), (777, 768, 3072), (130, 128, 128), (64, 64, 192), (257, 48, 6144), (500, 32, 512),
    (19999, 512, 1024),
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("tc", [True, False], ids=["tcgen05", "simt"])
def test_gemm_matches_torch(P, M, N, K, tc):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ w.float().t() + bias
    out = P.debug_gemm(a, w, bias, act=0, out_fp32=True, tcgen05=tc)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    print(f"gemm {M}x{N}x{K} tc={tc}: max abs err {err:.3e}")
    assert err < 2e-3 * max(1.0, ref.abs().max().item())
    # GELU epilogue + bf16 output
    out2 = P.debug_gemm(a, w, bias, act=1, out_fp32=False, tcgen05=tc).float()
    ref2 = torch.nn.functional.gelu(ref)
    assert (out2 - ref2).abs().max().item() < 2e-2 * max(1.0, ref2.abs().max().item())


def test_gemm_tcgen05_equals_validation_kernel_closely(P):
    a = torch.randn(515, 1024, device="cuda").bfloat16()
    w = (torch.randn(768, 1024, device="cuda") / 32).bfloat16()
    o1 = P.debug_gemm(a, w, None, tcgen05=True)
    o2 = P.debug_gemm(a, w, None, tcgen05=False)
    assert (o1 - o2).abs().max().item() < 1e-3


@pytest.fixture(scope="module")
def tiny_engine(P):
    from helpers import TINY, build_model
    model = build_model(TINY)
    eng = P.Engine(model, TINY, max_batch=8)
    yield eng
    eng.close()


@pytest.mark.parametrize("L,M", [(4000, 8), (4001, 7), (16000, 32), (80000, 100), (1237, 200)])
def test_mask_kernel_is_bit_exact(P, tiny_engine, L, M):
    rng = np.random.default_rng(L)
    clip = rng.standard_normal(L).astype(np.float32)
    Z = rng.integers(0, 2, size=(13, M)).astype(np.uint8)
    Z[0] = 0
    Z[1] = 1
    for baseline in (0.0, -1.5):
        tiny_engine.set_clip(clip, num_segments=M, baseline=baseline)
        out = tiny_engine.mask(tiny_engine.bits_to_device(Z)).cpu().numpy()
        ref = CB.materialize(clip, Z, CB.segment_bounds(L, M), baseline)
        assert np.array_equal(out, ref)


@pytest.mark.parametrize("M,K,D", [(8, 24, 3), (32, 256, 45), (100, 2048, 249), (200, 8192, 64)])
def test_wls_matches_oracle_solve(P, tiny_engine, M, K, D):
    rng = np.random.default_rng(M)
    lin = rng.standard_normal((M, D))

    def f(Z):
        Z = np.asarray(Z, dtype=np.float64)
        return (Z @ lin + 0.05 * Z.sum(1, keepdims=True) ** 1.5).astype(np.float32)

    Z, kw, _ = P.sample_coalitions(M, K, seed=0)
    y = f(Z)
    fx, fnull = f(np.ones((1, M)))[0].astype(np.float64), f(np.zeros((1, M)))[0].astype(np.float64)
    ref = KernelExplainerRef(f, M)
    np.random.seed(0)
    ref.sample(K)
    phi_ref = ref.solve(y, fx, fnull)
    dev = tiny_engine.device
    phi, status = tiny_engine.wls(tiny_engine.bits_to_device(Z), torch.from_numpy(kw).to(dev),
                                  torch.from_numpy(y).to(dev), torch.from_numpy(fx).to(dev),
                                  torch.from_numpy(fnull).to(dev), M)
    phi = phi.cpu().numpy()
    assert int(status.item()) == 0
    print(f"wls M={M} K={K} D={D}: max abs err {np.abs(phi - phi_ref).max():.3e}")
    assert np.abs(phi - phi_ref).max() < 1e-8 * max(1.0, np.abs(phi_ref).max())
    assert np.abs(phi.sum(0) - (fx - fnull)).max() < 1e-9      # efficiency


@pytest.mark.parametrize("M,K,D", [(6, 4, 2), (40, 17, 5), (200, 64, 33)])
def test_wls_singular_design_gives_the_lstsq_solution(P, tiny_engine, M, K, D):
    """Fewer distinct coalitions than unknowns: shap's solve falls back to numpy.linalg.lstsq on the sqrt-weighted system
    (minimum-norm least squares).  The device path reports status 2 and returns the same."""
    rng = np.random.default_rng(M + K)
    Z = (rng.random((K, M)) < 0.5).astype(np.uint8)
    kw = rng.random(K) + 0.1
    y = rng.standard_normal((K, D)).astype(np.float32)
    fx, fnull = rng.standard_normal(D), rng.standard_normal(D)
    dev = tiny_engine.device
    phi, status = tiny_engine.wls(tiny_engine.bits_to_device(Z), torch.from_numpy(kw).to(dev), torch.from_numpy(y).to(dev),
                                  torch.from_numpy(fx).to(dev), torch.from_numpy(fnull).to(dev), M)
    assert int(status.item()) == 2
    Zf = Z.astype(np.float64)
    X = Zf[:, :-1] - Zf[:, -1:]
    sw = np.sqrt(kw)
    ref = np.empty((M, D))
    for d in range(D):
        r = y[:, d].astype(np.float64) - fnull[d] - Zf[:, -1] * (fx[d] - fnull[d])
        w = np.linalg.lstsq(sw[:, None] * X, sw * r, rcond=None)[0]
        ref[:-1, d] = w
        ref[-1, d] = (fx[d] - fnull[d]) - w.sum()
    err = np.abs(phi.cpu().numpy() - ref).max()
    print(f"singular wls M={M} K={K}: max abs err vs lstsq {err:.3e}")
    assert err < 1e-7 * max(1.0, np.abs(ref).max())


# shapes with >= 296 output tiles take the CTA-pair kernel (TMA-store / TMA reduce-add epilogues); the small one the
# single-CTA kernel (direct stores)
@pytest.mark.parametrize("M,N,K", [(19999, 512, 1024), (37848, 768, 768), (9960, 1024, 4096), (1000, 768, 512)])
def test_gemm_residual_epilogues_match_torch(P, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    lin = a.float() @ w.float().t() + bias
    tol = 2e-3 * max(1.0, lin.abs().max().item())
    # (1) bf16 residual, fp32 output (post-LN wav2vec2 without the fused LayerNorm residual)
    res16 = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    out = P.debug_gemm(a, w, bias, residual=res16)
    assert (out - (lin + res16.float())).abs().max().item() < tol
    # (2) fp32 residual into a different fp32 buffer, alpha = 0.5 (conformer macaron feed-forward)
    res32 = torch.randn(M, N, device="cuda", generator=g)
    out = P.debug_gemm(a, w, bias, residual=res32, alpha=0.5)
    assert (out - (0.5 * lin + res32)).abs().max().item() < tol
    # (3) in-place accumulation into the fp32 residual stream: TMA reduce-add in the pair kernel
    acc = res32.clone()
    P.debug_gemm(a, w, bias, alpha=0.5, accumulate_into=acc)
    torch.cuda.synchronize()
    err = (acc - (0.5 * lin + res32)).abs().max().item()
    print(f"gemm {M}x{N}x{K} in-place accumulate: max abs err {err:.3e}")
    assert err < tol
    # (4) and twice in a row (stable-LN / conformer layers chain such accumulations)
    P.debug_gemm(a, w, None, accumulate_into=acc)
    assert (acc - (0.5 * lin + res32 + (lin - bias))).abs().max().item() < 2 * tol


def test_eval_edge_cases_empty_single_and_ragged(P, tiny_engine):
    """Empty coalition matrix, a single row, one segment, M not dividing L, non-zero baseline, ragged last batch tile."""
    rng = np.random.default_rng(11)
    clip = rng.standard_normal(4001).astype(np.float32)
    tiny_engine.set_clip(clip, num_segments=7, baseline=0.25)
    tiny_engine.set_targets("max")
    empty = tiny_engine.eval_bits(torch.empty((0, 1), dtype=torch.int32, device="cuda"))
    assert tuple(empty.shape) == (0, tiny_engine.num_frames(4001))
    Z = rng.integers(0, 2, size=(19, 7)).astype(np.uint8)          # 19 rows, batch tile 8 -> tiles of 8, 8, 3
    y = tiny_engine.eval_bits(tiny_engine.bits_to_device(Z))
    one = tiny_engine.eval_bits(tiny_engine.bits_to_device(Z[5:6]))
    assert torch.equal(y[5:6], one)
    # explicit materialisation through the waveform entry point gives the same numbers
    X = torch.from_numpy(CB.materialize(clip, Z, CB.segment_bounds(4001, 7), 0.25)).cuda()
    assert torch.equal(y, tiny_engine.eval_waveforms(X))
    # a single segment: the two coalitions are "all baseline" and "the clip"
    tiny_engine.set_clip(clip, num_segments=1, baseline=0.0)
    y1 = tiny_engine.eval_bits(tiny_engine.bits_to_device(np.array([[0], [1]], np.uint8)))
    ref = tiny_engine.eval_waveforms(torch.stack([torch.zeros(4001), torch.from_numpy(clip)]).cuda())
    assert torch.equal(y1, ref)
    # errors are reported, not swallowed
    tiny_engine.set_targets("logprob", [10 ** 6], [3])
    with pytest.raises(RuntimeError, match="target frame"):
        tiny_engine.eval_bits(tiny_engine.bits_to_device(Z[:1, :1]))
    # the frame check follows the clip of the call, not the workspace a previous call with another length left behind
    T_clip = tiny_engine.num_frames(4001)
    long_x = torch.from_numpy(rng.standard_normal((1, 9000)).astype(np.float32)).cuda()
    T_long = tiny_engine.num_frames(9000)
    tiny_engine.set_targets("logit", [T_long - 1], [3])
    assert tiny_engine.eval_waveforms(long_x).shape == (1, 1)         # valid for the long rows
    with pytest.raises(RuntimeError, match="target frame"):            # ... but beyond the set clip's T'
        tiny_engine.eval_bits(tiny_engine.bits_to_device(np.ones((1, 1), np.uint8)))
    tiny_engine.set_targets("logit", [T_clip - 1], [3])
    a = tiny_engine.eval_bits(tiny_engine.bits_to_device(np.ones((1, 1), np.uint8)))
    b = tiny_engine.eval_waveforms(torch.from_numpy(clip)[None].cuda())
    assert torch.equal(a, b)
    with pytest.raises(RuntimeError):
        tiny_engine.set_clip(clip[:200], num_segments=2)          # shorter than the conv receptive field


def test_mean_mode_is_bit_reproducible(P, tiny_engine):
    """lime_predict_fn's reduction (mean over vocabulary and time): fixed summation order, no atomics."""
    rng = np.random.default_rng(3)
    x = torch.from_numpy(rng.standard_normal((11, 6000)).astype(np.float32)).cuda()
    tiny_engine.set_targets("mean")
    a = tiny_engine.eval_waveforms(x)
    for _ in range(3):
        assert torch.equal(a, tiny_engine.eval_waveforms(x))
    tiny_engine.set_targets("logits")
    lg = tiny_engine.eval_waveforms(x).view(11, -1, 32)
    assert (a[:, 0] - lg.mean(-1).mean(1)).abs().max().item() < 1e-5


def test_graph_replay_equals_eager_launches(P):
    """The tile plan replayed as a CUDA graph gives bit-identical outputs to launching its kernels one by one, across
    changes of targets, mode and clip between replays (per-call arguments travel through the device argument block)."""
    from helpers import TINY, build_model
    model = build_model(TINY)
    rng = np.random.default_rng(8)
    clip = rng.standard_normal(8000).astype(np.float32)
    Z = rng.integers(0, 2, size=(21, 9)).astype(np.uint8)
    outs = []
    for graphs in (True, False):
        eng = P.Engine(model, TINY, max_batch=8, graphs=graphs)
        eng.set_clip(clip, num_segments=9)
        res = []
        for mode, fr, tk in (("max", None, None), ("logprob", [3, 5, 20], [1, 2, 31]), ("mean", None, None),
                             ("logit", [0], [7]), ("logits", None, None)):
            eng.set_targets(mode, fr, tk)
            res.append(eng.eval_bits(eng.bits_to_device(Z)).clone())
            res.append(eng.eval_bits(eng.bits_to_device(Z)).clone())       # replay with unchanged arguments
        eng.set_clip(clip[::-1].copy(), num_segments=3, baseline=0.5)          # same length: same plans, new arguments
        eng.set_targets("max")
        res.append(eng.eval_bits(eng.bits_to_device(Z[:, :3])).clone())
        n0 = eng.launch_count()
        eng.eval_bits(eng.bits_to_device(Z[:, :3]))
        assert eng.launch_count() - n0 > 3 * 20                                # 3 tiles x (plan kernels + argument kernel)
        outs.append(res)
        eng.close()
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_vocab_not_a_multiple_of_32_runs_on_the_contraction_kernel(P):
    """lm_head rows are zero-padded to a multiple of 32: fine-tuned checkpoints with e.g. 29 or 45 tokens need no other
    code path (the fused CUDA-core head kernel is a validation mode only)."""
    import dataclasses
    from helpers import TINY, build_model
    from oracle import w2v2_forward as W
    for V in (29, 45):
        cfg = dataclasses.replace(TINY, vocab_size=V)
        model = build_model(cfg)
        x = np.random.default_rng(V).standard_normal((3, 5000)).astype(np.float32)
        with torch.no_grad():
            ref = W.ctc_logits(W.state_dict_of(model), cfg.to_dict(), torch.from_numpy(x)).numpy()
        eng = P.Engine(model, cfg, max_batch=4)
        eng.set_targets("logits")
        out = eng.eval_waveforms(torch.from_numpy(x).cuda()).view(ref.shape).cpu().numpy()
        assert rel_err(out, ref) < 0.025
        eng.set_targets("logprob", [1, 4], [V - 1, 0])
        lp = eng.eval_waveforms(torch.from_numpy(x).cuda()).cpu().numpy()
        ref_lp = torch.log_softmax(torch.from_numpy(ref), -1)[:, [1, 4], [V - 1, 0]].numpy()
        assert np.abs(lp - ref_lp).max() < 0.025 * np.abs(ref).max()
        eng.close()
