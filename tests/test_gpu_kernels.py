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


def test_wls_reports_singular_design(P, tiny_engine):
    M, K, D = 6, 4, 2       # fewer samples than unknowns -> normal matrix not positive definite
    Z = np.eye(M, dtype=np.uint8)[:K]
    dev = tiny_engine.device
    phi, status = tiny_engine.wls(tiny_engine.bits_to_device(Z), torch.ones(K, dtype=torch.float64, device=dev),
                                  torch.zeros(K, D, device=dev), torch.zeros(D, dtype=torch.float64, device=dev),
                                  torch.zeros(D, dtype=torch.float64, device=dev), M)
    assert int(status.item()) == 1


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
    with pytest.raises(RuntimeError):
        tiny_engine.set_clip(clip[:200], num_segments=2)          # shorter than the conv receptive field
