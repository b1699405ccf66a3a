"""CPU models of two numerical constructions the CUDA kernels rely on (the kernels themselves are covered by the GPU
parity tests): the bf16 hi/lo K = 32 split of conv0_mma_kernel and the one-MUFU GELU of common.cuh."""
import math

import numpy as np
import torch


def _bf16(x):
    return torch.from_numpy(np.asarray(x, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def test_conv0_hi_lo_split_reproduces_conv_plus_groupnorm_affine():
    """conv0_stats_kernel / conv0_mma_kernel (frontend.cu): y[f, c] = sum_k A[f, k] B[c, k] with
    A[f] = [x_hi | x_lo | x_hi | 1 1], B[c] = [w_hi | w_hi | w_lo | s_hi s_lo], w = filter * rstd * gamma, s = shift.
    bf16 x bf16 products are exact in fp32, so the only error is the dropped lo*lo term (~2^-16 relative per product)."""
    rng = np.random.default_rng(0)
    T0, C, KW, stride = 400, 64, 10, 5
    x = rng.standard_normal((T0 - 1) * stride + KW).astype(np.float32)
    w = (rng.standard_normal((C, KW)) * 0.3).astype(np.float32)
    ga = rng.uniform(0.5, 3.0, C).astype(np.float32)      # rstd * gamma
    gb = rng.standard_normal(C).astype(np.float32)        # beta - mean * rstd * gamma
    win = np.stack([x[f * stride:f * stride + KW] for f in range(T0)])          # [T0, KW]
    ref = win.astype(np.float64) @ (w.astype(np.float64) * ga[:, None]).T + gb   # [T0, C]

    x_hi = _bf16(win); x_lo = _bf16(win - x_hi)
    wf = (w.astype(np.float64) * ga[:, None]).astype(np.float32)
    w_hi = _bf16(wf); w_lo = _bf16(wf - w_hi)
    s_hi = _bf16(gb); s_lo = _bf16(gb - s_hi)
    A = np.concatenate([x_hi, x_lo, x_hi, np.ones((T0, 2), np.float32)], 1)
    B = np.concatenate([w_hi, w_hi, w_lo, s_hi[:, None], s_lo[:, None]], 1)
    assert A.shape[1] == 32 and B.shape[1] == 32
    y = (A.astype(np.float32) @ B.T.astype(np.float32)).astype(np.float32)       # fp32 accumulate
    scale = np.abs(win).astype(np.float64) @ np.abs(wf).astype(np.float64).T + np.abs(gb)
    assert (np.abs(y - ref) / scale).max() < 3e-5        # far below the bf16 rounding (4e-3) of the stored output


def test_one_mufu_gelu_matches_exact_erf_gelu():
    """gelu_erf2 (common.cuh): max(v, 0) - |v| q with q = 0.5 erfc(|v| / sqrt 2) = (c0 + c1 |v| + ... + c6 |v|^6)^-16
    (Abramowitz-Stegun 7.1.28 with the 1/sqrt 2 and the 0.5 folded into the coefficients), evaluated in fp32."""
    f = np.float32
    c = [f(1.0442737824), f(5.2075163037e-02), f(2.2076998457e-02), f(3.4227392389e-03), f(3.9686137011e-05),
         f(5.1055209009e-05), f(5.6212996640e-06)]
    v = np.linspace(-12, 12, 400001).astype(f)
    av = np.abs(v)
    p = np.full_like(av, c[6])
    for k in range(5, -1, -1):
        p = (p * av + c[k]).astype(f)
    for _ in range(4):
        p = (p * p).astype(f)
    out = (np.maximum(v, f(0)) - av * (f(1) / p).astype(f)).astype(f)
    v64 = v.astype(np.float64)
    ref = 0.5 * v64 * (1.0 + np.vectorize(math.erf)(v64 / math.sqrt(2.0)))
    assert np.abs(out - ref).max() < 1.5e-6
    big = np.abs(ref) > 1e-3
    assert (np.abs(out - ref)[big] / np.abs(ref)[big]).max() < 5e-4   # bf16 rounding of the result is 4e-3
