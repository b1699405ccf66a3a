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


def test_fused_attention_backward_recurrence_matches_autograd():
    """attention_bwd_kernel (attention_bwd.cu) restated in numpy, tile by tile: key blocks j outside, query blocks i inside,
    P = exp2(S c - LSE_i) from the forward's log-sum-exp (no row maximum in the backward), D_i = dO_i . O_i,
    dS = P (dP - D_i) scale, dV_j += P^T dO_i, dK_j += dS^T Q_i, dQ_i += dS K_j (one partial per key block), P and dS
    rounded to bf16 before they feed the contractions, padding keys masked to P = 0.  Against torch autograd on
    softmax(Q K^T / sqrt d) V with an upstream gradient dO."""
    rng = np.random.default_rng(3)
    T, d, BLK = 300, 64, 128                    # three blocks, the last one ragged (44 valid rows / keys)
    q, k, v, dO = (_bf16(rng.standard_normal((T, d)) * s) for s in (1.0, 1.0, 1.0, 0.5))
    scale = 1.0 / math.sqrt(d)
    tq, tk, tv = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (q, k, v))
    out = torch.softmax(tq @ tk.T * scale, -1) @ tv
    out.backward(torch.tensor(dO, dtype=torch.float64))
    # what the forward kernel leaves behind: O (bf16) and the log2-domain LSE of the scaled scores
    c = scale * math.log2(math.e)
    S = q.astype(np.float64) @ k.astype(np.float64).T
    lse2 = np.log2(np.exp2(S * c - (S * c).max(1, keepdims=True)).sum(1)) + (S * c).max(1)
    O = _bf16(out.detach().numpy())
    delta = (dO.astype(np.float64) * O).sum(1)
    nb = (T + BLK - 1) // BLK
    pad = nb * BLK

    def padded(a):
        z = np.zeros((pad, d), np.float32)
        z[:T] = a
        return z

    qp, kp, vp, dop = padded(q), padded(k), padded(v), padded(dO)
    lse_p = np.full(pad, 1e30)
    lse_p[:T] = lse2
    del_p = np.zeros(pad)
    del_p[:T] = delta
    dq = np.zeros((pad, d), np.float32)
    dk = np.zeros((pad, d), np.float32)
    dv = np.zeros((pad, d), np.float32)
    for j in range(nb):
        ks, vs = kp[j * BLK:(j + 1) * BLK], vp[j * BLK:(j + 1) * BLK]
        nvalid = T - j * BLK
        for i in range(nb):
            sl = slice(i * BLK, (i + 1) * BLK)
            s_ij = qp[sl] @ ks.T                                  # tensor cores, fp32 accumulate
            dp_ij = dop[sl] @ vs.T
            p = np.exp2(s_ij * np.float32(c) - lse_p[sl, None]).astype(np.float32)
            p[:, max(nvalid, 0):] = 0.0                           # keys past the clip
            ds = (p * (dp_ij - del_p[sl, None]) * scale).astype(np.float32)
            pb, dsb = _bf16(p), _bf16(ds)
            dv[j * BLK:(j + 1) * BLK] += pb.T @ dop[sl]
            dk[j * BLK:(j + 1) * BLK] += dsb.T @ qp[sl]
            dq[sl] += dsb @ ks
    for mine, ref, name in ((dq[:T], tq.grad, "dQ"), (dk[:T], tk.grad, "dK"), (dv[:T], tv.grad, "dV")):
        r = ref.numpy()
        err = np.abs(mine - r).max() / np.abs(r).max()
        assert err < 6e-3, (name, err)          # bf16 rounding of P / dS / O, nothing else
    assert np.all(dq[T:] == 0) and np.all(dk[T:] == 0) and np.all(dv[T:] == 0)     # padded rows receive nothing
