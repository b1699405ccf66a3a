"""CPU, world_size 2 over gloo: coalition sharding + the single all-gather + replicated solve."""
import os
import sys

import numpy as np
import torch
import torch.distributed as td
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, K, M, D, q):
    sys.path.insert(0, ROOT)
    from oracle.kernelshap_ref import KernelExplainerRef
    from shap_transformer_asr_b200 import dist as wdist
    from shap_transformer_asr_b200.kernelshap import sample_coalitions

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    wdist.init_from_env(device_type="cpu")
    assert wdist.rank_world() == (rank, world)
    rng = np.random.default_rng(0)
    lin = rng.standard_normal((M, D))

    def f(Z):
        return np.asarray(Z, dtype=np.float64) @ lin + 0.1 * np.asarray(Z).sum(1, keepdims=True) ** 2

    Z, kw, _ = sample_coalitions(M, K, seed=0)           # identical on every rank (same seed)
    lo, hi = wdist.shard_range(Z.shape[0], rank, world)
    y_local = torch.from_numpy(f(Z[lo:hi]))              # this rank's rows only
    y = wdist.all_gather_rows(y_local, Z.shape[0], rank, world).numpy()
    ex = KernelExplainerRef(f, M)
    np.random.seed(0)
    ex.sample(K)
    phi = ex.solve(y, f(np.ones((1, M)))[0], f(np.zeros((1, M)))[0])
    q.put((rank, lo, hi, float(np.abs(y - f(Z)).max()), phi))
    wdist.barrier()
    td.destroy_process_group()


def test_two_rank_sharding_and_all_gather():
    K, M, D = 301, 24, 3      # K not divisible by the world size: ragged last shard
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, K, M, D, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, e0, phi0), (r1, lo1, hi1, e1, phi1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 151, 151, 301)
    assert e0 < 1e-12 and e1 < 1e-12              # gathered matrix == unsharded evaluation
    assert np.array_equal(phi0, phi1)             # replicated solve agrees bit for bit


def _eg_worker(rank, world, port, L, D, q):
    sys.path.insert(0, ROOT)
    from shap_transformer_asr_b200 import dist as wdist
    from shap_transformer_asr_b200.expected_gradients import ExpectedGradientsExplainer, make_background

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    wdist.init_from_env(device_type="cpu")
    a = np.linspace(0.5, 2.0, D).astype(np.float32)

    class Mock:                                  # f_j(x) = 0.5 a_j |x|^2: gradient a_j x
        device = torch.device("cpu")
        calls = 0

        def num_frames(self, n):
            return D

        def grad_waveforms(self, xs, frames):
            Mock.calls += xs.shape[0]
            aj = torch.from_numpy(a[np.asarray(frames)]).float()
            return aj[:, None] * xs, 0.5 * aj * (xs ** 2).sum(1)

    x = np.random.default_rng(1).standard_normal(L).astype(np.float32)
    bg = make_background(L, 5, seed=2)
    phi = ExpectedGradientsExplainer(Mock(), bg, nsamples=16, seed=3, batch=7).shap_values(x)
    ref = ExpectedGradientsExplainer(Mock(), bg, nsamples=16, seed=3, batch=7, shard_outputs=False).shap_values(x)
    q.put((rank, Mock.calls, phi, ref))
    wdist.barrier()
    td.destroy_process_group()


def test_expected_gradients_shard_output_frames_over_ranks():
    """Two ranks: each evaluates the passes of its own block of output frames (no collective on the data path), one
    all-gather assembles [1, L, D]; bit-identical to the unsharded estimator on every rank."""
    L, D = 50, 9              # D not divisible by the world size
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_eg_worker, args=(r, 2, port, L, D, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, c0, phi0, ref0), (_, c1, phi1, ref1) = res
    assert phi0.shape == (1, L, D)
    assert np.array_equal(phi0, phi1) and np.array_equal(phi0, ref0) and np.array_equal(ref0, ref1)
    # sharded run: 5 + 4 output frames x 16 samples; the unsharded reference run adds D x 16 on each rank
    assert (c0, c1) == (5 * 16 + D * 16, 4 * 16 + D * 16)


def test_shard_range_edge_cases():
    from shap_transformer_asr_b200.dist import shard_range
    for K in (0, 1, 7, 8, 9, 2048):
        for G in (1, 2, 4, 8):
            spans = [shard_range(K, r, G) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == K
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
