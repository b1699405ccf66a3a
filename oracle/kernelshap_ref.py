"""ORACLE (test infrastructure only) -- PARITY UNPINNED.

CPU restatement of ``shap.KernelExplainer`` (coalition sampler + constrained weighted
least squares).  ``shap`` is a third-party dependency of the reference pinned only as
``shap>=0.40.0`` (requirements.txt:4); it is absent from /root/reference, from this image
and from /opt/wheelhouse, and the reference itself never calls KernelExplainer (its call
sites are GradientExplainer shap_calculation.py:133,162 and DeepExplainer
feasability_tests/w2v2conformer.py:139-142).  There is therefore NO golden vector, test
or fixture of the reference, and no runnable copy of shap, to pin this file against:
it follows the published algorithm of ``shap/explainers/_kernel.py`` (0.41-0.46:
``KernelExplainer.explain`` / ``addsample`` / ``run`` / ``solve``) as restated in
SURVEY.md Appendix A, and is checked only by properties (efficiency, exactness on
additive games, brute-force Shapley under full enumeration; tests/test_kernelshap.py).

Conventions used here (segment formulation): features are the M segment indicators,
background ``data = zeros[1, M]``, instance ``x = ones[1, M]``, so ``synth_data`` equals the
mask matrix and the model function is the coalition callback ``f(Z[n, M]) -> [n, D]``.
``l1_reg`` is pinned to ``False`` and the link is the identity.
The random part draws from the GLOBAL legacy ``np.random`` state, in shap's call order,
so ``np.random.seed(s)`` fixes the coalition index sets.
"""
from __future__ import annotations

import copy
import itertools

import numpy as np
from scipy.special import binom


class KernelExplainerRef:
    def __init__(self, model_fn, num_features: int):
        self.f = model_fn
        self.M = int(num_features)

    # -- shap: KernelExplainer.allocate / addsample ------------------------------------
    def _allocate(self):
        self.maskMatrix = np.zeros((self.nsamples, self.M))
        self.kernelWeights = np.zeros(self.nsamples)
        self.nsamplesAdded = 0

    def _addsample(self, m, w):
        self.maskMatrix[self.nsamplesAdded, :] = m
        self.kernelWeights[self.nsamplesAdded] = w
        self.nsamplesAdded += 1

    # -- shap: KernelExplainer.explain, sampling part -------------------------------------
    def sample(self, nsamples="auto"):
        M = self.M
        self.nsamples = 2 * M + 2 ** 11 if nsamples == "auto" else int(nsamples)
        self.max_samples = 2 ** 30
        if M <= 30:
            self.max_samples = 2 ** M - 2
            if self.nsamples > self.max_samples:
                self.nsamples = self.max_samples
        self._allocate()

        num_subset_sizes = int(np.ceil((M - 1) / 2.0))
        num_paired_subset_sizes = int(np.floor((M - 1) / 2.0))
        weight_vector = np.array([(M - 1.0) / (i * (M - i)) for i in range(1, num_subset_sizes + 1)])
        weight_vector[:num_paired_subset_sizes] *= 2
        weight_vector /= np.sum(weight_vector)

        num_full_subsets = 0
        num_samples_left = self.nsamples
        group_inds = np.arange(M, dtype="int64")
        mask = np.zeros(M)
        remaining_weight_vector = copy.copy(weight_vector)
        for subset_size in range(1, num_subset_sizes + 1):
            nsubsets = binom(M, subset_size)
            if subset_size <= num_paired_subset_sizes:
                nsubsets *= 2
            if num_samples_left * remaining_weight_vector[subset_size - 1] / nsubsets >= 1.0 - 1e-8:
                num_full_subsets += 1
                num_samples_left -= nsubsets
                if remaining_weight_vector[subset_size - 1] < 1.0:
                    remaining_weight_vector /= (1 - remaining_weight_vector[subset_size - 1])
                w = weight_vector[subset_size - 1] / binom(M, subset_size)
                if subset_size <= num_paired_subset_sizes:
                    w /= 2.0
                for inds in itertools.combinations(group_inds, subset_size):
                    mask[:] = 0.0
                    mask[np.array(inds, dtype="int64")] = 1.0
                    self._addsample(mask, w)
                    if subset_size <= num_paired_subset_sizes:
                        mask[:] = np.abs(mask - 1)
                        self._addsample(mask, w)
            else:
                break

        nfixed_samples = self.nsamplesAdded
        samples_left = self.nsamples - self.nsamplesAdded
        if num_full_subsets != num_subset_sizes:
            remaining_weight_vector = copy.copy(weight_vector)
            remaining_weight_vector[:num_paired_subset_sizes] /= 2
            remaining_weight_vector = remaining_weight_vector[num_full_subsets:]
            remaining_weight_vector /= np.sum(remaining_weight_vector)
            ind_set = np.random.choice(len(remaining_weight_vector), 4 * samples_left, p=remaining_weight_vector)
            ind_set_pos = 0
            used_masks = {}
            while samples_left > 0 and ind_set_pos < len(ind_set):
                mask.fill(0.0)
                ind = ind_set[ind_set_pos]
                ind_set_pos += 1
                subset_size = ind + num_full_subsets + 1
                mask[np.random.permutation(M)[:subset_size]] = 1.0
                mask_tuple = tuple(mask)
                new_sample = False
                if mask_tuple not in used_masks:
                    new_sample = True
                    used_masks[mask_tuple] = self.nsamplesAdded
                    samples_left -= 1
                    self._addsample(mask, 1.0)
                else:
                    self.kernelWeights[used_masks[mask_tuple]] += 1.0
                if samples_left > 0 and subset_size <= num_paired_subset_sizes:
                    mask[:] = np.abs(mask - 1)
                    if new_sample:
                        samples_left -= 1
                        self._addsample(mask, 1.0)
                    else:
                        self.kernelWeights[used_masks[mask_tuple] + 1] += 1.0
            weight_left = np.sum(weight_vector[num_full_subsets:])
            self.kernelWeights[nfixed_samples:] *= weight_left / self.kernelWeights[nfixed_samples:].sum()
        self.num_full_subsets = num_full_subsets
        self.nfixed_samples = nfixed_samples
        return self.maskMatrix[: self.nsamplesAdded].copy(), self.kernelWeights[: self.nsamplesAdded].copy()

    # -- shap: KernelExplainer.solve (l1_reg=False) ----------------------------------------
    def solve(self, ey, fx, fnull):
        """ey[nsamplesAdded, D], fx[D], fnull[D] -> phi[M, D]."""
        M = self.M
        Zm = self.maskMatrix[: self.nsamplesAdded]
        kw = self.kernelWeights[: self.nsamplesAdded]
        ey = np.asarray(ey, dtype=np.float64)
        D = ey.shape[1]
        phi = np.zeros((M, D))
        for d in range(D):
            eyAdj = ey[:, d] - fnull[d]
            eyAdj2 = eyAdj - Zm[:, -1] * (fx[d] - fnull[d])
            etmp = np.transpose(np.transpose(Zm[:, :-1]) - Zm[:, -1])
            WX = kw[:, None] * etmp
            try:
                w = np.linalg.solve(etmp.T @ WX, WX.T @ eyAdj2)
            except np.linalg.LinAlgError:
                sqrt_W = np.sqrt(kw)
                w = np.linalg.lstsq(sqrt_W[:, None] * etmp, sqrt_W * eyAdj2, rcond=None)[0]
            phi[:-1, d] = w
            phi[-1, d] = (fx[d] - fnull[d]) - sum(w)
        phi[np.abs(phi) < 1e-10] = 0
        return phi

    # -- shap: KernelExplainer.shap_values for one instance -------------------------------------
    def shap_values(self, nsamples="auto"):
        fnull = np.atleast_1d(np.asarray(self.f(np.zeros((1, self.M))), dtype=np.float64)[0])
        fx = np.atleast_1d(np.asarray(self.f(np.ones((1, self.M))), dtype=np.float64)[0])
        Z, _ = self.sample(nsamples)
        ey = np.asarray(self.f(Z), dtype=np.float64)
        if ey.ndim == 1:
            ey = ey[:, None]
        return self.solve(ey, fx, fnull), fx, fnull


def brute_force_shapley(f, M: int) -> np.ndarray:
    """Exact Shapley values by enumeration (test helper for small M)."""
    from math import factorial

    allZ = np.array(list(itertools.product([0, 1], repeat=M)), dtype=np.float64)
    vals = np.asarray(f(allZ), dtype=np.float64)
    if vals.ndim == 1:
        vals = vals[:, None]
    index = {tuple(z): i for i, z in enumerate(allZ)}
    phi = np.zeros((M, vals.shape[1]))
    for i in range(M):
        for z in allZ:
            if z[i] == 0:
                s = int(z.sum())
                wgt = factorial(s) * factorial(M - s - 1) / factorial(M)
                z1 = z.copy()
                z1[i] = 1
                phi[i] += wgt * (vals[index[tuple(z1)]] - vals[index[tuple(z)]])
    return phi
