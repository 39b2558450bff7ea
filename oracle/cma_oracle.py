"""CPU oracle for the CMA-ES rank-mu covariance update.  TEST INFRASTRUCTURE (see nes_oracle.py header).

PARITY UNPINNED.  The reference delegates all CMA arithmetic to the third-party package `cma`
(PyPI "cma", N. Hansen's pycma; version unpinned — README.md:11 links PyPI without a version, there is
no requirements file), called at cma_es.py:43-49 (options/ctor), :62 (ask) and :90 (tell).  `cma` is not
in /root/reference, not installed, and cannot be fetched (no network); the reference holds no tests or
golden vectors at that boundary.  This file therefore restates the PUBLISHED algorithm — Hansen, "The CMA
Evolution Strategy: A Tutorial", arXiv:1604.00772 (cited at README.md:16), equations (47)-(58) — in
fp64 numpy, and "matches the reference" can only mean "matches this restatement".  The pinned contract
is the positive-weights (mu = floor(lambda/2)) recombination; active (negative) weights are an option.
"""
from __future__ import annotations

import numpy as np


def cma_constants(n, lam, active=False):
    """Default strategy parameters, tutorial Table 1 (eqs. 49-58)."""
    wp = np.log((lam + 1) / 2.0) - np.log(np.arange(1, lam + 1))            # eq. 49
    mu = lam // 2
    mu_eff = wp[:mu].sum() ** 2 / (wp[:mu] ** 2).sum()
    mu_eff_neg = wp[mu:].sum() ** 2 / max((wp[mu:] ** 2).sum(), 1e-300)
    alpha_cov = 2.0
    cc = (4 + mu_eff / n) / (n + 4 + 2 * mu_eff / n)                        # eq. 56
    c1 = alpha_cov / ((n + 1.3) ** 2 + mu_eff)                              # eq. 57
    cmu = min(1 - c1, alpha_cov * (0.25 + mu_eff + 1 / mu_eff - 2) / ((n + 2) ** 2 + alpha_cov * mu_eff / 2))  # eq. 58
    cs = (mu_eff + 2) / (n + mu_eff + 5)                                    # eq. 55
    ds = 1 + 2 * max(0.0, np.sqrt((mu_eff - 1) / (n + 1)) - 1) + cs
    w = np.zeros(lam)
    w[:mu] = wp[:mu] / wp[:mu].sum()                                        # eq. 53, positive part
    if active:
        a_mu = 1 + c1 / cmu                                                 # eq. 50
        a_mueff = 1 + 2 * mu_eff_neg / (mu_eff + 2)                         # eq. 51
        a_posdef = (1 - c1 - cmu) / (n * cmu)                               # eq. 52
        w[mu:] = min(a_mu, a_mueff, a_posdef) * wp[mu:] / (-wp[mu:].sum())  # eq. 53, negative part
    return dict(w=w, mu=mu, mu_eff=mu_eff, cc=cc, c1=c1, cmu=cmu, cs=cs, ds=ds)


def eigen_gap(n, c1, cmu):
    """Generations between eigendecompositions of C: the tutorial's lazy update (its reference code B.2 refreshes B, D when
    counteval - eigeneval > lambda/(c1+cmu)/n/10, i.e. every 1/((c1+cmu) n 10) generations), at least every generation.
    For BASELINE configs[2] (n=1024, lambda=256) and configs[4] (n=4096, lambda=1024) this is 1: lambda ~ n/4 makes cmu large."""
    return max(1, int(1.0 / ((c1 + cmu) * n * 10.0)))


def sort_and_scale(X, cost, m_old, sigma):
    """y_{i:lambda} = (x_{i:lambda} - m_old)/sigma, members sorted by cost ascending (eq. 41-42);
    ties by index.  (cma_es.py:89 rank-shapes the costs first; ranks preserve the order.)"""
    order = np.argsort(np.asarray(cost, dtype=np.float64), kind='stable')
    Y = (np.asarray(X, dtype=np.float64)[order] - np.asarray(m_old, dtype=np.float64)) / float(sigma)
    return Y, order


def rank_mu_delta(Y_sorted, w):
    """dC_mu = sum_i w_i y_i y_i^T   (eq. 47, rank-mu term).  Y_sorted [lambda, n], w [lambda]."""
    Y = np.asarray(Y_sorted, dtype=np.float64)
    return (Y * np.asarray(w, dtype=np.float64)[:, None]).T @ Y


def cov_update(C, dC, pc, c1, cmu, sum_w, hsig=1.0, cc=0.0):
    """C <- (1 + c1*dh - c1 - cmu*sum_w) C + c1 pc pc^T + cmu dC  (eq. 47), dh = (1-hsig) cc (2-cc)."""
    dh = (1 - hsig) * cc * (2 - cc)
    decay = 1 + c1 * dh - c1 - cmu * sum_w
    C = np.asarray(C, dtype=np.float64)
    pc = np.asarray(pc, dtype=np.float64)
    return decay * C + c1 * np.outer(pc, pc) + cmu * np.asarray(dC, dtype=np.float64), decay


def sphere(X):
    return (np.asarray(X, dtype=np.float64) ** 2).sum(axis=-1)


# ----------------------------------------------------------------------------------------------------------------
# One full CMA-ES generation — still a restatement of the tutorial, still parity-UNPINNED.
# ----------------------------------------------------------------------------------------------------------------
class CMAState:
    """(mu/mu_w, lambda)-CMA-ES, tutorial arXiv:1604.00772 Fig. 6 / eqs. 38-47 with the Table 1 defaults, positive
    weights, c_m = 1, eigendecomposition refreshed every generation.  Stands in for cma.CMAEvolutionStrategy
    (cma_es.py:49) with ask() (:62) and tell() (:90).  fp64 throughout."""

    def __init__(self, x0, sigma0, lam):
        self.n = n = len(x0)
        self.lam = lam
        self.k = cma_constants(n, lam)
        self.m = np.asarray(x0, dtype=np.float64).copy()
        self.sigma = float(sigma0)
        self.C = np.eye(n)
        self.pc = np.zeros(n)
        self.ps = np.zeros(n)
        self.B = np.eye(n)
        self.D = np.ones(n)
        self.gen = 0
        self.chiN = np.sqrt(n) * (1 - 1.0 / (4 * n) + 1.0 / (21 * n * n))
        self.gap = eigen_gap(n, self.k['c1'], self.k['cmu'])

    def ask(self, z):
        """x_i = m + sigma * B D z_i (eq. 38-40); z [lambda, n] standard normal."""
        self.Y = (np.asarray(z, dtype=np.float64) * self.D) @ self.B.T
        return self.m + self.sigma * self.Y

    def tell(self, X, cost):
        k, n = self.k, self.n
        w, mu_eff, cc, c1, cmu, cs, ds = k['w'], k['mu_eff'], k['cc'], k['c1'], k['cmu'], k['cs'], k['ds']
        Y, order = sort_and_scale(X, cost, self.m, self.sigma)
        yw = w @ Y                                                       # eq. 41
        self.m = self.m + self.sigma * yw                                # eq. 42 (c_m = 1)
        Cinvsqrt_yw = self.B @ ((self.B.T @ yw) / self.D)
        self.ps = (1 - cs) * self.ps + np.sqrt(cs * (2 - cs) * mu_eff) * Cinvsqrt_yw          # eq. 43
        norm_ps = np.linalg.norm(self.ps)
        hsig = float(norm_ps / np.sqrt(1 - (1 - cs) ** (2 * (self.gen + 1))) / self.chiN < 1.4 + 2.0 / (n + 1))
        self.pc = (1 - cc) * self.pc + hsig * np.sqrt(cc * (2 - cc) * mu_eff) * yw            # eq. 45
        dC = rank_mu_delta(Y, w)
        self.C, self.decay = cov_update(self.C, dC, self.pc, c1, cmu, w.sum(), hsig=hsig, cc=cc)   # eq. 47
        self.dC = dC
        self.sigma = self.sigma * np.exp((cs / ds) * (norm_ps / self.chiN - 1))               # eq. 44
        self.C = 0.5 * (self.C + self.C.T)
        self.gen += 1
        if self.gen % self.gap == 0:                                     # lazy eigendecomposition (tutorial B.2)
            d2, self.B = np.linalg.eigh(self.C)
            self.D = np.sqrt(np.maximum(d2, 1e-300))
        return order
