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
