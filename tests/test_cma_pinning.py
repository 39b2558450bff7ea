"""Pins for the CMA-ES restatement (oracle/cma_oracle.py) that do not come from the restatement itself.  pycma — the
package the reference delegates to (cma_es.py:43-49,62,90) — is absent, so these are the independent anchors available:

  * strategy constants worked by hand from the published formulas (Hansen, arXiv:1604.00772, Table 1 / eqs. 49-58), and
    the (mu_w, w_1) pairs pycma prints in its start-up banner for its default population sizes
    ("(5_w,10)-aCMA-ES (mu_w=3.2,w_1=45%)" for dimension 10, etc.);
  * one complete generation in dimension 3 with lambda = 4 carried out by hand below (mean, evolution paths, rank-one and
    rank-mu terms, step size), written out as literal arithmetic rather than through the oracle's helpers.
CPU only."""
import math

import numpy as np

from oracle import cma_oracle as co


def test_default_weights_match_pycma_banner_values():
    # (n, default lambda = 4 + floor(3 ln n), mu, mu_w as printed to one decimal, w_1 in per cent)
    for n, lam, mu, mu_w, w1 in [(2, 6, 3, 2.0, 63), (10, 10, 5, 3.2, 45), (20, 12, 6, 3.7, 40), (40, 15, 7, 4.5, 34)]:
        assert lam == 4 + int(3 * math.log(n))
        k = co.cma_constants(n, lam)
        assert k['mu'] == mu
        assert round(k['mu_eff'], 1) == mu_w, (n, k['mu_eff'])
        assert int(100 * k['w'][0]) == w1, (n, k['w'][0])
        assert abs(k['w'][:mu].sum() - 1) < 1e-15 and np.all(k['w'][mu:] == 0) and np.all(np.diff(k['w'][:mu]) < 0)


def test_constants_by_hand_n10():
    """n = 10, lambda = 10, mu = 5: every constant from literal arithmetic (four significant digits worked on paper)."""
    lw = [math.log(5.5) - math.log(i) for i in range(1, 6)]          # 1.7047 1.0116 0.6061 0.3185 0.0953
    assert [round(v, 4) for v in lw] == [1.7047, 1.0116, 0.6061, 0.3185, 0.0953]
    s = sum(lw)
    w = [v / s for v in lw]
    assert [round(v, 4) for v in w] == [0.4563, 0.2708, 0.1622, 0.0852, 0.0255]
    mu_eff = 1.0 / sum(v * v for v in w)
    assert round(mu_eff, 3) == 3.167
    k = co.cma_constants(10, 10)
    assert np.allclose(k['w'][:5], w, rtol=1e-14) and abs(k['mu_eff'] - mu_eff) < 1e-12
    assert abs(k['cc'] - (4 + mu_eff / 10) / (10 + 4 + 2 * mu_eff / 10)) < 1e-15 and round(k['cc'], 4) == 0.2950
    assert abs(k['cs'] - (mu_eff + 2) / (10 + mu_eff + 5)) < 1e-15 and round(k['cs'], 4) == 0.2844
    assert abs(k['c1'] - 2 / ((10 + 1.3) ** 2 + mu_eff)) < 1e-15 and round(k['c1'], 5) == 0.01528
    # eq. 58 of the 2016 tutorial (the 1/4 is pycma's `rankmu_offset`): 2 (0.25 + 3.1673 - 2 + 0.3157) / 147.167 = 0.023552
    assert abs(k['cmu'] - 2 * (0.25 + mu_eff - 2 + 1 / mu_eff) / ((10 + 2) ** 2 + mu_eff)) < 1e-15 and round(k['cmu'], 5) == 0.02355
    assert abs(k['ds'] - (1 + k['cs'])) < 1e-15                      # sqrt((mu_eff-1)/(n+1)) < 1: the max() term vanishes


def test_one_generation_by_hand_n3():
    """n = 3, lambda = 4 (mu = 2), C = I, sigma = 0.5, m = (1, 2, 3), sphere costs.  All quantities below are literal."""
    n, lam = 3, 4
    z = np.array([[1.0, 0.0, -1.0], [0.5, -0.5, 0.0], [-1.0, 1.0, 1.0], [0.0, 2.0, 0.0]])
    m0, sigma0 = np.array([1.0, 2.0, 3.0]), 0.5
    st = co.CMAState(m0, sigma0, lam)
    X = st.ask(z)
    assert np.array_equal(X, m0 + sigma0 * z)                         # B = D = I
    cost = (X ** 2).sum(1)                                            # 2.25+4+6.25=12.5 | 1.5625+3.0625+9=13.625 | .25+6.25+12.25=18.75 | 1+9+9=19
    assert np.allclose(cost, [12.5, 13.625, 18.75, 19.0])
    order = st.tell(X, cost)
    assert list(order) == [0, 1, 2, 3]
    # weights: ln(2.5) - ln(1) = 0.91629, ln(2.5) - ln(2) = 0.22314 -> w = (0.80415, 0.19585); mu_eff = 1/(w1^2+w2^2) = 1.4599
    w1 = math.log(2.5) / (math.log(2.5) + math.log(1.25))
    w2 = 1 - w1
    assert round(w1, 4) == 0.8042 and round(w2, 4) == 0.1958
    mu_eff = 1 / (w1 * w1 + w2 * w2)
    assert round(mu_eff, 3) == 1.460
    yw = w1 * z[0] + w2 * z[1]                                        # (0.90208, -0.09792, -0.80415)
    assert np.allclose(yw, [0.90208, -0.09792, -0.80416], atol=1e-5)
    assert np.allclose(st.m, m0 + 0.5 * yw, rtol=1e-15)
    cs = (mu_eff + 2) / (n + mu_eff + 5)
    cc = (4 + mu_eff / n) / (n + 4 + 2 * mu_eff / n)
    c1 = 2 / ((n + 1.3) ** 2 + mu_eff)
    cmu = min(1 - c1, 2 * (0.25 + mu_eff - 2 + 1 / mu_eff) / ((n + 2) ** 2 + mu_eff))
    ds = 1 + 2 * max(0.0, math.sqrt((mu_eff - 1) / (n + 1)) - 1) + cs
    ps = math.sqrt(cs * (2 - cs) * mu_eff) * yw                       # ps_0 = 0, C^-1/2 = I
    assert np.allclose(st.ps, ps, rtol=1e-14)
    chiN = math.sqrt(3) * (1 - 1 / 12 + 1 / 189)
    norm_ps = float(np.linalg.norm(ps))
    hsig = norm_ps / math.sqrt(1 - (1 - cs) ** 2) / chiN < 1.4 + 2 / 4
    assert hsig                                                       # 0.94 / 0.758 / 1.596 = 0.78 < 1.9
    pc = math.sqrt(cc * (2 - cc) * mu_eff) * yw
    assert np.allclose(st.pc, pc, rtol=1e-14)
    dC = w1 * np.outer(z[0], z[0]) + w2 * np.outer(z[1], z[1])        # y = z while C = I
    C1 = (1 - c1 - cmu) * np.eye(3) + c1 * np.outer(pc, pc) + cmu * dC
    assert np.allclose(st.dC, dC, rtol=1e-14) and np.allclose(st.C, C1, rtol=1e-13)
    assert abs(C1[0, 2] - (c1 * pc[0] * pc[2] + cmu * (-w1))) < 1e-15    # one entry spelled out: only member 0 has z0*z2 != 0
    assert abs(st.sigma - 0.5 * math.exp((cs / ds) * (norm_ps / chiN - 1))) < 1e-15
    # the eigen-system the next ask() samples from reproduces C
    assert np.allclose((st.B * st.D ** 2) @ st.B.T, st.C, atol=1e-14)


def test_product_constants_equal_oracle_constants():
    from distributedes_b200.cma_es import cma_constants
    for n, lam in [(3, 4), (10, 10), (1024, 256), (4096, 1024)]:
        a, b = cma_constants(n, lam), co.cma_constants(n, lam)
        for key in ('mu', 'mu_eff', 'cc', 'c1', 'cmu', 'cs', 'ds'):
            assert a[key] == b[key], (n, lam, key)
        assert np.array_equal(a['w'], b['w'])


def test_lazy_eigendecomposition_gap():
    """tutorial B.2: B, D refreshed every 1/((c1+cmu) n 10) generations (>= 1).  Small lambda against n -> several
    generations between decompositions; the BASELINE configs (lambda ~ n/4) -> every generation."""
    k = co.cma_constants(1024, 256)
    assert co.eigen_gap(1024, k['c1'], k['cmu']) == 1
    k = co.cma_constants(4096, 1024)
    assert co.eigen_gap(4096, k['c1'], k['cmu']) == 1
    k = co.cma_constants(1000, 12)                                    # pycma's default lambda for n = 1000 is 4+floor(3 ln n) = 24
    gap = co.eigen_gap(1000, k['c1'], k['cmu'])
    assert gap == int(1.0 / ((k['c1'] + k['cmu']) * 1000 * 10)) and gap >= 2
    st = co.CMAState(np.zeros(40), 1.0, 6)
    assert st.gap >= 1
    rs = np.random.RandomState(1)
    B0 = st.B.copy()
    for g in range(st.gap):
        X = st.ask(rs.randn(6, 40))
        st.tell(X, co.sphere(X))
        if g + 1 < st.gap:
            assert np.array_equal(st.B, B0)                           # untouched between refreshes
    assert not np.array_equal(st.B, B0)
