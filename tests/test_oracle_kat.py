"""Pin the CPU oracle: 50-digit known-answer vectors, an independent mpmath
re-derivation (oracle/arbiter.py), finite-difference identities and zero-lag limits.

The reference has no tests or golden vectors of its own, so these are the pins.
"""
import numpy as np
import pytest

import oracle
from oracle import arbiter
from oracle.gradient import cross_cov_rows, gradient_draws, ncomp_of, v0_matrix
from oracle.hlm import hlm_bayes_sp, r_chol, r_chol2inv

from conftest import (HLM_KAT, KAT, KAT_COORDS, KAT_EPS, KAT_GRID, KAT_PHI, KAT_SIG2, KAT_Z)

TYPES = ("gaussian", "matern1", "matern2")


@pytest.mark.parametrize("cov_type", TYPES)
def test_cov_kat(cov_type):
    D = oracle.dist_matrix(KAT_COORDS)
    K = oracle.cov_by_name(cov_type)(D, KAT_PHI, 1.0)
    want = KAT[cov_type]["K"]
    got = (K[0, 1], K[0, 2], K[2, 3])
    np.testing.assert_allclose(got, want, rtol=1e-14)
    assert np.all(np.diag(K) == 1.0)            # exact-zero diagonal of dist() => K_ii = sig2 exactly (Q17)
    assert np.array_equal(K, K.T)


@pytest.mark.parametrize("cov_type", TYPES)
def test_gradient_kat(cov_type):
    c = ncomp_of(cov_type)
    draws, mean, var = gradient_draws(KAT_COORDS, KAT_GRID, cov_type, [KAT_PHI], [KAT_SIG2], KAT_Z[None, :],
                                      KAT_EPS[None, None, :c], want_moments=True)
    k = KAT[cov_type]
    delta = KAT_COORDS - KAT_GRID[0]
    d = np.sqrt((delta ** 2).sum(1))
    _, At = cross_cov_rows(cov_type, KAT_PHI, d, delta)
    np.testing.assert_allclose(At[:, 0], k["row1"], rtol=1e-13)
    np.testing.assert_allclose(mean[0, 0], k["mean"], rtol=1e-12)
    iu = np.triu_indices(c)
    np.testing.assert_allclose(var[0, 0][iu], k["var_upper"], rtol=1e-11)
    np.testing.assert_allclose(draws[0, 0], k["draw"], rtol=1e-12)
    if c == 5:   # asymmetry of Q2: lower [3,1] entry is +2.53.. where the upper [1,3] is -2.53..
        assert var[0, 0][2, 0] == pytest.approx(-var[0, 0][0, 2], rel=1e-11)


@pytest.mark.parametrize("cov_type", TYPES)
def test_arbiter_rederives_kat(cov_type):
    c = ncomp_of(cov_type)
    a = arbiter.gradient_point(KAT_COORDS, cov_type, KAT_PHI, KAT_SIG2, KAT_Z, KAT_GRID[0], KAT_EPS[:c])
    k = KAT[cov_type]
    np.testing.assert_allclose(a["rows"][:, 0], k["row1"], rtol=5e-14)
    np.testing.assert_allclose(a["mean"], k["mean"], rtol=5e-14)
    np.testing.assert_allclose(a["var"][np.triu_indices(c)], k["var_upper"], rtol=5e-14)
    np.testing.assert_allclose(a["draw"], k["draw"], rtol=5e-14)


@pytest.mark.parametrize("cov_type", TYPES)
def test_rows_are_kernel_derivatives(cov_type):
    """nabla.K.t rows = [-d1 K, -d2 K, d11 K, d12 K, d22 K] w.r.t. the data coordinate (finite differences)."""
    import mpmath as mp
    mp.mp.dps = 50
    phi = mp.mpf("2.25")
    x0 = (mp.mpf("0.31"), mp.mpf("-0.17"))

    def Kf(d1, d2):
        return arbiter._cov(cov_type, phi, mp.sqrt(d1 * d1 + d2 * d2))

    h = mp.mpf("1e-12")
    d1, d2 = x0
    g1 = (Kf(d1 + h, d2) - Kf(d1 - h, d2)) / (2 * h)
    g2 = (Kf(d1, d2 + h) - Kf(d1, d2 - h)) / (2 * h)
    h11 = (Kf(d1 + h, d2) - 2 * Kf(d1, d2) + Kf(d1 - h, d2)) / (h * h)
    h22 = (Kf(d1, d2 + h) - 2 * Kf(d1, d2) + Kf(d1, d2 - h)) / (h * h)
    h12 = (Kf(d1 + h, d2 + h) - Kf(d1 + h, d2 - h) - Kf(d1 - h, d2 + h) + Kf(d1 - h, d2 - h)) / (4 * h * h)
    want = [-g1, -g2, h11, h12, h22]
    delta = np.array([[float(d1), float(d2)]])
    d = np.sqrt((delta ** 2).sum(1))
    _, At = cross_cov_rows(cov_type, float(phi), d, delta)
    c = ncomp_of(cov_type)
    np.testing.assert_allclose(At[:, 0], [float(v) for v in want[:c]], rtol=1e-12)


def test_v0_constants():
    phi = 1.75
    # matern1 / matern2 V.0 are the zero-lag limits; gaussian deliberately is not (Q4): 2 phi^2 vs 2 phi
    np.testing.assert_allclose(np.diag(v0_matrix("matern1", phi)), [3 * phi ** 2] * 2)
    V = v0_matrix("matern2", phi)
    np.testing.assert_allclose(np.diag(V), [5 / 3 * phi ** 2] * 2 + [25 * phi ** 4, 25 / 3 * phi ** 4, 25 * phi ** 4])
    assert V[2, 4] == V[4, 2] == pytest.approx(25 / 3 * phi ** 4)
    Vg = v0_matrix("gaussian", phi)
    np.testing.assert_allclose(np.diag(Vg), [2 * phi ** 2] * 2 + [12 * phi ** 4, 4 * phi ** 4, 12 * phi ** 4])


def test_chol_helpers_follow_lapack_upper():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((6, 6))
    S = A @ A.T + 6 * np.eye(6)
    S_bad_lower = S.copy()
    S_bad_lower[np.tril_indices(6, -1)] = 99.0      # chol must ignore the lower triangle
    U = r_chol(S_bad_lower)
    np.testing.assert_allclose(U.T @ U, S, rtol=1e-13)
    np.testing.assert_allclose(r_chol2inv(U), np.linalg.inv(S), rtol=1e-12)
    with pytest.raises(np.linalg.LinAlgError):
        r_chol(-np.eye(3))


def test_hlm_kat():
    y = HLM_KAT["y"]
    niter = 1
    norm = np.array([HLM_KAT["normal"], 0.0, 0.0])
    unif = np.array([0.5, 0.5, 0.5])
    out = hlm_bayes_sp(y, KAT_COORDS, niter=niter, nburn=0, report=1, cov_type="matern2",
                       norm=norm, unif=unif, trgt_fn_compute=True, keep_trace=True)
    assert out["trace_r"][0, 0] == pytest.approx(HLM_KAT["r_sig2"], rel=1e-12)
    assert out["trgt_fn"][0] == pytest.approx(HLM_KAT["trgt_init"], rel=1e-13)
    quad, logdiag = arbiter.hlm_quantities(KAT_COORDS, y, "matern2", 1.0, 1.0, 1.0)
    assert quad == pytest.approx(HLM_KAT["quad0"], rel=1e-14)
    assert logdiag == pytest.approx(HLM_KAT["logdiag0"], rel=1e-14)


def test_hlm_rng_order_and_bounds():
    """The phi uniform is consumed when the proposal is out of bounds; state is unchanged (Q10, a9)."""
    rng = np.random.default_rng(5)
    coords = rng.random((12, 2))
    y = rng.standard_normal(12)
    niter = 6
    norm = rng.standard_normal(3 * niter)
    unif = rng.random(3 * niter)
    norm[2] = 1e3        # iteration 1 phi proposal: exp(log 1 + 0.1e3) way above upper.phi
    out = hlm_bayes_sp(y, coords, niter=niter, nburn=2, report=3, cov_type="gaussian", norm=norm, unif=unif,
                       keep_trace=True)
    assert out["chain"]["phi"][0] == 1.0 and np.isnan(out["trace_r"][0, 2])
    assert out["sig2"].shape == (4,) and out["accept_m"].shape == (2, 3)
    assert out["nu"] == np.inf


def test_hlm_adaptation():
    rng = np.random.default_rng(6)
    coords = rng.random((15, 2))
    y = rng.standard_normal(15)
    niter = 40
    out = hlm_bayes_sp(y, coords, niter=niter, nburn=20, report=10, cov_type="matern1",
                       norm=rng.standard_normal(3 * niter), unif=rng.random(3 * niter))
    acc, tun = out["accept_m"], out["tuning"]
    assert np.allclose(tun[0], 0.1)
    for b in range(3):
        for p in range(3):
            x = max(0.17, min(acc[b, p], 0.75))
            f = x / 0.5 if x > 0.5 else (x / 0.25 if x < 0.25 else 1.0)
            assert tun[b + 1, p] == pytest.approx(tun[b, p] * f, rel=1e-15)
