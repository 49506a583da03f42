"""bayes_cwomb(type = "rectilinear") (SURVEY section 8f rank 3, R/spatial_wombling.R:159-244, :308-393, :464-711):
oracle identities on CPU, GPU parity against the oracle through the C ABI."""
import numpy as np
import pytest

from oracle.wombling import (_gamma_kernels, _k_terms, _u_perp, bayes_cwomb_rectilinear, mvrnorm2,
                             rectilinear_moments)


def _inputs(n, S, cov_type, seed=0):
    rng = np.random.Generator(np.random.Philox(seed))
    coords = rng.random((n, 2))
    curve = np.array([[0.2, 0.2], [0.35, 0.4], [0.5, 0.45], [0.7, 0.7], [0.8, 0.55]])
    # ranges that keep cond(K) small enough for a 1e-10 comparison and sig2 * var.womb positive definite
    phir = ((40.0, 80.0) if n < 100 else (250.0, 400.0)) if cov_type == "gaussian" else (8.0, 14.0)
    f = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1]))
    model = {"phi": rng.uniform(*phir, S), "sig2": rng.uniform(1.5, 2.5, S),    # sig2 well below ~1.2: var.womb not PD
             "z": (f - f.mean())[None, :] + 0.3 * rng.standard_normal((S, n))}
    eps = rng.standard_normal((S, curve.shape[0] - 1, 2))
    return coords, curve, model, eps


@pytest.mark.parametrize("cov_type,phi", [("gaussian", 30.0), ("matern2", 6.0)])
def test_gamma_kernels_are_directional_derivatives_of_the_covariance(cov_type, phi):
    """Gamma.t.1 = d/dn K(delta) along u_perp and Gamma.t.2 = d^2/dn^2 K(delta) (finite differences of cov_*)."""
    import oracle
    cov = oracle.cov_by_name(cov_type)
    u = np.array([0.6, 0.8])
    up = _u_perp(u)
    delta = np.array([0.13, -0.21])
    g1, g2 = _gamma_kernels(cov_type, delta, up, phi)
    K = lambda d: float(cov(np.array([np.sqrt((d ** 2).sum())]), phi, 1.0)[0])     # noqa: E731
    h = 1e-4
    fd1 = (K(delta + h * up) - K(delta - h * up)) / (2 * h)
    fd2 = (K(delta + h * up) - 2 * K(delta) + K(delta - h * up)) / h ** 2
    assert g1 == pytest.approx(fd1, rel=1e-6)
    assert g2 == pytest.approx(fd2, rel=1e-5)


def test_k_terms_zero_lag_branches_and_mvrnorm():
    u = np.array([0.6, 0.8])
    up = _u_perp(u)
    for cov_type, phi in (("gaussian", 30.0), ("matern2", 6.0)):
        K11, K12, K22 = _k_terms(cov_type, 0.0, u, up, phi)
        assert K12 == 0.0 and K11 < 0.0 and K22 > 0.0                 # R/spatial_wombling.R:532,541 / :690,706
    Sigma = np.array([[4.0, 99.0], [1.0, 3.0]])                       # mvrnorm reads the LOWER triangle
    x = mvrnorm2([1.0, 2.0], Sigma, [0.0, 0.0])
    assert np.array_equal(x, [1.0, 2.0])
    ev, V = np.linalg.eigh(np.array([[4.0, 1.0], [1.0, 3.0]]))
    x = mvrnorm2([0.0, 0.0], Sigma, [1.0, 0.0])
    assert np.linalg.norm(x) == pytest.approx(np.sqrt(ev[1]))
    with pytest.raises(np.linalg.LinAlgError):
        mvrnorm2([0.0, 0.0], np.array([[1.0, 0.0], [0.0, -1.0]]), [0.0, 0.0])


def test_oracle_return_structure():
    coords, curve, model, eps = _inputs(30, 6, "matern2")
    out = bayes_cwomb_rectilinear(coords, model, "matern2", 3, curve, eps)
    assert list(out) == ["tval", "uval", "womb1.mcmc", "womb2.mcmc", "womb.grad.inf", "womb.curv.inf", "womb.measure.inf"]
    assert out["womb1.mcmc"].shape == (4, 6) and out["womb.grad.inf"].shape == (4, 4) and out["womb.measure.inf"].shape == (2, 4)


@pytest.mark.gpu
@pytest.mark.parametrize("n,S,cov_type", [(40, 6, "matern2"), (40, 6, "gaussian"), (300, 4, "matern2"), (155, 5, "gaussian")])
def test_gpu_rectilinear_vs_oracle(n, S, cov_type):
    import spwombling_b200 as spw
    coords, curve, model, eps = _inputs(n, S, cov_type, seed=n)
    ref = bayes_cwomb_rectilinear(coords, model, cov_type, 2, curve, eps)
    got = spw.bayes_cwomb(coords=coords, model=model, cov_type=cov_type, nbatch=2, curve=curve, type="rectilinear",
                          eps=eps, return_moments=True)
    assert list(got)[:7] == list(ref)
    for s in range(S):
        mean, var, _, _ = rectilinear_moments(coords, curve, cov_type, model["phi"][s], model["sig2"][s], model["z"][s])
        assert np.max(np.abs(got["mean.womb"][s] - mean) / np.max(np.abs(mean), axis=1, keepdims=True)) < 1e-10
        assert np.max(np.abs(got["var.womb"][s] - var) / np.max(np.abs(var), axis=(1, 2), keepdims=True)) < 1e-10
    for key in ("womb1.mcmc", "womb2.mcmc"):
        assert np.max(np.abs(got[key] - ref[key])) <= 1e-9 * np.max(np.abs(ref[key]))
    for key in ("womb.grad.inf", "womb.curv.inf", "womb.measure.inf"):
        np.testing.assert_allclose(got[key], ref[key], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(got["tval"], ref["tval"], rtol=1e-15)


@pytest.mark.gpu
def test_gpu_rectilinear_error_behaviour():
    """mvrnorm stops when sig2 * var.womb is not positive definite (sig2 well below 1 with a smooth kernel); matern1 is
    refused because the reference's K.11.m1 is NaN on the diagonal of the double quadrature."""
    import spwombling_b200 as spw
    from spwombling_b200._lib import NotPositiveDefinite
    coords, curve, model, eps = _inputs(40, 3, "matern2", seed=3)
    model = dict(model, phi=np.full(3, 3.0), sig2=np.full(3, 0.5))
    with pytest.raises(np.linalg.LinAlgError):
        bayes_cwomb_rectilinear(coords, model, "matern2", 3, curve, eps)
    with pytest.raises(NotPositiveDefinite):
        spw.bayes_cwomb(coords=coords, model=model, cov_type="matern2", nbatch=3, curve=curve, type="rectilinear", eps=eps)
    with pytest.raises(NotImplementedError):
        spw.bayes_cwomb(coords=coords, model=model, cov_type="matern1", nbatch=3, curve=curve, type="rectilinear", eps=eps)
