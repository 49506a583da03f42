"""spsample (SURVEY section 8f rank 1, R/spsample_beta.R:19-113): oracle identities on CPU, GPU parity vs the oracle."""
import numpy as np
import pytest

import oracle
from oracle.spsample import spsample as oracle_spsample


def _inputs(n, S, cov_type, phir, p=2, seed=0):
    rng = np.random.Generator(np.random.Philox(seed))
    coords = rng.random((n, 2))
    X = rng.standard_normal((n, p))
    y = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1])) + X @ np.arange(1, p + 1) \
        + rng.standard_normal(n)
    phi = rng.uniform(*phir, S)
    sig2 = rng.uniform(0.5, 3.0, S)
    tau2 = rng.uniform(0.5, 2.0, S)
    return coords, X, y, phi, sig2, tau2, rng.standard_normal((S, p)), rng.standard_normal((S, n))


def test_oracle_conditional_mean_identity():
    """eps = 0: z = mu.Z = sig2 R (sig2 R + tau2 I)^-1 (y - X beta), the closed form of (R^-1/sig2 + I/tau2)^-1 w/tau2."""
    coords, X, y, phi, sig2, tau2, eb, ez = _inputs(40, 3, "matern2", (3.0, 8.0))
    out = oracle_spsample(y, X, coords, phi, sig2, tau2, "matern2", eb, np.zeros_like(ez))
    D = oracle.dist_matrix(coords)
    for s in range(3):
        R = oracle.cov_matern2(D, phi[s], 1.0)
        want = sig2[s] * R @ np.linalg.solve(sig2[s] * R + tau2[s] * np.eye(40), y - X @ out["beta"][s])
        np.testing.assert_allclose(out["z"][s], want, rtol=1e-8, atol=1e-9)
        assert out["beta0"][s] == pytest.approx(out["z"][s].mean())
    assert not out["bad"].any()


def test_oracle_jitter_fallback_and_chunks():
    coords, X, y, phi, sig2, tau2, eb, ez = _inputs(80, 4, "gaussian", (0.5, 1.0))     # singular R.Z -> + 1e-4 I
    a = oracle_spsample(y, X, coords, phi, sig2, tau2, "gaussian", eb, ez)
    b = oracle_spsample(y, X, coords, phi, sig2, tau2, "gaussian", eb, ez, chunks=2)
    assert np.isfinite(a["z"]).all()
    np.testing.assert_array_equal(a["z"], b["z"])          # chunking only matters for bad iterations


@pytest.mark.gpu
@pytest.mark.parametrize("n,S,cov_type,phir,p", [(60, 5, "matern2", (4.0, 10.0), 1), (155, 4, "matern1", (2.0, 6.0), 2),
                                                 (300, 3, "matern2", (10.0, 25.0), 3), (130, 3, "gaussian", (80.0, 160.0), 2)])
def test_gpu_spsample_vs_oracle(n, S, cov_type, phir, p):
    import spwombling_b200 as spw
    coords, X, y, phi, sig2, tau2, eb, ez = _inputs(n, S, cov_type, phir, p, seed=n)
    ref = oracle_spsample(y, X, coords, phi, sig2, tau2, cov_type, eb, ez)
    got = spw.spsample(y=y, X=X, coords=coords, phi=phi, sig2=sig2, tau2=tau2, cov_type=cov_type, eps_beta=eb, eps_z=ez)
    assert list(got) == ["z", "beta0", "beta"]
    np.testing.assert_allclose(got["beta"], ref["beta"], rtol=1e-10, atol=1e-12)
    scale = np.abs(ref["z"]).max(axis=1, keepdims=True)
    assert np.max(np.abs(got["z"] - ref["z"]) / scale) < 1e-9        # two explicit inverses: cond(K)-limited
    np.testing.assert_allclose(got["beta0"], ref["beta0"], rtol=1e-8, atol=1e-9)


@pytest.mark.gpu
def test_gpu_spsample_jitter_path():
    """chol(R.Z) fails for a numerically singular gaussian kernel; both sides retry with R.Z + 1e-4 I (:70-71)."""
    import spwombling_b200 as spw
    coords, X, y, phi, sig2, tau2, eb, ez = _inputs(100, 3, "gaussian", (0.5, 1.0), 1, seed=5)
    ref = oracle_spsample(y, X, coords, phi, sig2, tau2, "gaussian", eb, ez)
    got = spw.spsample(y=y, X=X, coords=coords, phi=phi, sig2=sig2, tau2=tau2, cov_type="gaussian", eps_beta=eb, eps_z=ez)
    scale = np.abs(ref["z"]).max(axis=1, keepdims=True)
    assert np.max(np.abs(got["z"] - ref["z"]) / scale) < 1e-6        # cond(R + 1e-4 I) ~ 1e6


@pytest.mark.gpu
def test_gpu_spsample_feeds_spatial_gradient():
    """The chain hlmBayes_sp -> spsample -> spatial_gradient of SURVEY section 1 runs end to end on the device path."""
    import spwombling_b200 as spw
    rng = np.random.Generator(np.random.Philox(9))
    n = 100
    coords = rng.random((n, 2))
    y = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1])) + rng.standard_normal(n)
    yc = y - y.mean()
    mc = spw.hlmBayes_sp(y=yc, coords=coords, niter=60, nburn=30, report=10, cov_type="matern2", verbose=False, rng=rng)
    zs = spw.spsample(y=yc, X=np.ones((n, 1)), coords=coords, phi=mc["phi"], sig2=mc["sig2"], tau2=mc["tau2"],
                      cov_type="matern2", rng=rng)
    gs = np.linspace(0.1, 0.9, 4)
    grid = np.array([[a, b] for b in gs for a in gs])
    out = spw.spatial_gradient(coords, {"phi": mc["phi"], "sig2": mc["sig2"], "z": zs["z"]}, "matern2", grid, nbatch=5,
                               rng=rng)
    assert out["grad1.est"].shape == (16, 3) and np.isfinite(out["grad.s11.mcmc"]).all()
