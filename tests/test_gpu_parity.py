"""GPU parity tests: the CUDA path (through the C ABI, host buffers) against the CPU oracle on the same
seeded inputs.  Tolerances (stated per BASELINE.json north_star): deterministic quantities agree within
1e-10 relative in fp64, normalised by the largest magnitude of the block they belong to (entries of
V.0 - M cancel); phi ranges keep cond(K) <~ 1e6 where that bound is meaningful.
"""
import numpy as np
import pytest

import oracle
from oracle import arbiter
from oracle.gradient import cross_cov_rows, gradient_draws, ncomp_of
from oracle.hlm import hlm_bayes_sp, r_chol

from conftest import HLM_KAT, KAT, KAT_COORDS, KAT_EPS, KAT_GRID, KAT_PHI, KAT_SIG2, KAT_Z

pytestmark = pytest.mark.gpu
TYPES = ("gaussian", "matern1", "matern2")
RTOL = 1e-10


@pytest.fixture(scope="module")
def spw():
    import spwombling_b200 as m
    m.load()
    return m


def rel_err(got, want, axis=None):
    """max |got - want| / max |want| over the trailing block axes."""
    got, want = np.asarray(got), np.asarray(want)
    if axis is None:
        return np.max(np.abs(got - want)) / max(np.max(np.abs(want)), 1e-300)
    num = np.max(np.abs(got - want), axis=axis)
    den = np.maximum(np.max(np.abs(want), axis=axis), 1e-300)
    return np.max(num / den)


def synth(n, g_side, S, cov_type, phi_range, seed=0):
    rng = np.random.Generator(np.random.Philox(seed))
    coords = rng.random((n, 2))
    gs = np.linspace(0.05, 0.95, g_side)
    grid = np.array([[a, b] for b in gs for a in gs])
    phi = rng.uniform(*phi_range, S)
    sig2 = rng.uniform(0.5, 2.0, S)
    f = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1]))
    z = (f - f.mean())[None, :] + 0.3 * rng.standard_normal((S, n))
    eps = rng.standard_normal((S, grid.shape[0], ncomp_of(cov_type)))
    return coords, grid, phi, sig2, z, eps


# ---- covariance kernels ---------------------------------------------------------------------------------
@pytest.mark.parametrize("cov_type", TYPES)
@pytest.mark.parametrize("n", [1, 4, 63, 100, 155, 506])
def test_cov_matrix(spw, cov_type, n):
    rng = np.random.default_rng(n)
    coords = KAT_COORDS if n == 4 else rng.random((n, 2))
    phi, sig2 = (KAT_PHI, 1.7) if n == 4 else (rng.uniform(0.5, 20), rng.uniform(0.5, 2))
    K = spw.cov_matrix(coords, cov_type, phi, sig2)
    want = oracle.cov_by_name(cov_type)(oracle.dist_matrix(coords), phi, sig2)
    np.testing.assert_allclose(K, want, rtol=1e-13, atol=1e-300)
    assert np.all(np.diag(K) == sig2)                       # exact diagonal
    assert np.array_equal(K, K.T)
    if n == 4:
        np.testing.assert_allclose([K[0, 1], K[0, 2], K[2, 3]], np.array(KAT[cov_type]["K"]) * 1.7, rtol=1e-13)


@pytest.mark.parametrize("cov_type", TYPES)
def test_cov_delta_signature(spw, cov_type):
    rng = np.random.default_rng(3)
    D = rng.random((37, 11)) * 3
    D[0, 0] = 0.0
    fn = {"gaussian": spw.cov_gaussian, "matern1": spw.cov_matern1, "matern2": spw.cov_matern2}[cov_type]
    got = fn(Delta=D, phi=2.5, sig2=1.3)
    np.testing.assert_allclose(got, oracle.cov_by_name(cov_type)(D, 2.5, 1.3), rtol=1e-14)
    assert got[0, 0] == 1.3


@pytest.mark.parametrize("cov_type", TYPES)
@pytest.mark.parametrize("n,g", [(4, 1), (101, 7), (300, 40)])
def test_cross_cov(spw, cov_type, n, g):
    rng = np.random.default_rng(n + g)
    coords = KAT_COORDS if n == 4 else rng.random((n, 2))
    grid = KAT_GRID if g == 1 else rng.random((g, 2))
    phi = KAT_PHI if n == 4 else 7.5
    got = spw.cross_cov(coords, grid, cov_type, phi)
    for j in range(g):
        delta = coords - grid[j]
        d = np.sqrt((delta ** 2).sum(1))
        _, At = cross_cov_rows(cov_type, phi, d, delta)
        np.testing.assert_allclose(got[j], At, rtol=1e-12, atol=1e-14 * np.abs(At).max())
    if n == 4:
        np.testing.assert_allclose(got[0][:, 0], KAT[cov_type]["row1"], rtol=1e-13)


# ---- Cholesky -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 5, 63, 64, 65, 100, 128, 129, 155, 200, 506, 1000])
def test_chol_matches_lapack(spw, n):
    rng = np.random.default_rng(n)
    coords = rng.random((n, 2))
    A = oracle.cov_matern2(oracle.dist_matrix(coords), 6.0, 1.3) + 0.7 * np.eye(n)
    A_junk_lower = A.copy()
    A_junk_lower[np.tril_indices(n, -1)] = -123.0           # chol() must read the upper triangle only
    U, half_logdet = spw.chol(A_junk_lower)
    want = r_chol(A)
    assert rel_err(U, want) < 1e-12
    assert np.all(U[np.tril_indices(n, -1)] == 0.0)
    assert half_logdet == pytest.approx(np.sum(np.log(np.diag(want))), rel=1e-12, abs=1e-12)


def test_chol_not_positive_definite(spw):
    A = np.eye(70)
    A[66, 66] = -1.0
    with pytest.raises(spw.NotPositiveDefinite) as ei:
        spw.chol(A)
    assert ei.value.info[1] == 67                            # LAPACK: leading minor of order 67


# ---- spatial_gradient -----------------------------------------------------------------------------------
@pytest.mark.parametrize("cov_type", TYPES)
def test_gradient_kat(spw, cov_type):
    c = ncomp_of(cov_type)
    # nbatch = 1 has no HPD (coda needs nsamp > 1): use two identical samples and two batches instead
    model = {"phi": [KAT_PHI, KAT_PHI], "sig2": [KAT_SIG2, KAT_SIG2], "z": np.stack([KAT_Z, KAT_Z])}
    eps = np.stack([KAT_EPS[None, :c], KAT_EPS[None, :c]])
    out = spw.spatial_gradient(KAT_COORDS, model, cov_type, KAT_GRID, nbatch=2, eps=eps, return_moments=True,
                               rounded=False)
    k = KAT[cov_type]
    names = ("s1", "s2", "s11", "s12", "s22")[:c]
    draw = np.array([out["grad.%s.mcmc" % nm][0, 0] for nm in names])
    np.testing.assert_allclose(out["mean.grad"][0, 0], k["mean"], rtol=1e-11)
    np.testing.assert_allclose(out["var.grad"][0, 0][np.triu_indices(c)], k["var_upper"], rtol=1e-10)
    np.testing.assert_allclose(draw, k["draw"], rtol=1e-11)
    np.testing.assert_allclose(out["mean.grad"][1], out["mean.grad"][0], rtol=0, atol=0)


CASES = [
    # n, grid side, S, type, phi range  (ranges from SURVEY section 8d / 9.2: cond(K) <~ 1e6)
    (100, 5, 6, "matern2", (10, 30)),
    (100, 5, 6, "matern1", (3, 12)),
    (100, 5, 6, "gaussian", (60, 150)),
    (155, 6, 4, "matern1", (2, 8)),
    (257, 4, 3, "matern2", (15, 40)),
    (506, 7, 3, "matern2", (30, 60)),
    (506, 4, 2, "gaussian", (600, 1200)),
    (1100, 5, 2, "matern2", (60, 100)),
]


@pytest.mark.parametrize("n,gside,S,cov_type,phir", CASES)
def test_gradient_vs_oracle(spw, n, gside, S, cov_type, phir):
    coords, grid, phi, sig2, z, eps = synth(n, gside, S, cov_type, phir, seed=n + S)
    want_d, want_m, want_v = gradient_draws(coords, grid, cov_type, phi, sig2, z, eps, want_moments=True)
    out = spw.spatial_gradient(coords, {"phi": phi, "sig2": sig2, "z": z}, cov_type, grid, nbatch=2, eps=eps,
                               return_moments=True, rounded=False)
    c = ncomp_of(cov_type)
    names = ("s1", "s2", "s11", "s12", "s22")[:c]
    got_d = np.stack([out["grad.%s.mcmc" % nm] for nm in names], axis=2)       # S x G x c
    assert rel_err(out["mean.grad"], want_m, axis=2) < RTOL
    assert rel_err(out["var.grad"].reshape(S, -1, c * c), want_v.reshape(S, -1, c * c), axis=2) < RTOL
    assert rel_err(got_d, want_d, axis=2) < RTOL
    # the (non-symmetric) var.grad of the reference: gradient/curvature blocks carry opposite signs (Q2)
    if c == 5:
        v = out["var.grad"]
        np.testing.assert_allclose(v[..., 2, 0], -v[..., 0, 2], rtol=1e-9, atol=1e-9 * np.abs(v).max())


def test_gradient_ill_conditioned_uses_arbiter(spw):
    """cond(K) ~ 1e9: float64 algorithms disagree beyond 1e-10; accept the GPU result if it is at least as
    close to the 50-digit arbiter as the float64 oracle is (SURVEY section 8c two-tier oracle)."""
    rng = np.random.default_rng(11)
    n = 40
    coords = rng.random((n, 2))
    grid = np.array([[0.5, 0.5]])
    phi, sig2 = 1.5, 1.2
    z = rng.standard_normal((1, n))
    eps = rng.standard_normal((1, 1, 5))
    _, om, ov = gradient_draws(coords, grid, "matern2", [phi], [sig2], z, eps, want_moments=True)
    model = {"phi": [phi, phi], "sig2": [sig2, sig2], "z": np.vstack([z, z])}
    out = spw.spatial_gradient(coords, model, "matern2", grid, nbatch=2, eps=np.vstack([eps, eps]),
                               return_moments=True, rounded=False)
    arb = arbiter.gradient_point(coords, "matern2", phi, sig2, z[0], grid[0], eps[0, 0])
    err_gpu = rel_err(out["mean.grad"][0, 0], arb["mean"])
    err_cpu = rel_err(om[0, 0], arb["mean"])
    assert err_gpu <= max(10 * err_cpu, 1e-10)
    assert rel_err(out["var.grad"][0, 0], arb["var"]) <= max(10 * rel_err(ov[0, 0], arb["var"]), 1e-10)


def test_gradient_not_positive_definite(spw):
    """Gaussian kernel with small phi is numerically singular: the reference's chol() stops; so do we."""
    coords, grid, phi, sig2, z, eps = synth(200, 3, 2, "gaussian", (0.5, 1.0), seed=5)
    with pytest.raises(spw.NotPositiveDefinite):
        spw.spatial_gradient(coords, {"phi": phi, "sig2": sig2, "z": z}, "gaussian", grid, nbatch=2, eps=eps)


def test_gradient_sample_sharding_invariance(spw):
    """Results for a sample do not depend on which other samples share its batch / call (bit-identical)."""
    coords, grid, phi, sig2, z, eps = synth(130, 6, 7, "matern2", (10, 30), seed=9)
    full = spw.spatial_gradient(coords, {"phi": phi, "sig2": sig2, "z": z}, "matern2", grid, nbatch=2, eps=eps,
                                rounded=False)
    part = spw.spatial_gradient(coords, {"phi": phi[3:6], "sig2": sig2[3:6], "z": z[3:6]}, "matern2", grid,
                                nbatch=2, eps=eps[3:6], rounded=False)
    for nm in ("s1", "s2", "s11", "s12", "s22"):
        assert np.array_equal(full["grad.%s.mcmc" % nm][3:6], part["grad.%s.mcmc" % nm])


# ---- summary --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("S,g,nbatch", [(50, 7, 5), (200, 33, 100), (5000, 19, 200), (37, 3, 10), (100, 130, 100)])
def test_hpd_summary(spw, S, g, nbatch):
    rng = np.random.default_rng(S + g)
    draws = rng.standard_normal((S, g)) * rng.uniform(0.1, 10, g) + rng.uniform(-5, 5, g)
    draws[:, 0] = np.round(draws[:, 0], 1)                  # ties
    got = spw.hpd_summary(draws, nbatch)
    want = oracle.summarise_draws(draws, nbatch)
    np.testing.assert_allclose(got, want, rtol=0, atol=0)   # order statistics: bit-exact


def test_spatial_gradient_summary_and_names(spw):
    coords, grid, phi, sig2, z, eps = synth(100, 4, 40, "matern2", (10, 30), seed=21)
    model = {"phi": phi, "sig2": sig2, "z": z}
    out = spw.spatial_gradient(coords, model, "matern2", grid, nbatch=10, eps=eps)
    ref = oracle.spatial_gradient(coords, model, "matern2", grid, nbatch=10, eps=eps)
    assert list(out) == list(ref)
    for key in ("grad1.est", "grad2.est", "grad11.est", "grad12.est", "grad22.est"):
        assert out[key].shape == (16, 3)
        # rounded to 4 digits like the reference; allow one unit in the last place for boundary flips
        assert np.max(np.abs(out[key] - ref[key])) <= 1.0001e-4
    m1 = spw.spatial_gradient(coords, {"phi": phi / 3, "sig2": sig2, "z": z}, "matern1", grid, nbatch=10,
                              eps=eps[:, :, :2], return_mcmc=False)
    assert list(m1) == ["grad1.est", "grad2.est"]


# ---- hlmBayes_sp ------------------------------------------------------------------------------------------
def test_hlm_kat(spw):
    out = spw.hlmBayes_sp(y=HLM_KAT["y"], coords=KAT_COORDS, niter=1, nburn=0, report=1, cov_type="matern2",
                          norm=[HLM_KAT["normal"], 0.0, 0.0], unif=[0.5, 0.5, 0.5], trgtFn_compute=True,
                          verbose=False)
    assert out["trgt_fn"][0] == pytest.approx(HLM_KAT["trgt_init"], rel=1e-12)
    ref = hlm_bayes_sp(HLM_KAT["y"], KAT_COORDS, niter=1, nburn=0, report=1, cov_type="matern2",
                       norm=[HLM_KAT["normal"], 0.0, 0.0], unif=[0.5, 0.5, 0.5], trgt_fn_compute=True)
    np.testing.assert_allclose(out["trgt_fn"], ref["trgt_fn"], rtol=1e-12)
    np.testing.assert_allclose(out["sig2"], ref["sig2"], rtol=1e-13)


@pytest.mark.parametrize("n,cov_type,niter,report", [(30, "matern2", 60, 10), (100, "matern1", 120, 20),
                                                     (100, "gaussian", 90, 30), (155, "matern2", 40, 10),
                                                     (300, "matern2", 30, 10),
                                                     # single-CTA chain: degenerate sizes, the one / two owner-warp
                                                     # boundary (n = 124 | 125) and the largest system it takes
                                                     (1, "matern2", 12, 4), (2, "matern1", 12, 4), (5, "gaussian", 12, 4),
                                                     (124, "matern2", 18, 6), (125, "matern1", 18, 6),
                                                     (131, "gaussian", 18, 6), (200, "matern2", 15, 5)])
def test_hlm_chain_vs_oracle(spw, n, cov_type, niter, report):
    rng = np.random.Generator(np.random.Philox(n + niter))
    coords = rng.random((n, 2))
    y = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1])) + rng.standard_normal(n)
    norm = rng.standard_normal(3 * niter)
    unif = rng.random(3 * niter)
    norm[5] = 80.0                                          # one out-of-bounds phi proposal (a9)
    nburn = niter // 2
    ref = hlm_bayes_sp(y, coords, niter=niter, nburn=nburn, report=report, cov_type=cov_type, norm=norm, unif=unif,
                       trgt_fn_compute=True)
    out = spw.hlmBayes_sp(y=y, coords=coords, niter=niter, nburn=nburn, report=report, cov_type=cov_type,
                          norm=norm, unif=unif, trgtFn_compute=True, verbose=False)
    for key in ("sig2", "tau2", "phi"):
        np.testing.assert_allclose(out[key], ref[key], rtol=1e-10)
        np.testing.assert_allclose(out["diagnostics"]["chain"][key], ref["chain"][key], rtol=1e-10)
        assert out[key].shape == (niter - nburn,)
    np.testing.assert_allclose(out["trgt_fn"], ref["trgt_fn"], rtol=1e-10)
    np.testing.assert_allclose(out["diagnostics"]["accept_m"], ref["accept_m"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(out["diagnostics"]["tuning"], ref["tuning"], rtol=1e-13)
    assert out["nu"] == ref["nu"]
    assert list(out)[:2] == ["y", "coords"]


def test_hlm_defaults_and_graph_equivalence(spw, monkeypatch):
    """Graph replay and eager launches of the blocked path produce the same chain (bit-identical)."""
    monkeypatch.setenv("SPW_HLM_SMALL", "0")
    rng = np.random.default_rng(2)
    coords = rng.random((64, 2))
    y = rng.standard_normal(64)
    niter = 20
    kw = dict(y=y, coords=coords, niter=niter, nburn=10, report=5, cov_type="matern1",
              norm=rng.standard_normal(3 * niter), unif=rng.random(3 * niter), verbose=False)
    a = spw.hlmBayes_sp(**kw)
    monkeypatch.setenv("SPW_HLM_GRAPH", "0")
    b = spw.hlmBayes_sp(**kw)
    for key in ("sig2", "tau2", "phi"):
        assert np.array_equal(a[key], b[key])


# ---- workspace planning: batching over samples and chunking over the grid change no arithmetic ---------------
def test_workspace_chunking_bit_identical(spw):
    coords, grid, phi, sig2, z, eps = synth(300, 9, 21, "matern2", (15, 40), seed=33)      # G = 81, S = 21
    model = {"phi": phi, "sig2": sig2, "z": z}
    lib = spw.load()
    full = spw.spatial_gradient(coords, model, "matern2", grid, nbatch=3, eps=eps, return_moments=True, rounded=False)
    try:
        # ~6 MB: forces several sample batches AND several grid chunks (per sample: 2 x 300 x 304 x 8 B of factors
        # plus 5 x 304 x 8 B per grid point)
        lib.spw_set_workspace_limit(6 << 20)
        small = spw.spatial_gradient(coords, model, "matern2", grid, nbatch=3, eps=eps, return_moments=True,
                                     rounded=False)
    finally:
        lib.spw_set_workspace_limit(0)
    for k in full:
        assert np.array_equal(full[k], small[k]), k


def test_argument_errors(spw):
    coords, grid, phi, sig2, z, eps = synth(50, 3, 4, "matern2", (10, 30), seed=1)
    with pytest.raises(ValueError):
        spw.spatial_gradient(coords, {"phi": phi, "sig2": sig2[:2], "z": z}, "matern2", grid, nbatch=2, eps=eps)
    with pytest.raises(ValueError):
        spw.spatial_gradient(coords, {"phi": phi, "sig2": sig2, "z": z}, "matern", grid, nbatch=2, eps=eps)
    with pytest.raises(ValueError):        # coda::HPDinterval needs more than one batch median
        spw.spatial_gradient(coords, {"phi": phi, "sig2": sig2, "z": z}, "matern2", grid, nbatch=1, eps=eps)


# ---- posterior surface analysis (README.md:123-177) -------------------------------------------
def test_surface_analysis(spw):
    from oracle.surface import surface_analysis as oracle_surface
    rng = np.random.default_rng(4)
    S, G = 60, 23
    est = {"grad.%s.mcmc" % k: rng.standard_normal((S, G)) * sc for k, sc in
           zip(("s1", "s2", "s11", "s12", "s22"), (3, 3, 40, 25, 40))}
    got = spw.surface_analysis(est, nbatch=10, return_mcmc=True)
    ref = oracle_surface(*(est["grad.%s.mcmc" % k] for k in ("s1", "s2", "s11", "s12", "s22")), nbatch=10)
    for name in ("det", "eigen1", "eigen2", "laplace", "div"):
        scale = np.abs(ref[name + ".mcmc"]).max()
        assert np.max(np.abs(got[name + ".mcmc"] - ref[name + ".mcmc"])) < 1e-12 * max(scale, 1.0) * 50
        np.testing.assert_allclose(got[name][:, :3], ref[name][:, :3], rtol=1e-11, atol=1e-11 * scale)
        assert np.array_equal(got[name][:, 3], ref[name][:, 3])


# ---- bayes_cwomb(type = "riemann.sum") rides on spatial_gradient (R/spatial_wombling.R:114-157, :245-306) --------
@pytest.mark.parametrize("cov_type,phir", [("matern2", (10, 30)), ("matern1", (3, 10))])
def test_bayes_cwomb_riemann(spw, cov_type, phir):
    from oracle.wombling import bayes_cwomb as oracle_cwomb
    rng = np.random.Generator(np.random.Philox(12))
    n, S = 120, 40
    coords = rng.random((n, 2))
    t = np.linspace(0.15, 0.85, 9)
    curve = np.column_stack([t, 0.5 + 0.25 * np.sin(2 * np.pi * t)])
    c = ncomp_of(cov_type)
    model = {"phi": rng.uniform(*phir, S), "sig2": rng.uniform(0.5, 2, S), "z": rng.standard_normal((S, n))}
    eps = rng.standard_normal((S, curve.shape[0], c))
    got = spw.bayes_cwomb(coords=coords, model=model, cov_type=cov_type, nbatch=8, curve=curve, type="riemann.sum",
                          eps=eps, rounded=False)
    ref = oracle_cwomb(coords, model, cov_type, 8, curve, eps, rounded=False)
    assert list(got) == list(ref)
    for k in ref:
        scale = max(np.abs(ref[k]).max(), 1.0)
        assert np.max(np.abs(np.asarray(got[k]) - ref[k])) < 1e-9 * scale, k
    with pytest.raises(ValueError):
        spw.bayes_cwomb(coords=coords, model=model, cov_type=cov_type, nbatch=8, curve=curve, type="simpson")


def test_hlm_summary(spw):
    """R/hlm_summary.R:19-50: medians + coda::HPDinterval on the raw samples."""
    from oracle.summary import hpd_interval
    rng = np.random.default_rng(8)
    S, n = 400, 17
    sig2, tau2, phi = rng.gamma(2, 1, S), rng.gamma(3, 0.5, S), rng.uniform(1, 5, S)
    beta = rng.standard_normal((S, 2))
    z = rng.standard_normal((S, n)) + np.linspace(-1, 1, n)
    got = spw.hlm_summary(sig2=sig2, tau2=tau2, phi=phi, beta=beta, z=z)
    cols = [sig2, tau2, phi, beta[:, 0], beta[:, 1]]
    for r, x in enumerate(cols):
        lo, hi = hpd_interval(x[:, None])
        np.testing.assert_allclose(got["param"][r], [np.median(x), lo[0], hi[0]], rtol=0, atol=0)
    lo, hi = hpd_interval(z)
    np.testing.assert_allclose(got["z"][:, :3], np.column_stack([np.median(z, axis=0), lo, hi]), rtol=0, atol=0)
    assert set(np.unique(got["z"][:, 3])) <= {-1.0, 0.0, 1.0}


# ---- opt-in schedules give the same results as the default ones ---------------------------------------------------
def test_hlm_opt_in_paths(spw, monkeypatch):
    rng = np.random.Generator(np.random.Philox(77))
    n, niter = 90, 30
    coords = rng.random((n, 2))
    y = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1])) + rng.standard_normal(n)
    kw = dict(y=y, coords=coords, niter=niter, nburn=10, report=10, cov_type="matern2",
              norm=rng.standard_normal(3 * niter), unif=rng.random(3 * niter), trgtFn_compute=True, verbose=False)
    monkeypatch.setenv("SPW_HLM_SMALL", "0")          # blocked multi-kernel path (CUDA graph per iteration)
    base = spw.hlmBayes_sp(**kw)
    monkeypatch.setenv("SPW_HLM_SMALL", "1")          # single-CTA shared-memory chain (default for n <= 200)
    small = spw.hlmBayes_sp(**kw)
    monkeypatch.delenv("SPW_HLM_SMALL")
    for key in ("sig2", "tau2", "phi", "trgt_fn"):
        np.testing.assert_allclose(small[key], base[key], rtol=1e-10)
    np.testing.assert_array_equal(small["diagnostics"]["accept_m"], base["diagnostics"]["accept_m"])
    monkeypatch.setenv("SPW_LOOKAHEAD", "1")          # two-stream lookahead in the single-matrix Cholesky
    A = oracle.cov_matern2(oracle.dist_matrix(rng.random((900, 2))), 6.0, 1.3) + 0.7 * np.eye(900)
    U1, l1 = spw.chol(A)
    monkeypatch.delenv("SPW_LOOKAHEAD")
    U0, l0 = spw.chol(A)
    assert np.array_equal(U0, U1) and l0 == l1


# ---- edge shapes of the fused kernel: N a multiple of the 128-row block (no ragged block), tiny / ragged grids -----
@pytest.mark.parametrize("n,G,cov_type,phir", [(128, 1, "matern2", (10, 30)), (256, 17, "matern2", (15, 40)),
                                               (384, 33, "matern1", (20, 40)), (127, 16, "gaussian", (80, 160)),
                                               (129, 31, "matern2", (10, 30)), (3, 2, "matern2", (2, 4))])
def test_gradient_edge_shapes(spw, n, G, cov_type, phir):
    rng = np.random.Generator(np.random.Philox(n * 1000 + G))
    coords = rng.random((n, 2))
    grid = rng.random((G, 2))
    S = 3
    c = ncomp_of(cov_type)
    phi, sig2 = rng.uniform(*phir, S), rng.uniform(0.5, 2, S)
    z = rng.standard_normal((S, n))
    eps = rng.standard_normal((S, G, c))
    want_d, want_m, want_v = gradient_draws(coords, grid, cov_type, phi, sig2, z, eps, want_moments=True)
    out = spw.spatial_gradient(coords, {"phi": phi, "sig2": sig2, "z": z}, cov_type, grid, nbatch=3, eps=eps,
                               return_moments=True, rounded=False)
    names = ("s1", "s2", "s11", "s12", "s22")[:c]
    got_d = np.stack([out["grad.%s.mcmc" % nm] for nm in names], axis=2)
    assert rel_err(out["mean.grad"], want_m, axis=2) < RTOL
    assert rel_err(out["var.grad"].reshape(S, -1, c * c), want_v.reshape(S, -1, c * c), axis=2) < RTOL
    assert rel_err(got_d, want_d, axis=2) < RTOL


@pytest.mark.parametrize("S,G,nbatch", [(3000, 5, 3000), (4000, 37, 4), (100000, 3, 100000), (2600, 4, 1300), (70000, 2, 7)])
def test_summary_scales_by_sorting(S, G, nbatch):
    """hlm_summary / the wombling measure use one batch per sample (nb = S) and small nbatch means long batches: both
    leave the rank-counting kernels for the bitonic-sort path (in shared memory up to 16384 values per row, global
    passes above).  Bit-exact against the oracle's coda::HPDinterval restatement."""
    import spwombling_b200 as spw
    from oracle.summary import summarise_draws
    rng = np.random.Generator(np.random.Philox(S + G))
    draws = rng.standard_normal((S, G))
    draws[::7] = np.round(draws[::7], 1)                     # ties
    got = spw.hpd_summary(draws, nbatch)
    assert np.array_equal(got, summarise_draws(draws, nbatch))


def test_summary_propagates_nan():
    """A NaN draw makes the affected estimates NaN (R: median / HPDinterval return NA) instead of silently 0."""
    import spwombling_b200 as spw
    rng = np.random.Generator(np.random.Philox(5))
    for S, nbatch in ((60, 12), (3000, 3000)):
        draws = rng.standard_normal((S, 4))
        draws[S // 2, 1] = np.nan
        got = spw.hpd_summary(draws, nbatch)
        assert np.isnan(got[1]).all() and np.isfinite(got[[0, 2, 3]]).all()


@pytest.mark.parametrize("n", [60, 300])
def test_hlm_stops_at_a_non_finite_ratio(n):
    """A proposal whose log-ratio is NaN must not be accepted through fmin(NaN, 0) = 0: the reference's
    `if(log(runif(1)) < accept.prob)` raises an error at that iteration; here the chain rejects, records the iteration
    and the call fails (both the single-CTA and the blocked chain)."""
    import spwombling_b200 as spw
    from spwombling_b200._lib import SpwError
    rng = np.random.Generator(np.random.Philox(n))
    coords = rng.random((n, 2))
    y = rng.standard_normal(n)
    y[3] = np.nan
    with pytest.raises(SpwError) as e:
        spw.hlmBayes_sp(y=y, coords=coords, niter=4, nburn=0, report=2, cov_type="matern2", norm=rng.standard_normal(12),
                        unif=rng.random(12), verbose=False)
    assert e.value.info[0] == 1                      # the first iteration
