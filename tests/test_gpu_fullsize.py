"""BASELINE.json's full matrix orders (N = 2000 gaussian, N = 5000 matern2) through size-independent properties and
an independent fp64 evaluation on the same GPU (torch.linalg: cuSOLVER Cholesky + triangular solves -- a different
algorithm from both the reference's explicit inverse and this library's triangular inverse).  The CPU oracle would
need minutes per sample at these sizes, so it checks a single (sample, few grid points) slice only.
Tolerance: 1e-10 relative, normalised per 5-vector / 5x5 block (north_star), at phi ranges with cond(K) <~ 1e6."""
import numpy as np
import pytest

from oracle.gradient import cross_cov_rows, ncomp_of, v0_matrix

pytestmark = pytest.mark.gpu
SIGN = np.array([-1.0, -1.0, 1.0, 1.0, 1.0])


def _inputs(n, cov_type, phir, S, gside, seed):
    rng = np.random.Generator(np.random.Philox(seed))
    coords = rng.random((n, 2))
    gs = np.linspace(0.05, 0.95, gside)
    grid = np.array([[a, b] for b in gs for a in gs])
    phi = rng.uniform(*phir, S)
    sig2 = rng.uniform(0.5, 2.0, S)
    f = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1]))
    z = (f - f.mean())[None, :] + 0.3 * rng.standard_normal((S, n))
    eps = rng.standard_normal((S, grid.shape[0], ncomp_of(cov_type)))
    return coords, grid, phi, sig2, z, eps


def _torch_moments(coords, grid, cov_type, phi, z):
    """mean / var for one sample by Cholesky solves in torch fp64 on the GPU."""
    import torch
    import oracle
    c = ncomp_of(cov_type)
    dev = torch.device("cuda", 0)
    K = torch.from_numpy(oracle.cov_by_name(cov_type)(oracle.dist_matrix(coords), phi, 1.0)).to(dev)
    L = torch.linalg.cholesky(K)
    A = np.empty((grid.shape[0], c, coords.shape[0]))
    for g in range(grid.shape[0]):
        delta = coords - grid[g]
        _, A[g] = cross_cov_rows(cov_type, phi, np.sqrt((delta ** 2).sum(1)), delta)
    At = torch.from_numpy(A.reshape(-1, coords.shape[0]).T.copy()).to(dev)          # N x (G c)
    W = torch.linalg.solve_triangular(L, At, upper=False)
    v = torch.linalg.solve_triangular(L, torch.from_numpy(z).to(dev)[:, None], upper=False)
    mean = (W.T @ v).reshape(grid.shape[0], c).cpu().numpy()
    Wg = W.T.reshape(grid.shape[0], c, -1)
    M = torch.einsum("gak,gbk->gab", Wg, Wg).cpu().numpy()
    var = v0_matrix(cov_type, phi)[None] - M * SIGN[:c][None, None, :]
    return mean, var


def _blk_err(got, want):
    got, want = got.reshape(got.shape[0], -1), want.reshape(want.shape[0], -1)
    return np.max(np.max(np.abs(got - want), axis=1) / np.max(np.abs(want), axis=1))


@pytest.mark.parametrize("n,cov_type,phir", [(2000, "gaussian", (3000.0, 6000.0)), (5000, "matern2", (100.0, 200.0)),
                                              (2000, "matern1", (30.0, 60.0))])
def test_fullsize_against_independent_fp64(n, cov_type, phir):
    import spwombling_b200 as spw
    S = 2
    coords, grid, phi, sig2, z, eps = _inputs(n, cov_type, phir, S, 8, seed=n)
    c = ncomp_of(cov_type)
    out = spw.spatial_gradient(coords, {"phi": phi, "sig2": sig2, "z": z}, cov_type, grid, nbatch=2, eps=eps,
                               return_moments=True, rounded=False)
    for s in range(S):
        mean, var = _torch_moments(coords, grid, cov_type, phi[s], z[s])
        assert _blk_err(out["mean.grad"][s], mean) < 1e-10
        assert _blk_err(out["var.grad"][s], var) < 1e-10
    # properties that need no reference --------------------------------------------------------------------
    names = ("s1", "s2", "s11", "s12", "s22")[:c]
    draws = np.stack([out["grad.%s.mcmc" % nm] for nm in names], axis=2)
    # (i) eps = 0  =>  draw == mean exactly; (ii) var does not depend on z or sig2; (iii) mean is linear in z
    z2 = z[::-1].copy()
    out0 = spw.spatial_gradient(coords, {"phi": phi, "sig2": sig2 * 3.0, "z": z2}, cov_type, grid, nbatch=2,
                                eps=np.zeros_like(eps), return_moments=True, rounded=False)
    d0 = np.stack([out0["grad.%s.mcmc" % nm] for nm in names], axis=2)
    assert np.array_equal(d0, out0["mean.grad"])
    assert np.array_equal(out0["var.grad"], out["var.grad"])
    outs = spw.spatial_gradient(coords, {"phi": phi, "sig2": sig2, "z": z + z2}, cov_type, grid, nbatch=2,
                                eps=eps, return_moments=True, rounded=False)
    assert _blk_err(outs["mean.grad"].reshape(S, -1), (out["mean.grad"] + out0["mean.grad"]).reshape(S, -1)) < 1e-10
    # (iv) (draw - mean) scales with sqrt(sig2): U eps is sig2-free
    out3 = spw.spatial_gradient(coords, {"phi": phi, "sig2": sig2 * 4.0, "z": z}, cov_type, grid, nbatch=2, eps=eps,
                                return_moments=True, rounded=False)
    d3 = np.stack([out3["grad.%s.mcmc" % nm] for nm in names], axis=2)
    np.testing.assert_allclose(d3 - out3["mean.grad"], 2.0 * (draws - out["mean.grad"]), rtol=1e-9,
                               atol=1e-9 * np.abs(draws).max())
    # (v) one slice against the CPU oracle (explicit-inverse sandwich of the reference)
    if n <= 2000:
        from oracle.gradient import gradient_draws
        gd, gm, gv = gradient_draws(coords, grid[:3], cov_type, phi[:1], sig2[:1], z[:1], eps[:1, :3],
                                    want_moments=True)
        assert _blk_err(out["mean.grad"][0, :3], gm[0]) < 1e-10
        assert _blk_err(out["var.grad"][0, :3], gv[0]) < 1e-10
        assert _blk_err(draws[0, :3], gd[0]) < 1e-10


@pytest.mark.parametrize("n", [2000, 5000])
def test_fullsize_cholesky_and_hlm_quantities(n):
    """chol(): U'U reproduces the matrix and matches cuSOLVER's factor; hlm's first target value (the (sic) formula of
    R/hlmBayes_sp.R:103-105) agrees with quantities computed from an independent factorisation."""
    import torch
    import oracle
    import spwombling_b200 as spw
    rng = np.random.Generator(np.random.Philox(n + 1))
    coords = rng.random((n, 2))
    y = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1])) + rng.standard_normal(n)
    Sigma = oracle.cov_matern2(oracle.dist_matrix(coords), 1.0, 1.0) + np.eye(n)        # phi = sig2 = tau2 = 1
    U, half_logdet = spw.chol(Sigma)
    dev = torch.device("cuda", 0)
    Lt = torch.linalg.cholesky(torch.from_numpy(Sigma).to(dev))
    assert np.max(np.abs(U.T - Lt.cpu().numpy())) / np.max(np.abs(U)) < 1e-11
    want_half = float(torch.log(torch.diagonal(Lt)).sum())
    assert half_logdet == pytest.approx(want_half, rel=1e-12)
    w = torch.linalg.solve_triangular(Lt, torch.from_numpy(y).to(dev)[:, None], upper=False)
    quad = float((w * w).sum())
    out = spw.hlmBayes_sp(y=y, coords=coords, niter=1, nburn=0, report=1, cov_type="matern2", norm=np.zeros(3),
                          unif=np.full(3, 0.5), trgtFn_compute=True, verbose=False)
    want_trgt = -(2 + 1) * 0.0 - 1.0 - -(2 + 1) * 0.0 - 1.0 - want_half / 4 - quad / 2
    assert out["trgt_fn"][0] == pytest.approx(want_trgt, rel=1e-10)


def test_hlm_chain_blocked_path_vs_oracle():
    """A chain at a size that exercises the multi-block Cholesky schedule (16 diagonal blocks, 4 outer blocks, the
    appended y row) against the reference's explicit-inverse arithmetic."""
    import spwombling_b200 as spw
    from oracle.hlm import hlm_bayes_sp
    n, niter = 1000, 6
    rng = np.random.Generator(np.random.Philox(4242))
    coords = rng.random((n, 2))
    y = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1])) + rng.standard_normal(n)
    y = y - y.mean()
    norm, unif = rng.standard_normal(3 * niter), rng.random(3 * niter)
    ref = hlm_bayes_sp(y, coords, niter=niter, nburn=2, report=3, cov_type="matern1", norm=norm, unif=unif,
                       trgt_fn_compute=True)
    out = spw.hlmBayes_sp(y=y, coords=coords, niter=niter, nburn=2, report=3, cov_type="matern1", norm=norm, unif=unif,
                          trgtFn_compute=True, verbose=False)
    for key in ("sig2", "tau2", "phi"):
        np.testing.assert_allclose(out["diagnostics"]["chain"][key], ref["chain"][key], rtol=1e-10)
    np.testing.assert_allclose(out["trgt_fn"], ref["trgt_fn"], rtol=1e-10)
    np.testing.assert_array_equal(out["diagnostics"]["accept_m"], ref["accept_m"])


_SCHEDULE_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r)
import spwombling_b200 as spw
from oracle.hlm import hlm_bayes_sp, r_chol
import oracle
n, niter = 1100, 4
rng = np.random.Generator(np.random.Philox(77))
coords = rng.random((n, 2))
y = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1])) + rng.standard_normal(n)
y = y - y.mean()
norm, unif = rng.standard_normal(3 * niter), rng.random(3 * niter)
ref = hlm_bayes_sp(y, coords, niter=niter, nburn=1, report=2, cov_type="matern2", norm=norm, unif=unif, trgt_fn_compute=True)
out = spw.hlmBayes_sp(y=y, coords=coords, niter=niter, nburn=1, report=2, cov_type="matern2", norm=norm, unif=unif,
                      trgtFn_compute=True, verbose=False)
for key in ("sig2", "tau2", "phi"):
    np.testing.assert_allclose(out["diagnostics"]["chain"][key], ref["chain"][key], rtol=1e-10)
np.testing.assert_allclose(out["trgt_fn"], ref["trgt_fn"], rtol=1e-10)
A = oracle.cov_matern1(oracle.dist_matrix(coords), 8.0, 2.0) + 0.5 * np.eye(n)
U, _ = spw.chol(A)
Uo = r_chol(A)
assert np.max(np.abs(U - Uo)) / np.max(np.abs(Uo)) < 1e-11
print("ok")
"""


@pytest.mark.parametrize("mode", ["0", "1", "2", "3"])
def test_cholesky_schedules_against_oracle(mode):
    """Every single-matrix Cholesky schedule (SPW_LOOKAHEAD 0..3; the switch is read once per process, hence the
    subprocess) with the bulk split forced on at a size the oracle handles: ragged pieces, a partial last outer block
    and the appended y row, against the reference's arithmetic."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SPW_LOOKAHEAD=mode, SPW_SHADOW_MIN="1")
    res = subprocess.run([sys.executable, "-c", _SCHEDULE_SCRIPT % {"root": root}], env=env, capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0 and res.stdout.strip().endswith("ok"), res.stdout[-2000:] + res.stderr[-4000:]
