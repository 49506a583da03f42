"""Correctness of the single-matrix Cholesky schedule selected by the environment (SPW_LA=1 / SPW_DAG=1 / default)
against the oracle: python tools/la_check.py [orders...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import spwombling_b200 as spw
from oracle.hlm import hlm_bayes_sp, r_chol
import oracle

orders = [int(a) for a in sys.argv[1:]] or [65, 130, 257, 300, 513, 1100, 1985]
for n in orders:
    rng = np.random.Generator(np.random.Philox(n))
    coords = rng.random((n, 2))
    A = oracle.cov_matern1(oracle.dist_matrix(coords), 8.0, 2.0) + 0.5 * np.eye(n)
    U, half = spw.chol(A)
    Uo = r_chol(A)
    err = np.max(np.abs(U - Uo)) / np.max(np.abs(Uo))
    U2, _ = spw.chol(A)
    print("chol n=%d: max rel err %.2e, deterministic %s" % (n, err, np.array_equal(U, U2)), flush=True)
    assert err < 1e-11 and np.array_equal(U, U2), n
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
print("hlm n=%d: chain and target trace agree with the oracle" % n)
print("ok")
