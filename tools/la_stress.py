"""Stress the flag protocol of the look-ahead Cholesky: many factorisations of the same matrices must be bit-identical,
and two runs of a long hlm chain must agree bit for bit: python tools/la_stress.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import spwombling_b200 as spw
import oracle

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
for n in (1100, 1985, 3001, 5000):
    rng = np.random.Generator(np.random.Philox(n))
    coords = rng.random((n, 2))
    A = oracle.cov_matern1(oracle.dist_matrix(coords), 8.0, 2.0) + 0.5 * np.eye(n)
    U0, h0 = spw.chol(A)
    r = np.linalg.norm(U0.T @ U0 - A) / np.linalg.norm(A)
    bad = 0
    for i in range(reps if n < 5000 else max(4, reps // 10)):
        U, h = spw.chol(A)
        bad += not (np.array_equal(U, U0) and h == h0)
    print("n=%d: residual %.2e, %d mismatching repeats" % (n, r, bad), flush=True)
    assert r < 1e-13 and bad == 0
n, niter = 2000, 300
rng = np.random.Generator(np.random.Philox(5))
coords = rng.random((n, 2))
y = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1])) + rng.standard_normal(n)
norm, unif = rng.standard_normal(3 * niter), rng.random(3 * niter)
runs = [spw.hlmBayes_sp(y=y - y.mean(), coords=coords, niter=niter, nburn=0, report=50, cov_type="matern2", norm=norm,
                        unif=unif, verbose=False) for _ in range(2)]
for key in ("sig2", "tau2", "phi"):
    assert np.array_equal(runs[0]["diagnostics"]["chain"][key], runs[1]["diagnostics"]["chain"][key]), key
print("hlm n=%d niter=%d: two runs bit-identical; acceptance %s" % (n, niter, runs[0]["diagnostics"].get("accept", "")))
print("ok")
