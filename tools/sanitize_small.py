"""Small end-to-end run for compute-sanitizer (memcheck): every kernel family on tiny inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import spwombling_b200 as spw

rng = np.random.Generator(np.random.Philox(1))
for n, cov, phir in ((37, "matern2", (5, 10)), (130, "matern1", (3, 8)), (70, "gaussian", (60, 120))):
    coords = rng.random((n, 2))
    grid = rng.random((21, 2))
    c = 2 if cov == "matern1" else 5
    S = 5
    model = {"phi": rng.uniform(*phir, S), "sig2": rng.uniform(0.5, 2, S), "z": rng.standard_normal((S, n))}
    out = spw.spatial_gradient(coords, model, cov, grid, nbatch=2, eps=rng.standard_normal((S, 21, c)), return_moments=True)
    y = rng.standard_normal(n)
    spw.hlmBayes_sp(y=y, coords=coords, niter=6, nburn=2, report=3, cov_type=cov, verbose=False, rng=rng,
                    upper_phi=200.0)
    spw.spsample(y=y, X=np.ones((n, 1)), coords=coords, phi=model["phi"], sig2=model["sig2"], tau2=model["sig2"],
                 cov_type=cov, rng=rng)
    spw.cov_matrix(coords, cov, 3.0, 1.2)
    spw.cross_cov(coords, grid, cov, 3.0)
    K = spw.cov_matrix(coords, "matern2", 3.0, 1.0) + np.eye(n)
    spw.chol(K)
    if c == 5:
        spw.surface_analysis(out, nbatch=2)
print("sanitize_small done")
