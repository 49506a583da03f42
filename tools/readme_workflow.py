"""The README workflow end to end on the GPU path (config C1 of BASELINE.json): timing of each stage."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, spwombling_b200 as spw
rng = np.random.default_rng(1)
coords = rng.random((100, 2))
y = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1])) + rng.standard_normal(100)
yc = y - y.mean()
T = {}
t = time.perf_counter(); spw.chol(np.eye(4)); T["init"] = time.perf_counter() - t
t = time.perf_counter()
mc = spw.hlmBayes_sp(y=yc, coords=coords, niter=10000, nburn=5000, report=100, cov_type="matern2", verbose=False, rng=rng)
T["hlmBayes_sp (1e4 iterations)"] = time.perf_counter() - t
t = time.perf_counter()
zs = spw.spsample(y=yc, X=np.ones((100, 1)), coords=coords, phi=mc["phi"], sig2=mc["sig2"], tau2=mc["tau2"],
                  cov_type="matern2", rng=rng)
T["spsample (5000 samples)"] = time.perf_counter() - t
g = np.arange(0.05, 0.951, 0.05)
grid = np.array([[a, b] for b in g for a in g])
model = {"phi": mc["phi"], "sig2": mc["sig2"], "z": zs["z"]}
t = time.perf_counter()
est = spw.spatial_gradient(coords=coords, model=model, cov_type="matern2", grid_points=grid, nbatch=100, rng=rng)
T["spatial_gradient (5000 x 361)"] = time.perf_counter() - t
t = time.perf_counter(); surf = spw.surface_analysis(est, nbatch=100); T["surface_analysis"] = time.perf_counter() - t
t = time.perf_counter()
womb = spw.bayes_cwomb(coords=coords, model=model, cov_type="matern2", nbatch=100, curve=grid[:19], type="riemann.sum", rng=rng)
T["bayes_cwomb riemann.sum (19 vertices)"] = time.perf_counter() - t
for k, v in T.items():
    print("%-40s %8.3f s" % (k, v))
print("posterior medians: sig2 %.3f tau2 %.3f phi %.3f; grad1 significant at %d of 361 grid points; Gamma1(C) = %s"
      % (np.median(mc["sig2"]), np.median(mc["tau2"]), np.median(mc["phi"]),
         int(np.sum((est["grad1.est"][:, 1] > 0) | (est["grad1.est"][:, 2] < 0))), np.round(womb["womb.measure.inf"][0], 3)))
