"""hlmBayes_sp timing helper: python tools/hlm_bench.py N NITER [cov_type]  (prints iterations/s per repetition)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import spwombling_b200 as spw

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
niter = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cov = sys.argv[3] if len(sys.argv) > 3 else "matern2"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
rng = np.random.Generator(np.random.Philox(2))
coords = rng.random((n, 2))
y = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1])) + rng.standard_normal(n)
r6 = np.random.Generator(np.random.Philox(6))
norm, unif = r6.standard_normal(3 * niter), r6.random(3 * niter)
kw = dict(y=y, coords=coords, nburn=0, report=max(1, niter // 2), cov_type=cov, verbose=False)
spw.hlmBayes_sp(niter=2, norm=norm, unif=unif, **kw)
best = 0.0
for rep in range(reps):
    spw.launch_count(reset=True)
    t0 = time.perf_counter()
    out = spw.hlmBayes_sp(niter=niter, norm=norm, unif=unif, **kw)
    dt = time.perf_counter() - t0
    print("n=%d niter=%d: %.2f it/s  (%.2f ms/iter, %.2f ms per factorisation, %.2f TFLOP/s in Cholesky, %d launches)"
          % (n, niter, niter / dt, dt / niter * 1e3, dt / niter / 3 * 1e3, niter * n ** 3 / dt / 1e12,
             spw.launch_count()), flush=True)
    best = max(best, niter / dt)
print("best of %d: n=%d %.2f it/s (%.3f ms per factorisation)" % (reps, n, best, 1e3 / best / 3))
