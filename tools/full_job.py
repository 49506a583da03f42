"""Run a BASELINE.json configuration in full through the public entry point and check it.

    python tools/full_job.py c5 --ngpu 8      # N=5000, matern2, 100x100 grid, 5000 posterior samples
    python tools/full_job.py c4 --ngpu 1      # N=2000, gaussian, 60x60 grid, 5000 posterior samples

Checks: (i) mean / var / draws of two samples against an independent fp64 evaluation on the GPU (torch.linalg:
cuSOLVER Cholesky + triangular solves), 1e-10 relative per block; (ii) the returned estimates against the oracle's
batch-median / HPD summary of the returned draws on a subset of grid points (bit-exact)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import spwombling_b200 as spw
from oracle.summary import summarise_draws

ap = argparse.ArgumentParser()
ap.add_argument("workload")
ap.add_argument("--ngpu", type=int, default=1)
ap.add_argument("--samples", type=int, default=5000)
ap.add_argument("--nbatch", type=int, default=200)
a = ap.parse_args()
w = bench.WORKLOADS[a.workload]
S = a.samples
t0 = time.perf_counter()
coords, grid, phi, sig2, z, eps = bench.make_inputs(w, S)
t_in = time.perf_counter() - t0
c = bench.ncomp(w["cov"])
G = grid.shape[0]
model = {"phi": phi, "sig2": sig2, "z": z}
t0 = time.perf_counter()
out = spw.spatial_gradient(coords, model, w["cov"], grid, nbatch=a.nbatch, eps=eps, ngpu=a.ngpu, rounded=False)
dt = time.perf_counter() - t0
units = S * G
names = ("s1", "s2", "s11", "s12", "s22")[:c]
# (ii) summary consistency on a subset of grid points
sub = np.linspace(0, G - 1, 25).astype(int)
ok_sum = True
for k, nm in enumerate(names):
    est = summarise_draws(np.asarray(out["grad.%s.mcmc" % nm])[:, sub], a.nbatch)
    key = ("grad1.est", "grad2.est", "grad11.est", "grad12.est", "grad22.est")[k]
    ok_sum &= bool(np.array_equal(est, out[key][sub]))
# (i) two samples against an independent fp64 evaluation (moments recomputed through the library for those samples)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_fullsize import _torch_moments, _blk_err
pick = [0, S - 1]
gsub = grid[sub]
chk = spw.spatial_gradient(coords, {"phi": phi[pick], "sig2": sig2[pick], "z": z[pick]}, w["cov"], gsub, nbatch=2,
                           eps=eps[pick][:, sub], return_moments=True, rounded=False)
errs = []
for j, s in enumerate(pick):
    mean, var = _torch_moments(coords, gsub, w["cov"], phi[s], z[s])
    errs.append(max(_blk_err(chk["mean.grad"][j], mean), _blk_err(chk["var.grad"][j], var)))
    full_d = np.stack([np.asarray(out["grad.%s.mcmc" % nm])[s, sub] for nm in names], axis=1)
    sub_d = np.stack([chk["grad.%s.mcmc" % nm][j] for nm in names], axis=1)
    errs.append(float(np.max(np.abs(full_d - sub_d))))          # same draws whether computed in the full job or alone
print(json.dumps({"workload": w["name"], "samples": S, "grid_points": G, "ngpu": a.ngpu, "nbatch": a.nbatch,
                  "wall_s": dt, "units_per_s_wall": units / dt, "input_generation_s": t_in,
                  "h2d_gb": (eps.nbytes + z.nbytes) / 1e9, "d2h_gb": units * c * 8 / 1e9,
                  "summary_matches_oracle_on_returned_draws": ok_sum,
                  "max_rel_err_vs_independent_fp64": max(errs[0], errs[2]),
                  "draws_identical_full_vs_alone": bool(errs[1] == 0.0 and errs[3] == 0.0)}))
