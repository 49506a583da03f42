"""Inputs for oracle/r_pin/run_reference.R (the recipe that pins the oracle to the UNMODIFIED reference on a machine
with R): small seeded cases written as plain CSV (17 significant digits, no header) under tests/golden/r_pin/in/.
The same arrays are rebuilt by tests/test_r_pin.py, so the committed files cannot drift from what the tests feed the
oracle and the GPU path.

    python tools/make_r_pin_inputs.py          # rewrites tests/golden/r_pin/in/*.csv
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "r_pin", "in")
COVS = ("gaussian", "matern1", "matern2")
PHI_RANGE = {"gaussian": (40.0, 80.0), "matern1": (3.0, 8.0), "matern2": (4.0, 9.0)}


def ncomp(cov):
    return 2 if cov == "matern1" else 5


def cases():
    """name -> dict of arrays (all float64, 2-D or 1-D)."""
    out = {}
    rng = np.random.Generator(np.random.Philox(20261020))
    # ---- hlmBayes_sp: one chain per covariance type ----
    n, niter = 40, 12
    coords = rng.random((n, 2))
    y = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1])) + rng.standard_normal(n)
    y = y - y.mean()
    for cov in COVS:
        out["hlm_" + cov] = {"coords": coords, "y": y, "norm": rng.standard_normal(3 * niter), "unif": rng.random(3 * niter),
                             "par": np.array([niter, 4, 4])}          # niter, nburn, report
    # ---- spatial_gradient: S samples x G grid points ----
    n, G, S = 30, 4, 6
    coords = rng.random((n, 2))
    grid = rng.uniform(0.1, 0.9, (G, 2))
    f = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1]))
    for cov in COVS:
        c = ncomp(cov)
        out["grad_" + cov] = {"coords": coords, "grid": grid, "phi": rng.uniform(*PHI_RANGE[cov], S),
                              "sig2": rng.uniform(0.5, 2.0, S),
                              "z": (f - f.mean())[None, :] + 0.3 * rng.standard_normal((S, n)),
                              "eps": rng.standard_normal((S, G, c)).reshape(S * G, c),   # rnorm(c) per (sample, grid point)
                              "par": np.array([3])}                     # nbatch
    # ---- spsample (the X-taking definition of R/spsample_beta.R) ----
    n, S, p = 30, 4, 2
    coords = rng.random((n, 2))
    X = np.column_stack([np.ones(n), rng.standard_normal(n)])
    y = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1])) + X @ np.array([1.0, 2.0]) \
        + rng.standard_normal(n)
    out["spsample_matern2"] = {"coords": coords, "X": X, "y": y, "phi": rng.uniform(4.0, 9.0, S),
                               "sig2": rng.uniform(0.5, 3.0, S), "tau2": rng.uniform(0.5, 2.0, S),
                               "eps_beta": rng.standard_normal((S, p)), "eps_z": rng.standard_normal((S, n))}
    # ---- covariance kernels on a distance vector ----
    out["cov"] = {"delta": np.concatenate([[0.0], rng.uniform(0.0, 1.5, 15)]), "par": np.array([3.5, 1.7])}   # phi, sig2
    return out


def write_all(out_dir=OUT):
    os.makedirs(out_dir, exist_ok=True)
    for name, arrs in cases().items():
        for key, a in arrs.items():
            a = np.atleast_2d(np.asarray(a, dtype=np.float64))
            if a.shape[0] == 1 and np.asarray(arrs[key]).ndim == 1:
                a = a.T                                                # vectors are written as one column
            np.savetxt(os.path.join(out_dir, "%s.%s.csv" % (name, key)), a, fmt="%.17g", delimiter=",")


if __name__ == "__main__":
    write_all()
    print("wrote", len(os.listdir(OUT)), "files to", OUT)
