"""CPU baseline legs for bench.py: the oracle (the restated reference algorithm: NumPy + OpenBLAS, NOT R)
timed on the host cores.  Test infrastructure only -- bench.py's ``cpu_baseline`` / ``--impl reference`` legs
are the only callers outside tests.

Parallel shape mirrors the reference: ``spatial_gradient`` forks P = cores - 1 workers over contiguous sample
chunks, each single-threaded in BLAS (R/spatial_gradient.R:29-32,200); ``hlmBayes_sp`` is one sequential
process using all BLAS threads.
"""
import multiprocessing as mp
import os
import time

import numpy as np

from .cov import cov_by_name, dist_matrix
from .gradient import cross_cov_rows, ncomp_of, v0_matrix
from .hlm import r_chol, r_chol2inv


def _limit_blas(nthreads):
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=nthreads)
    except Exception:       # pragma: no cover
        return None


def _one_sample(args):
    """One posterior sample of the reference loop on `g_sample` grid points; returns (t_factor, t_grid_total)."""
    coords, grid, cov_type, phi, sig2, z, eps = args
    lim = _limit_blas(1)    # noqa: F841  (kept alive for the duration of the call)
    cov = cov_by_name(cov_type)
    t0 = time.perf_counter()
    Delta = dist_matrix(coords)
    K = cov(Delta, phi, 1.0)                               # R/spatial_gradient.R:215/264/303
    Kinv = r_chol2inv(r_chol(K))                           # :216
    t1 = time.perf_counter()
    V0 = v0_matrix(cov_type, phi)
    out = np.empty((grid.shape[0], ncomp_of(cov_type)))
    for g in range(grid.shape[0]):
        delta = coords - grid[g]
        d = np.sqrt((delta ** 2).sum(axis=1))
        nabla_K, nabla_K_t = cross_cov_rows(cov_type, phi, d, delta)
        tmp = (nabla_K_t @ Kinv).T                         # :238
        mean = tmp.T @ z                                   # :239
        var = V0 - tmp.T @ nabla_K                         # :240
        out[g] = np.sqrt(sig2) * (r_chol(var) @ eps[g]) + mean   # :242
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, float(out.sum())


def gradient_reference_step(coords, grid, cov_type, phi, sig2, z, eps, g_sample, workers):
    """Time `workers` concurrent single-threaded workers, one sample each, on the first `g_sample` grid points,
    and extrapolate each worker's rate to the full grid: rate_w = G / (t_factor_w + G * t_grid_w).
    Returns dict(units_per_s, t_factor, t_grid, wall)."""
    G = grid.shape[0]
    gs = grid[:g_sample]
    jobs = [(coords, gs, cov_type, float(phi[w]), float(sig2[w]), z[w], eps[w][:g_sample]) for w in range(workers)]
    t0 = time.perf_counter()
    if workers == 1:
        res = [_one_sample(jobs[0])]
    else:
        ctx = mp.get_context("fork")
        with ctx.Pool(workers) as pool:
            res = pool.map(_one_sample, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    tf = np.array([r[0] for r in res])
    tg = np.array([r[1] for r in res]) / g_sample
    rate = float(np.sum(G / (tf + G * tg)))
    return {"units_per_s": rate, "t_factor": float(tf.mean()), "t_grid": float(tg.mean()), "wall": wall}


def hlm_reference_iteration(coords, y, cov_type, phi=1.0, sig2=1.0, tau2=1.0):
    """One iteration's linear algebra of R/hlmBayes_sp.R:108-203 with all BLAS threads: three proposals, each
    chol + chol2inv + quadratic form (+ one covariance assembly for the phi proposal).  Returns seconds."""
    cov = cov_by_name(cov_type)
    n = y.shape[0]
    Delta = dist_matrix(coords)
    R = cov(Delta, phi, 1.0)
    I = np.eye(n)
    inv_cur = r_chol2inv(r_chol(sig2 * R + tau2 * I))
    t0 = time.perf_counter()
    acc = 0.0
    for step in range(3):
        Rd = cov(Delta, phi * 1.01, 1.0) if step == 2 else R
        U = r_chol((sig2 * (1.01 if step == 0 else 1.0)) * Rd + (tau2 * (1.01 if step == 1 else 1.0)) * I)
        inv_d = r_chol2inv(U)
        acc += (y @ (inv_d - inv_cur)) @ y / 2 + np.sum(np.log(np.diag(U)))
    return time.perf_counter() - t0, acc


def host_info():
    info = {"cores": os.cpu_count()}
    try:
        from threadpoolctl import threadpool_info
        blas = [d for d in threadpool_info() if d.get("user_api") == "blas"]
        if blas:
            info["blas"] = "%s %s (%s threads)" % (blas[0].get("internal_api"), blas[0].get("version"),
                                                   blas[0].get("num_threads"))
    except Exception:       # pragma: no cover
        pass
    return info
