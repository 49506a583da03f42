"""Parity at the BENCHMARKED sizes against the CPU oracle itself (the reference's explicit-inverse op sequence), not
only against an independent fp64 evaluation: the C5 matrix order for spatial_gradient and hlmBayes_sp, spsample at
N >= 1000, and the all-pointer `.C()` entry points against their by-value twins.  All calls go through the C ABI."""
import ctypes

import numpy as np
import pytest

from oracle.gradient import gradient_draws, ncomp_of

pytestmark = pytest.mark.gpu


def _blk_err(got, want):
    got, want = got.reshape(got.shape[0], -1), want.reshape(want.shape[0], -1)
    return np.max(np.max(np.abs(got - want), axis=1) / np.max(np.abs(want), axis=1))


def test_c5_slice_against_oracle():
    """N = 5000, matern2, phi in the C5 range: one posterior sample x 3 grid points through oracle.gradient_draws
    (dpotrf + dpotri + per-grid-point sandwich, R/spatial_gradient.R:303-333).  1e-10 relative per 5-vector / 5x5 block."""
    import spwombling_b200 as spw
    n, cov_type = 5000, "matern2"
    rng = np.random.Generator(np.random.Philox(5005))
    coords = rng.random((n, 2))
    grid = np.array([[0.31, 0.62], [0.05, 0.95], [0.77, 0.18]])
    phi, sig2 = np.array([150.0, 110.0]), np.array([1.3, 0.8])     # the summary needs two batches; sample 0 is checked
    f = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1]))
    z = (f - f.mean())[None, :] + 0.3 * rng.standard_normal((2, n))
    eps = rng.standard_normal((2, 3, 5))
    out = spw.spatial_gradient(coords, {"phi": phi, "sig2": sig2, "z": z}, cov_type, grid, nbatch=2, eps=eps,
                               return_moments=True, rounded=False)
    gd, gm, gv = gradient_draws(coords, grid, cov_type, phi[:1], sig2[:1], z[:1], eps[:1], want_moments=True)
    draws = np.stack([out["grad.%s.mcmc" % nm] for nm in ("s1", "s2", "s11", "s12", "s22")], axis=2)
    assert _blk_err(out["mean.grad"][0], gm[0]) < 1e-10
    assert _blk_err(out["var.grad"][0], gv[0]) < 1e-10
    assert _blk_err(draws[0], gd[0]) < 1e-10


def test_hlm_n5000_chain_against_oracle():
    """hlmBayes_sp at N = 5000 under the DEFAULT single-matrix Cholesky schedule: two full iterations (six proposals
    with accept / reject, the phi proposal re-assembling the matrix) and the target trace against the reference's
    arithmetic (chol + chol2inv + quadratic form, R/hlmBayes_sp.R:108-203,234-238)."""
    import spwombling_b200 as spw
    from oracle.hlm import hlm_bayes_sp
    n, niter = 5000, 2
    rng = np.random.Generator(np.random.Philox(5001))
    coords = rng.random((n, 2))
    y = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1])) + rng.standard_normal(n)
    y = y - y.mean()
    norm = np.array([0.3, -0.2, 0.5, -0.4, 0.1, -0.6])
    unif = np.array([0.9, 0.05, 0.5, 0.3, 0.99, 0.2])
    ref = hlm_bayes_sp(y, coords, niter=niter, nburn=0, report=1, cov_type="matern2", norm=norm, unif=unif,
                       trgt_fn_compute=True)
    out = spw.hlmBayes_sp(y=y, coords=coords, niter=niter, nburn=0, report=1, cov_type="matern2", norm=norm, unif=unif,
                          trgtFn_compute=True, verbose=False)
    for key in ("sig2", "tau2", "phi"):
        np.testing.assert_allclose(out["diagnostics"]["chain"][key], ref["chain"][key], rtol=1e-10)
    np.testing.assert_allclose(out["trgt_fn"], ref["trgt_fn"], rtol=1e-10)
    np.testing.assert_array_equal(out["diagnostics"]["accept_m"], ref["accept_m"])
    # the chain must have moved and rejected at least once, or the test would not exercise both branches
    assert 0 < ref["accept_m"].sum() < 3 * niter


@pytest.mark.parametrize("n,S,cov_type,phir", [(1000, 3, "matern2", (20.0, 40.0)), (1500, 2, "matern1", (10.0, 25.0))])
def test_spsample_large_n_against_oracle(n, S, cov_type, phir):
    """spsample at N >= 1000 (several outer blocks of the batched factorisation, the Gram product of the inverse)."""
    import spwombling_b200 as spw
    from oracle.spsample import spsample as oracle_spsample
    rng = np.random.Generator(np.random.Philox(n))
    coords = rng.random((n, 2))
    X = np.column_stack([np.ones(n), rng.standard_normal(n)])
    y = 10 * (np.sin(3 * np.pi * coords[:, 0]) + np.cos(3 * np.pi * coords[:, 1])) + X @ np.array([1.0, 2.0]) \
        + rng.standard_normal(n)
    phi, sig2, tau2 = rng.uniform(*phir, S), rng.uniform(0.5, 3.0, S), rng.uniform(0.5, 2.0, S)
    eb, ez = rng.standard_normal((S, 2)), rng.standard_normal((S, n))
    ref = oracle_spsample(y, X, coords, phi, sig2, tau2, cov_type, eb, ez)
    got = spw.spsample(y=y, X=X, coords=coords, phi=phi, sig2=sig2, tau2=tau2, cov_type=cov_type, eps_beta=eb, eps_z=ez)
    np.testing.assert_allclose(got["beta"], ref["beta"], rtol=1e-10, atol=1e-12)
    scale = np.abs(ref["z"]).max(axis=1, keepdims=True)
    assert np.max(np.abs(got["z"] - ref["z"]) / scale) < 1e-9
    assert not ref["bad"].any()


# ---- the `.C()` entry points (every argument a pointer) against their by-value twins -------------------------------
def _ip(v):
    return ctypes.byref(ctypes.c_int(int(v)))


def _dp(v):
    return ctypes.byref(ctypes.c_double(float(v)))


def _inputs(n=90, S=6, g=7, seed=11):
    rng = np.random.Generator(np.random.Philox(seed))
    coords = np.asfortranarray(rng.random((n, 2)))
    grid = np.asfortranarray(rng.random((g, 2)))
    return rng, coords, grid


def test_cov_p_matches_cov():
    from spwombling_b200._lib import load, ptr
    lib = load()
    rng, coords, _ = _inputs()
    n = coords.shape[0]
    for code in (0, 1, 2):
        a = np.empty((n, n), order="F")
        b = np.empty((n, n), order="F")
        assert lib.spw_cov(code, n, ptr(coords), 7.5, 1.3, ptr(a)) == 0
        st = ctypes.c_int(-1)
        assert lib.spw_cov_p(_ip(code), _ip(n), ptr(coords), _dp(7.5), _dp(1.3), ptr(b), ctypes.byref(st)) == 0
        assert st.value == 0 and np.array_equal(a, b)
    st = ctypes.c_int(0)
    assert lib.spw_cov_p(_ip(7), _ip(n), ptr(coords), _dp(7.5), _dp(1.3), ptr(b), ctypes.byref(st)) != 0
    assert st.value != 0                      # bad cov type: the status slot carries the error code for .C() callers


def test_hpd_summary_p_matches_hpd_summary():
    from spwombling_b200._lib import load, ptr
    lib = load()
    rng = np.random.Generator(np.random.Philox(12))
    S, g, nbatch = 60, 9, 12
    draws = np.asfortranarray(rng.standard_normal((S, g)))
    a = np.empty((g, 3), order="F")
    b = np.empty((g, 3), order="F")
    assert lib.spw_hpd_summary(S, g, nbatch, 0.95, ptr(draws), ptr(a)) == 0
    st = ctypes.c_int(-1)
    assert lib.spw_hpd_summary_p(_ip(S), _ip(g), _ip(nbatch), _dp(0.95), ptr(draws), ptr(b), ctypes.byref(st)) == 0
    assert st.value == 0 and np.array_equal(a, b)


def test_hlm_run_p_matches_hlm_run():
    from spwombling_b200._lib import load, ptr
    lib = load()
    rng, coords, _ = _inputs(n=260)            # > 200: the blocked multi-kernel chain
    n = coords.shape[0]
    y = rng.standard_normal(n)
    niter, report = 6, 3
    norm, unif = rng.standard_normal(3 * niter), rng.random(3 * niter)
    hyper = np.array([2.0, 1.0, 2.0, 1.0, 1e-3, 30.0])

    def bufs():
        return [np.zeros(niter), np.zeros(niter), np.zeros(niter), np.zeros(niter + 1), np.zeros((2, 3), order="F"),
                np.zeros((2, 3), order="F"), np.zeros(4, dtype=np.int32)]
    A, B = bufs(), bufs()
    rc = lib.spw_hlm_run(n, ptr(coords), ptr(y), 2, niter, 0, report, *hyper, ptr(norm), ptr(unif), 1,
                         *[ptr(x) for x in A])
    assert rc == 0
    st = ctypes.c_int(-1)
    rc = lib.spw_hlm_run_p(_ip(n), ptr(coords), ptr(y), _ip(2), _ip(niter), _ip(0), _ip(report), ptr(hyper), ptr(norm),
                           ptr(unif), _ip(1), *[ptr(x) for x in B], ctypes.byref(st))
    assert rc == 0 and st.value == 0
    for a, b in zip(A, B):
        assert np.array_equal(a, b)
    assert np.ptp(A[0]) > 0                   # the chain moved


def test_gradient_run_p_matches_gradient_run():
    from spwombling_b200._lib import load, ptr
    lib = load()
    rng, coords, grid = _inputs()
    n, g, S, c = coords.shape[0], grid.shape[0], 6, 5
    phi, sig2 = rng.uniform(8, 20, S), rng.uniform(0.5, 2, S)
    z = np.asfortranarray(rng.standard_normal((S, n)))
    eps = rng.standard_normal((S, g, c))

    def bufs():
        return np.empty((c, g, S)), np.empty((c, 3, g)), np.zeros(4, dtype=np.int32)
    d0, s0, i0 = bufs()
    d1, s1, i1 = bufs()
    rc = lib.spw_gradient_run(2, n, ptr(coords), g, ptr(grid), S, ptr(phi), ptr(sig2), ptr(z), ptr(eps), 1, 3, 0.95,
                              ptr(d0), ptr(s0), None, None, ptr(i0))
    assert rc == 0
    st = ctypes.c_int(-1)
    rc = lib.spw_gradient_run_p(_ip(2), _ip(n), ptr(coords), _ip(g), ptr(grid), _ip(S), ptr(phi), ptr(sig2), ptr(z),
                                ptr(eps), _ip(1), _ip(3), _dp(0.95), _ip(1), ptr(d1), ptr(s1), ptr(i1), ctypes.byref(st))
    assert rc == 0 and st.value == 0
    assert np.array_equal(d0, d1) and np.array_equal(s0, s1)
    # want_draws = 0: the summary alone, same numbers
    s2 = np.empty((c, 3, g))
    rc = lib.spw_gradient_run_p(_ip(2), _ip(n), ptr(coords), _ip(g), ptr(grid), _ip(S), ptr(phi), ptr(sig2), ptr(z),
                                ptr(eps), _ip(1), _ip(3), _dp(0.95), _ip(0), None, ptr(s2), ptr(i1), ctypes.byref(st))
    assert rc == 0 and np.array_equal(s0, s2)


_DAG_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r)
import spwombling_b200 as spw
from oracle.hlm import hlm_bayes_sp, r_chol
import oracle
for n in %(orders)r:
    rng = np.random.Generator(np.random.Philox(n))
    coords = rng.random((n, 2))
    A = oracle.cov_matern1(oracle.dist_matrix(coords), 8.0, 2.0) + 0.5 * np.eye(n)
    U, half = spw.chol(A)
    Uo = r_chol(A)
    assert np.max(np.abs(U - Uo)) / np.max(np.abs(Uo)) < 1e-11, n
    U2, _ = spw.chol(A)
    assert np.array_equal(U, U2), "the schedule must be deterministic"
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
print("ok")
"""


@pytest.mark.parametrize("switches, orders", [
    ({"SPW_DAG": "1", "SPW_DAG_MIN": "200"}, (257, 300, 1100, 1985)),
    ({"SPW_LA": "1", "SPW_LA_MIN": "60"}, (65, 130, 257, 300, 513, 1100, 1985)),
    ({"SPW_LA": "1", "SPW_LA_MIN": "60", "SPW_LA_HORIZON": "3"}, (1100, 1985)),
], ids=["tile_dag", "lookahead", "lookahead_horizon3"])
def test_opt_in_single_matrix_cholesky_schedules_against_oracle(switches, orders):
    """The opt-in single-matrix Cholesky schedules -- the persistent tile-DAG kernel (SPW_DAG=1) and the look-ahead
    schedule with a resident diagonal chain (SPW_LA=1).  The switches are read per call but plans are cached per
    process, hence the subprocess: ragged orders, the appended y row of the hlm system, run-to-run determinism."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, **switches)
    res = subprocess.run([sys.executable, "-c", _DAG_SCRIPT % {"root": root, "orders": tuple(orders)}], env=env,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and res.stdout.strip().endswith("ok"), res.stdout[-2000:] + res.stderr[-4000:]


def test_lookahead_cholesky_repeats_are_bit_identical():
    """The default single-matrix schedule from order 1024 on couples kernels that run at the same time through flags
    in global memory: a race there would show as run-to-run differences (tools/la_stress.py is the long version)."""
    import oracle
    import spwombling_b200 as spw
    for n in (1985, 3001):
        rng = np.random.Generator(np.random.Philox(n))
        coords = rng.random((n, 2))
        A = oracle.cov_matern1(oracle.dist_matrix(coords), 8.0, 2.0) + 0.5 * np.eye(n)
        U0, h0 = spw.chol(A)
        assert np.linalg.norm(U0.T @ U0 - A) / np.linalg.norm(A) < 1e-13
        for _ in range(8):
            U, h = spw.chol(A)
            assert np.array_equal(U, U0) and h == h0


_STALL_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r)
import spwombling_b200 as spw
from spwombling_b200._lib import SpwError
import oracle
n = 1100
rng = np.random.Generator(np.random.Philox(n))
coords = rng.random((n, 2))
A = oracle.cov_matern1(oracle.dist_matrix(coords), 8.0, 2.0) + 0.5 * np.eye(n)
try:
    spw.chol(A)
    print("no error")
except SpwError as exc:
    print("error:", exc)
"""


def test_lookahead_cholesky_reports_kernels_that_cannot_run_together():
    """The look-ahead schedule needs its kernels on the device at the same time.  With CUDA_LAUNCH_BLOCKING=1 every
    eager launch waits for the previous one to finish: the resident chain kernel gives up its first wait after two
    seconds and the call must come back with an error that says so -- not hang, not return numbers."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1")
    res = subprocess.run([sys.executable, "-c", _STALL_SCRIPT % {"root": root}], env=env, capture_output=True, text=True,
                         timeout=300)
    assert res.returncode == 0, res.stderr[-3000:]
    assert "did not run at the same time" in res.stdout and "SPW_LA=0" in res.stdout, res.stdout[-2000:]
