"""Golden fixtures built from the package's own datasets (meuse, boston.c; tools/make_golden.py).
CPU: the oracle reproduces the committed vectors.  GPU: the CUDA path reproduces them through the C ABI."""
import glob
import os

import numpy as np
import pytest

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def _load(path):
    d = np.load(path)
    return {k: d[k] for k in d.files}


def test_fixtures_present():
    names = [os.path.basename(p) for p in GOLDEN]
    assert names == ["boston_matern2.npz", "meuse_matern1.npz", "meuse_matern2.npz"]


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_golden(path):
    from oracle.gradient import gradient_draws
    from oracle.hlm import hlm_bayes_sp
    d = _load(path)
    cov = str(d["cov_type"])
    draws, mean, var = gradient_draws(d["coords"], d["grid"], cov, d["phi"], d["sig2"], d["z"], d["eps"],
                                      want_moments=True)
    np.testing.assert_allclose(mean, d["mean"], rtol=1e-9, atol=1e-9 * np.abs(d["mean"]).max())
    np.testing.assert_allclose(draws, d["draws"], rtol=1e-9, atol=1e-9 * np.abs(d["draws"]).max())
    niter = int(d["hlm_niter"])
    ch = hlm_bayes_sp(d["y"], d["coords"], niter=niter, nburn=niter // 2, report=int(d["hlm_report"]), cov_type=cov,
                      upper_phi=float(d["hlm_upper_phi"]), norm=d["hlm_norm"], unif=d["hlm_unif"],
                      trgt_fn_compute=True)
    np.testing.assert_allclose(ch["chain"]["phi"], d["hlm_phi"], rtol=1e-9)
    np.testing.assert_allclose(ch["trgt_fn"], d["hlm_trgt"], rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_gpu_reproduces_golden(path):
    import spwombling_b200 as spw
    d = _load(path)
    cov = str(d["cov_type"])
    c = d["eps"].shape[2]
    out = spw.spatial_gradient(d["coords"], {"phi": d["phi"], "sig2": d["sig2"], "z": d["z"]}, cov, d["grid"],
                               nbatch=2, eps=d["eps"], return_moments=True, rounded=False)
    names = ("s1", "s2", "s11", "s12", "s22")[:c]
    got = np.stack([out["grad.%s.mcmc" % nm] for nm in names], axis=2)
    blk = lambda a: np.max(np.abs(a), axis=-1, keepdims=True)      # noqa: E731  (normalise per 5-vector / block)
    assert np.max(np.abs(out["mean.grad"] - d["mean"]) / blk(d["mean"])) < 1e-10
    assert np.max(np.abs(got - d["draws"]) / blk(d["draws"])) < 1e-10
    v_got, v_want = out["var.grad"].reshape(got.shape[0], got.shape[1], -1), d["var"].reshape(got.shape[0], got.shape[1], -1)
    assert np.max(np.abs(v_got - v_want) / blk(v_want)) < 1e-10
    niter = int(d["hlm_niter"])
    ch = spw.hlmBayes_sp(y=d["y"], coords=d["coords"], niter=niter, nburn=niter // 2, report=int(d["hlm_report"]),
                         cov_type=cov, upper_phi=float(d["hlm_upper_phi"]), norm=d["hlm_norm"], unif=d["hlm_unif"],
                         trgtFn_compute=True, verbose=False)
    for key, ref in (("sig2", d["hlm_sig2"]), ("tau2", d["hlm_tau2"]), ("phi", d["hlm_phi"])):
        np.testing.assert_allclose(ch["diagnostics"]["chain"][key], ref, rtol=1e-10)
    np.testing.assert_allclose(ch["trgt_fn"], d["hlm_trgt"], rtol=1e-10)
    np.testing.assert_allclose(ch["diagnostics"]["accept_m"], d["hlm_accept"], atol=1e-15)
