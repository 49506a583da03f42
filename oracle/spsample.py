"""Oracle: ``spsample`` -- Gibbs recovery of the latent process z and the regression coefficients given the
(phi, sig2, tau2) draws (test infrastructure only).  SURVEY.md section 8(f) rank 1: the producer of ``model$z``
that sits between ``hlmBayes_sp`` and ``spatial_gradient``.

Follows R/spsample_beta.R:19-113 (the X-taking definition, which wins at collation; R/spsample.R:16-97 is the
X-free variant) with the variates injected: per sample the reference draws ``rnorm(ncol(X))`` (:78) and then
``rnorm(N)`` (:93).  The forked chunking (:34-38) only matters for the "Bad Iteration" fallback (:85-89), which
looks at the previous rows of the SAME chunk; ``chunks`` reproduces it (default: one chunk).
"""
import numpy as np

from .cov import cov_by_name, dist_matrix
from .hlm import r_chol, r_chol2inv


def spsample(y, X, coords, phi, sig2, tau2, cov_type, eps_beta, eps_z, chunks=1):
    """Returns dict(z [S x N], beta0 [S], beta [S x p], bad [S] bool)."""
    if cov_type is None:
        raise ValueError("Please provide a covariance kernel!")                  # :27
    cov = cov_by_name(cov_type)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    N = y.shape[0]                                                               # :28
    X = np.asarray(X, dtype=np.float64).reshape(N, -1)
    p = X.shape[1]
    Delta = dist_matrix(np.asarray(coords, dtype=np.float64))                    # :29
    XtX = X.T @ X                                                                # :30
    phi = np.asarray(phi, dtype=np.float64).reshape(-1)
    sig2 = np.asarray(sig2, dtype=np.float64).reshape(-1)
    tau2 = np.asarray(tau2, dtype=np.float64).reshape(-1)
    S = phi.shape[0]                                                             # :32
    eps_beta = np.asarray(eps_beta, dtype=np.float64).reshape(S, p)
    eps_z = np.asarray(eps_z, dtype=np.float64).reshape(S, N)
    z_all = np.zeros((S, N))
    beta0 = np.zeros(S)
    beta_all = np.zeros((S, p))
    bad = np.zeros(S, dtype=bool)
    # split(1:niter, ceiling(seq_along(1:niter)/(niter/ncores)))  :35
    lab = np.ceil(np.arange(1, S + 1) / (S / chunks)).astype(np.int64)
    I = np.eye(N)
    for chunk in np.unique(lab):
        ids = np.where(lab == chunk)[0]
        z_mcmc = np.zeros((len(ids), N))                                         # :42
        z_last = None
        for i, s in enumerate(ids, start=1):                                     # :50  (i is 1-based like R)
            R_Z = cov(Delta, phi[s], 1.0)                                        # :51-65
            try:
                R_inv = r_chol2inv(r_chol(R_Z))                                  # :70
            except np.linalg.LinAlgError:
                R_inv = r_chol2inv(r_chol(R_Z + 1e-4 * I))                       # :71
            post_sd_beta = r_chol2inv(r_chol(1e-4 * np.eye(p) + XtX / tau2[s]))  # :76
            post_mean_beta = post_sd_beta @ (X.T @ (y - y.mean())) / tau2[s]     # :77
            beta = r_chol(post_sd_beta).T @ eps_beta[s] + post_mean_beta         # :78  crossprod(chol(.), rnorm)
            beta_all[s] = beta
            Sig_in = R_inv / sig2[s] + I / tau2[s]                               # :83
            try:
                chol_in = r_chol(Sig_in)                                         # :84
            except np.linalg.LinAlgError:
                bad[s] = True                                                    # :85-89  "Bad Iteration"
                if i == 1:
                    z = np.zeros(N)
                elif i == 2:
                    z = z_mcmc[i - 2].copy()
                else:
                    z = np.median(z_mcmc[: i - 1], axis=0)
                z_mcmc[i - 1] = z
                z = None
            else:
                Sig_Z = r_chol2inv(chol_in)                                      # :91
                mu_Z = Sig_Z.T @ (y - X @ beta) / tau2[s]                        # :92
                z = r_chol(Sig_Z).T @ eps_z[s] + mu_Z                            # :93
                z_mcmc[i - 1] = z
                z_last = z
            # :99 mean(z): in the bad branch the R variable z still holds the last drawn vector (error if none)
            beta0[s] = z_last.mean() if z_last is not None else np.nan
        z_all[ids] = z_mcmc
    return {"z": z_all, "beta0": beta0, "beta": beta_all, "bad": bad}
