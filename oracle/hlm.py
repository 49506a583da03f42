"""Oracle: collapsed Metropolis-within-Gibbs sampler ``hlmBayes_sp`` (test infrastructure only).

Follows R/hlmBayes_sp.R:27-339 for the three named covariance types, with the random variates
injected (the reference draws ``rnorm(1)``, ``runif(1)`` three times per iteration in the fixed
order sigma2, tau2, phi -- R/hlmBayes_sp.R:112,123,138,149,162,192; the phi uniform is consumed
even when the proposal is out of bounds).  Linear algebra goes through the same LAPACK routines
base R reaches: ``chol`` = dpotrf('U'), ``chol2inv`` = dpotri('U').
"""
import numpy as np
from scipy.linalg import lapack

from .cov import cov_by_name, dist_matrix


def r_chol(A):
    """R ``chol(A)``: upper factor U with U'U = A, reads the upper triangle only; error if not PD."""
    U, info = lapack.dpotrf(np.asarray(A, dtype=np.float64, order="F"), lower=0, clean=1)
    if info != 0:
        raise np.linalg.LinAlgError("the leading minor of order %d is not positive" % info)
    return U


def r_chol2inv(U):
    """R ``chol2inv(U)``: full symmetric inverse of U'U via dpotri."""
    inv, info = lapack.dpotri(np.asarray(U, dtype=np.float64, order="F"), lower=0)
    if info != 0:
        raise np.linalg.LinAlgError("dpotri info=%d" % info)
    iu = np.triu_indices(inv.shape[0], 1)
    inv[(iu[1], iu[0])] = inv[iu]
    return inv


def _step_factor(x):
    # R/hlmBayes_sp.R:287-293
    out = 1.0
    x = max(0.17, min(x, 0.75))
    if x > 0.50:
        out = x / 0.50
    elif x < 0.25:
        out = x / 0.25
    return out


def hlm_bayes_sp(y, coords, niter=None, nburn=None, report=None,
                 a_sigma=None, b_sigma=None, a_tau=None, b_tau=None,
                 lower_phi=None, upper_phi=None, cov_type=None,
                 norm=None, unif=None, trgt_fn_compute=False, keep_trace=False):
    """Run the chain with injected variates ``norm[3*niter]``, ``unif[3*niter]``.

    Returns a dict with the reference's return structure (R/hlmBayes_sp.R:321-337: ``y``, ``coords``,
    [``trgt_fn``], ``nu``, ``sig2``, ``tau2``, ``phi`` -- post-burn-in slices) plus diagnostics the
    parity tests use: ``chain`` (full-length sig2/tau2/phi), ``accept_m`` (rows of acceptance rates
    per report batch, R/hlmBayes_sp.R:245-246), ``tuning`` (the scales printed at each report) and,
    when ``keep_trace``, the per-step log-ratios ``r`` (NaN where the phi proposal was out of bounds).
    """
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    coords = np.asarray(coords, dtype=np.float64)
    cov = cov_by_name(cov_type)
    Delta = dist_matrix(coords)                                    # :38
    N = y.shape[0]                                                 # :41
    if niter is None:                                              # :43-47
        niter = 5000
        nburn = niter // 2
        report = 100
    e_sig2 = e_tau2 = e_phi = 1e-1                                 # :50
    a_sigma = 2.0 if a_sigma is None else a_sigma                  # :59-64
    a_tau = 2.0 if a_tau is None else a_tau
    b_sigma = 1.0 if b_sigma is None else b_sigma
    b_tau = 1.0 if b_tau is None else b_tau
    lower_phi = 1e-3 if lower_phi is None else lower_phi
    upper_phi = 30.0 if upper_phi is None else upper_phi
    phi = sig2 = tau2 = 1.0                                        # :67
    nu = {"matern1": 1.5, "matern2": 2.5, "gaussian": np.inf}[cov_type]   # :70-72
    norm = np.asarray(norm, dtype=np.float64).reshape(-1)
    unif = np.asarray(unif, dtype=np.float64).reshape(-1)
    if norm.shape[0] < 3 * niter or unif.shape[0] < 3 * niter:
        raise ValueError("need 3*niter normal and uniform variates")

    accept_p = np.zeros(3)                                         # :76
    res_sig2 = np.zeros(niter)
    res_tau2 = np.zeros(niter)
    res_phi = np.zeros(niter)                                      # :80
    accept_m, tuning = [], []
    trgt = np.zeros(niter) if trgt_fn_compute else None
    trace_r = np.full((niter, 3), np.nan) if keep_trace else None

    R = cov(Delta, phi, 1.0)                                       # :88-96
    I = np.eye(N)
    chol_S = r_chol(sig2 * R + tau2 * I)                           # :99
    inv_S = r_chol2inv(chol_S)                                     # :100

    def target():
        # R/hlmBayes_sp.R:103-105 and :235-237 (sic: "- -(a.tau+1)*log(tau2)" and "/4")
        return (-(a_sigma + 1) * np.log(sig2) - b_sigma / sig2
                - -(a_tau + 1) * np.log(tau2) - b_tau / tau2
                - np.sum(np.log(np.diag(chol_S))) / 4 - y @ inv_S @ y / 2)

    trgt_init = target() if trgt_fn_compute else None

    for i in range(1, niter + 1):                                  # :108
        k = 3 * (i - 1)
        # ---- sigma2 ----
        lsig2 = np.log(sig2) + e_sig2 * norm[k]                    # :112
        chol_d = r_chol(np.exp(lsig2) * R + tau2 * I)              # :114
        inv_d = r_chol2inv(chol_d)                                 # :115
        r = (-b_sigma * (np.exp(-lsig2) - 1 / sig2) - (a_sigma + 1) * (lsig2 - np.log(sig2))
             - (y @ (inv_d - inv_S)) @ y / 2
             - np.sum(np.log(np.diag(chol_d) / np.diag(chol_S))))  # :117-119
        if keep_trace:
            trace_r[i - 1, 0] = r
        if np.log(unif[k]) < min(r, 0.0):                          # :121-123
            sig2 = np.exp(lsig2)
            chol_S, inv_S = chol_d, inv_d
            accept_p[0] += 1
        res_sig2[i - 1] = sig2
        # ---- tau2 ----
        ltau2 = np.log(tau2) + e_tau2 * norm[k + 1]                # :138
        chol_d = r_chol(sig2 * R + np.exp(ltau2) * I)              # :140
        inv_d = r_chol2inv(chol_d)
        r = (-b_tau * (np.exp(-ltau2) - 1 / tau2) - (a_tau + 1) * (ltau2 - np.log(tau2))
             - (y @ (inv_d - inv_S)) @ y / 2
             - np.sum(np.log(np.diag(chol_d) / np.diag(chol_S))))  # :143-145
        if keep_trace:
            trace_r[i - 1, 1] = r
        if np.log(unif[k + 1]) < min(r, 0.0):                      # :147-149
            tau2 = np.exp(ltau2)
            chol_S, inv_S = chol_d, inv_d
            accept_p[1] += 1
        res_tau2[i - 1] = tau2
        # ---- phi ----
        lphi = np.log(phi) + e_phi * norm[k + 2]                   # :162
        if (np.exp(lphi) < lower_phi) or (np.exp(lphi) > upper_phi):   # :164
            accept_prob = -np.inf
        else:
            R_d = cov(Delta, np.exp(lphi), 1.0)                    # :170-178
            chol_d = r_chol(sig2 * R_d + tau2 * I)                 # :181
            inv_d = r_chol2inv(chol_d)
            r = (-(lphi - np.log(phi)) - (y @ (inv_d - inv_S)) @ y / 2
                 - np.sum(np.log(np.diag(chol_d) / np.diag(chol_S))))   # :184-185
            if keep_trace:
                trace_r[i - 1, 2] = r
            accept_prob = min(r, 0.0)
        if np.log(unif[k + 2]) < accept_prob:                      # :192
            phi = np.exp(lphi)
            R = R_d
            chol_S, inv_S = chol_d, inv_d
            accept_p[2] += 1
        res_phi[i - 1] = phi
        if trgt_fn_compute:
            trgt[i - 1] = target()                                 # :234-238
        if i % report == 0:                                        # :244
            accept_p = accept_p / report                           # :245
            accept_m.append(accept_p.copy())                       # :246
            tuning.append([e_sig2, e_tau2, e_phi])                 # the scales cat() prints, :268
            step = [_step_factor(x) for x in accept_p]             # :287-293
            e_sig2 *= step[0]
            e_tau2 *= step[1]
            e_phi *= step[2]                                       # :294-296
            accept_p = np.zeros(3)                                 # :299

    out = {"y": y, "coords": coords, "nu": nu,
           "sig2": res_sig2[nburn:niter], "tau2": res_tau2[nburn:niter], "phi": res_phi[nburn:niter],
           "chain": {"sig2": res_sig2, "tau2": res_tau2, "phi": res_phi},
           "accept_m": np.asarray(accept_m).reshape(-1, 3), "tuning": np.asarray(tuning).reshape(-1, 3),
           "final_scales": np.array([e_sig2, e_tau2, e_phi])}
    if trgt_fn_compute:
        out["trgt_fn"] = np.concatenate([[trgt_init], trgt])       # :325
    if keep_trace:
        out["trace_r"] = trace_r
    return out
