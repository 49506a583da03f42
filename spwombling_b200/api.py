"""Host-side mirror of the reference's R interface for the hot path (R is not available in this image, so
the host language above the C ABI is Python; INTEGRATION.md shows the equivalent R glue).

Same names, argument meaning and error behaviour as the R functions, with ``.`` in argument names
spelled ``_`` (``cov.type`` -> ``cov_type``):

  * ``cov_gaussian`` / ``cov_matern1`` / ``cov_matern2``  (R/cov_*.R:10-17)
  * ``hlmBayes_sp``                                        (R/hlmBayes_sp.R:27-339)
  * ``spatial_gradient``                                   (R/spatial_gradient.R:16-442)

Every function calls libspwombling through ctypes with host (NumPy) buffers; there is no CPU fallback.
Random variates are drawn on the host in the reference's order unless they are injected.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import COV_CODES, check, f64, load, ptr

COV_TYPES = ("gaussian", "matern1", "matern2")


def _code(cov_type):
    if cov_type not in COV_CODES:
        # the reference's general-Matern branch for other values is dead code
        raise ValueError("cov.type must be one of 'gaussian', 'matern1', 'matern2'")
    return COV_CODES[cov_type]


def ncomp_of(cov_type):
    return 2 if cov_type == "matern1" else 5


# ---- covariance kernels -----------------------------------------------------------------------------
def _cov_delta(code, Delta, phi, sig2):
    D = f64(Delta, "C")
    out = np.empty_like(D)
    check(load().spw_cov_delta(code, D.size, ptr(D), float(phi), float(sig2), ptr(out)))
    return out


def cov_gaussian(Delta=None, phi=None, sig2=None):
    """``sig2 * exp(-phi * Delta^2)`` -- R/cov_gaussian.R:10-17."""
    return _cov_delta(0, Delta, phi, sig2)


def cov_matern1(Delta=None, phi=None, sig2=None):
    """Matern nu = 3/2 -- R/cov_matern1.R:10-17."""
    return _cov_delta(1, Delta, phi, sig2)


def cov_matern2(Delta=None, phi=None, sig2=None):
    """Matern nu = 5/2 -- R/cov_matern2.R:10-17."""
    return _cov_delta(2, Delta, phi, sig2)


def cov_matrix(coords, cov_type, phi, sig2=1.0):
    """``cov_<type>(as.matrix(dist(coords)), phi, sig2)`` without materialising the distance matrix."""
    coords = f64(coords, "F")
    n = coords.shape[0]
    out = np.empty((n, n), dtype=np.float64, order="F")
    check(load().spw_cov(_code(cov_type), n, ptr(coords), float(phi), float(sig2), ptr(out)))
    return out


def cross_cov(coords, grid_points, cov_type, phi):
    """``nabla.K.t`` for every grid point: array [G, ncomp, N] (R/spatial_gradient.R:233-236, :272, :324-327)."""
    coords = f64(coords, "F")
    grid = f64(grid_points, "F")
    n, g, c = coords.shape[0], grid.shape[0], ncomp_of(cov_type)
    out = np.empty((g, c, n), dtype=np.float64)
    check(load().spw_cross_cov(_code(cov_type), n, ptr(coords), g, ptr(grid), float(phi), ptr(out)))
    return out


def chol(A):
    """R ``chol(A)``: upper factor from the upper triangle; returns (U, sum(log(diag U)))."""
    A = f64(A, "F")
    n = A.shape[0]
    U = np.empty((n, n), dtype=np.float64, order="F")
    ld = ctypes.c_double(0.0)
    info = ctypes.c_int(0)
    rc = load().spw_chol(n, ptr(A), ptr(U), ctypes.byref(ld), ctypes.byref(info))
    check(rc, [0, info.value])
    return U, ld.value


# ---- hlmBayes_sp ----------------------------------------------------------------------------------------
def hlmBayes_sp(y=None, coords=None, niter=None, nburn=None, report=None,
                a_sigma=None, b_sigma=None, a_tau=None, b_tau=None,
                lower_phi=None, upper_phi=None, cov_type=None,
                verbose=True, trgtFn_compute=False, digits=3,
                norm=None, unif=None, rng=None):
    """Collapsed Metropolis-within-Gibbs sampler for (sigma2, tau2, phi) -- R/hlmBayes_sp.R:27-339.

    Returns the reference's list as a dict: ``y``, ``coords``, [``trgt_fn``], ``nu``, ``sig2``, ``tau2``,
    ``phi`` (post-burn-in slices ``(nburn+1):niter``), plus ``diagnostics`` (full chains, acceptance
    matrix ``accept_m`` and the proposal scales in force during each report batch).
    ``norm`` / ``unif`` inject the ``rnorm(1)`` / ``runif(1)`` variates in the reference's order
    (index ``3*(i-1) + {0, 1, 2}`` for the sigma2 / tau2 / phi updates of iteration i).
    """
    code = _code(cov_type)
    y = f64(np.asarray(y, dtype=np.float64).reshape(-1), "C")
    coords = f64(coords, "F")
    n = y.shape[0]
    if coords.shape != (n, 2):
        raise ValueError("coords must be N x 2 with N = length(y)")
    if niter is None:                       # R/hlmBayes_sp.R:43-47
        niter, nburn, report = 5000, 2500, 100
    if nburn is None or report is None:
        raise ValueError("if niter is given, nburn and report must be given too (R/hlmBayes_sp.R:43-47)")
    niter, nburn, report = int(niter), int(nburn), int(report)
    a_sigma = 2.0 if a_sigma is None else float(a_sigma)    # R/hlmBayes_sp.R:59-64
    a_tau = 2.0 if a_tau is None else float(a_tau)
    b_sigma = 1.0 if b_sigma is None else float(b_sigma)
    b_tau = 1.0 if b_tau is None else float(b_tau)
    lower_phi = 1e-3 if lower_phi is None else float(lower_phi)
    upper_phi = 30.0 if upper_phi is None else float(upper_phi)
    if norm is None or unif is None:
        rng = np.random.default_rng() if rng is None else rng
        # the reference interleaves rnorm(1), runif(1) per update; any fixed order is equivalent in law
        norm = rng.standard_normal(3 * niter) if norm is None else norm
        unif = rng.random(3 * niter) if unif is None else unif
    norm = f64(np.asarray(norm).reshape(-1), "C")
    unif = f64(np.asarray(unif).reshape(-1), "C")
    if norm.shape[0] < 3 * niter or unif.shape[0] < 3 * niter:
        raise ValueError("need 3*niter normal and uniform variates")
    nrep = niter // report
    o_sig2, o_tau2, o_phi = (np.zeros(niter) for _ in range(3))
    o_trgt = np.zeros(niter + 1)
    o_acc = np.zeros((max(nrep, 1), 3), order="F")
    o_tun = np.zeros((max(nrep, 1), 3), order="F")
    info = np.zeros(4, dtype=np.int32)
    rc = load().spw_hlm_run(n, ptr(coords), ptr(y), code, niter, nburn, report, a_sigma, b_sigma, a_tau, b_tau,
                            lower_phi, upper_phi, ptr(norm), ptr(unif), int(bool(trgtFn_compute)),
                            ptr(o_sig2), ptr(o_tau2), ptr(o_phi), ptr(o_trgt), ptr(o_acc), ptr(o_tun), ptr(info))
    check(rc, info)
    o_acc, o_tun = o_acc[:nrep], o_tun[:nrep]
    if verbose:
        _print_report(o_sig2, o_tau2, o_phi, o_acc, o_tun, report, nburn, digits)
    nu = {"matern1": 1.5, "matern2": 2.5, "gaussian": np.inf}[cov_type]   # R/hlmBayes_sp.R:70-72
    out = {"y": y, "coords": np.asarray(coords)}
    if trgtFn_compute:
        out["trgt_fn"] = o_trgt                                            # R/hlmBayes_sp.R:325
    out.update({"nu": nu, "sig2": o_sig2[nburn:niter], "tau2": o_tau2[nburn:niter], "phi": o_phi[nburn:niter],
                "diagnostics": {"chain": {"sig2": o_sig2, "tau2": o_tau2, "phi": o_phi},
                                "accept_m": o_acc, "tuning": o_tun}})
    return out


def _print_report(sig2, tau2, phi, acc, tun, report, nburn, digits):
    """The per-batch ``cat()`` lines of R/hlmBayes_sp.R:266-281, printed after the device chain returns."""
    def summ(x):
        q = np.quantile(x, [0.5, 0.025, 0.975])
        return "%s ( %s , %s )" % tuple(np.round(q, digits - 1))
    bar = "=+++++++++++++++++++++++++++++++++++++++++++++++++="
    for b in range(acc.shape[0]):
        i = (b + 1) * report
        lo = (nburn - 1) if i > nburn else (i - report)      # res[nburn:i] vs res[(i-report+1):i] (1-based)
        print("Iteration::", i, "Acceptance:", *acc[b], "Tuning:", *np.round(tun[b], digits))
        print(bar)
        print("Sigma2::", summ(sig2[lo:i]), "\t", "Tau2::", summ(tau2[lo:i]), "\t", "Phi::", summ(phi[lo:i]))
        print(bar)


# ---- spatial_gradient -------------------------------------------------------------------------------------
COMP_NAMES = {2: ("s1", "s2"), 5: ("s1", "s2", "s11", "s12", "s22")}
EST_NAMES = {2: ("grad1.est", "grad2.est"), 5: ("grad1.est", "grad2.est", "grad11.est", "grad12.est", "grad22.est")}


def spatial_gradient(coords=None, model=None, cov_type=None, grid_points=None, nbatch=200, return_mcmc=True,
                     eps=None, rng=None, ngpu=1, return_moments=False, rounded=True, prob=0.95):
    """Gradient and curvature assessment over a grid -- R/spatial_gradient.R:16-442 (non-Windows branch).

    ``model`` holds the posterior samples: ``phi`` [S], ``sig2`` [S], ``z`` [S x N] (the reference reads
    ``model$phi``, ``model$sig2``, ``model$z``; its free variables niter/nburn/samples reduce to S).
    Returns a dict keyed like the reference's list: ``grad1.est`` ... (G x 3: estimate, lower, upper; rounded
    to 4 digits, 2 for matern1) and, when ``return_mcmc``, ``grad.s1.mcmc`` ... (S x G).
    ``eps`` [S, G, ncomp] injects the ``rnorm(ncomp)`` variates the reference draws per (sample, grid point).
    ``return_moments`` adds ``mean.grad`` [S, G, ncomp] and ``var.grad`` [S, G, ncomp, ncomp].
    """
    code = _code(cov_type)
    coords = f64(coords, "F")
    grid = f64(grid_points, "F")
    n, g, c = coords.shape[0], grid.shape[0], ncomp_of(cov_type)
    phi = f64(np.asarray(model["phi"]).reshape(-1), "C")
    sig2 = f64(np.asarray(model["sig2"]).reshape(-1), "C")
    S = phi.shape[0]
    z = f64(np.asarray(model["z"]).reshape(S, n), "F")
    if sig2.shape[0] != S:
        raise ValueError("model$phi, model$sig2 and model$z must hold the same number of samples")
    if eps is None:
        rng = np.random.default_rng() if rng is None else rng
        eps = rng.standard_normal((S, g, c))
    eps = f64(np.asarray(eps).reshape(S, g, c), "C")
    draws = np.empty((c, g, S), dtype=np.float64) if return_mcmc else None       # [a][j][s] == S x g col-major per a
    summ = np.empty((c, 3, g), dtype=np.float64)
    mean = np.empty((S, g, c), dtype=np.float64) if return_moments else None
    var = np.empty((S, g, c, c), dtype=np.float64) if return_moments else None
    info = np.zeros(4, dtype=np.int32)
    rc = load().spw_gradient_run(code, n, ptr(coords), g, ptr(grid), S, ptr(phi), ptr(sig2), ptr(z), ptr(eps),
                                 int(ngpu), int(nbatch), float(prob), ptr(draws), ptr(summ), ptr(mean), ptr(var),
                                 ptr(info))
    check(rc, info)
    digits = 2 if cov_type == "matern1" else 4                                    # R/spatial_gradient.R:358 vs :409
    out = {}
    for k in range(c):
        est = np.ascontiguousarray(summ[k].T)                                     # G x 3
        out[EST_NAMES[c][k]] = np.round(est, digits) if rounded else est
    if return_mcmc:
        for k in range(c):
            out["grad.%s.mcmc" % COMP_NAMES[c][k]] = draws[k].T                   # S x G view (column-major)
    if return_moments:
        out["mean.grad"] = mean
        out["var.grad"] = var.transpose(0, 1, 3, 2)                               # blocks are column-major
    return out


# ---- spsample ------------------------------------------------------------------------
def spsample(y=None, X=None, coords=None, phi=None, sig2=None, tau2=None, cov_type=None, silent=True,
             eps_beta=None, eps_z=None, rng=None):
    """Posterior samples of the latent process z (and intercept / regression coefficients) given the
    (phi, sig2, tau2) draws -- R/spsample_beta.R:19-113.  Returns ``{"z": S x N, "beta0": S, "beta": S x p}``
    (the reference ``unlist``s beta into a vector, :108).  The N x N work (three Cholesky factorisations and two
    inverses per sample, :70-93) runs on the GPU; the p x p regression block (:76-78) is evaluated here on the host.
    ``eps_beta`` [S, p] / ``eps_z`` [S, N] inject the ``rnorm(ncol(X))`` / ``rnorm(N)`` variates of :78 / :93.
    """
    if cov_type is None:
        raise ValueError("Please provide a covariance kernel!")              # R/spsample_beta.R:27
    code = _code(cov_type)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    n = y.shape[0]
    X = np.asarray(X, dtype=np.float64).reshape(n, -1)
    p = X.shape[1]
    coords = f64(coords, "F")
    phi = f64(np.asarray(phi).reshape(-1), "C")
    sig2 = f64(np.asarray(sig2).reshape(-1), "C")
    tau2 = f64(np.asarray(tau2).reshape(-1), "C")
    S = phi.shape[0]
    if eps_beta is None or eps_z is None:
        rng = np.random.default_rng() if rng is None else rng
        eps_beta = rng.standard_normal((S, p)) if eps_beta is None else eps_beta
        eps_z = rng.standard_normal((S, n)) if eps_z is None else eps_z
    eps_beta = np.asarray(eps_beta, dtype=np.float64).reshape(S, p)
    eps_z = f64(np.asarray(eps_z).reshape(S, n), "C")
    XtX = X.T @ X                                                             # :30
    Xty = X.T @ (y - y.mean())
    beta = np.empty((S, p))
    for s in range(S):                                                        # p x p block, :76-78
        post_sd = np.linalg.inv(1e-4 * np.eye(p) + XtX / tau2[s])
        post_sd = (post_sd + post_sd.T) / 2
        post_mean = post_sd @ Xty / tau2[s]
        beta[s] = np.linalg.cholesky(post_sd) @ eps_beta[s] + post_mean        # crossprod(chol(.), eps) = L eps
    w = f64((y[None, :] - beta @ X.T) / tau2[:, None], "C")                   # (y - X beta)/tau2, :92
    z = np.empty((n, S), dtype=np.float64)                                    # S x n column-major
    bad = np.zeros(S, dtype=np.int32)
    info = np.zeros(4, dtype=np.int32)
    rc = load().spw_spsample_run(code, n, ptr(coords), S, ptr(phi), ptr(sig2), ptr(tau2), ptr(w), ptr(eps_z),
                                 ptr(z), ptr(bad), ptr(info))
    check(rc, info)
    z = np.ascontiguousarray(z.T)
    beta0 = np.empty(S)
    z_last = None
    for s in range(S):                                                        # "Bad Iteration" rule, :85-89 (one chunk)
        if bad[s]:
            if not silent:
                print("Bad Iteration::", s + 1)
            z[s] = 0.0 if s == 0 else (z[s - 1] if s == 1 else np.median(z[:s], axis=0))
            beta0[s] = z_last.mean() if z_last is not None else np.nan        # :99 reads the last drawn z
        else:
            z_last = z[s]
            beta0[s] = z[s].mean()
    return {"z": z, "beta0": beta0, "beta": beta}


def bayes_cwomb(coords=None, model=None, cov_type=None, nbatch=None, curve=None, type="riemann.sum", approx=None,
                verbose=True, rule_len=10, eps=None, rng=None, ngpu=1, rounded=True, return_moments=False):
    """Curvilinear wombling measures along a polyline -- R/spatial_wombling.R:28-36, ``type = "riemann.sum"`` branches
    (:114-157 matern1, :245-306 matern2, :394-455 gaussian): ``spatial_gradient`` at the curve vertices (GPU path),
    projection on each segment's unit normal, batch medians + HPD per segment, length-weighted measures.
    Returns the reference's list as a dict: ``tval, uval, womb1.mcmc, [womb2.mcmc], womb.grad.inf, [womb.curv.inf],
    womb.measure.inf``.

    ``type = "rectilinear"`` (:159-244 matern2, :308-393 gaussian; ``approx`` must be None like the only branch the
    reference implements): the line-integral measures by quadrature over ``rule_len`` nodes -- cross-covariances
    ``Gamma.*``, the double quadrature of ``K.*``, the ``K^-1`` sandwich and the ``mvrnorm`` draw all on the GPU
    (``spw_wombling_rect_run``).  That branch does not run upstream as written (``model$phis``, ``1:ntimes`` on a
    vector); the semantics here are ntimes = length(model$phi), everything else literally, see oracle/wombling.py.
    matern1 is refused: its ``K.11.m1`` divides 0 by 0, so the reference could only stop in ``eigen``.  ``eps``
    [S, nseg, 2] injects the variates of ``mvrnorm`` (``return_moments`` adds ``mean.womb`` / ``var.womb``)."""
    if type == "rectilinear":
        if approx is not None:
            raise NotImplementedError("bayes_cwomb: only approx = NULL is implemented upstream (R/spatial_wombling.R:47,160,309)")
        return _bayes_cwomb_rectilinear(coords, model, cov_type, nbatch, curve, rule_len, eps, rng, return_moments)
    if type != "riemann.sum":
        raise ValueError("type must be 'rectilinear' or 'riemann.sum'")
    curve = np.asarray(curve, dtype=np.float64)
    d = curve[1:] - curve[:-1]
    tval = np.sqrt((d ** 2).sum(axis=1))                                   # R/spatial_wombling.R:42
    uval = d / tval[:, None]                                               # :43
    u_perp = np.column_stack([uval[:, 1], -uval[:, 0]])                    # :115-119
    g = spatial_gradient(coords=coords, model=model, cov_type=cov_type, grid_points=curve, nbatch=nbatch,
                         return_mcmc=True, eps=eps, rng=rng, ngpu=ngpu)
    s1, s2 = g["grad.s1.mcmc"][:, :-1], g["grad.s2.mcmc"][:, :-1]
    estval = s1 * u_perp[:, 0] + s2 * u_perp[:, 1]                         # :131-133
    S = estval.shape[0]

    def seg_inf(mat, strict_pair):
        inf = hpd_summary(mat, nbatch)                                     # :134-141 (batch medians, median, HPD)
        if rounded:
            inf = np.round(inf, 3)
        if strict_pair:
            sig = np.where((inf[:, 1] > 0) & (inf[:, 2] > 0), 1.0, np.where((inf[:, 1] < 0) & (inf[:, 2] < 0), -1.0, 0.0))
        else:
            sig = np.where(inf[:, 1] > 0, 1.0, np.where(inf[:, 2] < 0, -1.0, 0.0))    # :142-146 (matern1 branch)
        return np.column_stack([inf, sig])

    def measure(inf, mat):                                                 # :148
        per_sample = (mat * tval).sum(axis=1) / tval.sum()
        h = hpd_summary(per_sample[:, None], S)[0]                         # HPD over the S per-sample measures
        return np.array([(inf[:, 0] * tval).sum() / tval.sum(), h[0], h[1], h[2]])

    out = {"tval": tval, "uval": uval, "womb1.mcmc": estval}
    grad_inf = seg_inf(estval, cov_type != "matern1")
    rows = [measure(grad_inf, estval)]
    if cov_type != "matern1":
        s11, s12, s22 = (g["grad.%s.mcmc" % k][:, :-1] for k in ("s11", "s12", "s22"))
        estval1 = s11 * u_perp[:, 0] ** 2 + s22 * u_perp[:, 1] ** 2 + 2 * s12 * u_perp[:, 0] * u_perp[:, 1]   # :266
        out["womb2.mcmc"] = estval1
        out["womb.grad.inf"] = grad_inf
        curv_inf = seg_inf(estval1, True)
        out["womb.curv.inf"] = curv_inf
        rows.append(measure(curv_inf, estval1))
    else:
        out["womb.grad.inf"] = grad_inf
    out["womb.measure.inf"] = np.stack(rows)
    return out


def _bayes_cwomb_rectilinear(coords, model, cov_type, nbatch, curve, rule_len, eps, rng, return_moments):
    if cov_type == "matern1":
        raise NotImplementedError("bayes_cwomb(type='rectilinear', cov.type='matern1'): K.11.m1 divides 0 by 0 on the node "
                                  "pairs with t1 = t2 (R/spatial_wombling.R:605-607), so the reference stops in mvrnorm")
    code = _code(cov_type)
    coords = f64(coords, "F")
    curve = f64(curve, "F")
    n, nseg = coords.shape[0], curve.shape[0] - 1
    phi = f64(np.asarray(model["phi"]).reshape(-1), "C")
    sig2 = f64(np.asarray(model["sig2"]).reshape(-1), "C")
    S = phi.shape[0]
    z = f64(np.asarray(model["z"]).reshape(S, n), "F")
    if eps is None:
        rng = np.random.default_rng() if rng is None else rng
        eps = rng.standard_normal((S, nseg, 2))
    eps = f64(np.asarray(eps).reshape(S, nseg, 2), "C")
    w = np.empty((2, S, nseg), dtype=np.float64)
    mean = np.empty((S, nseg, 2), dtype=np.float64) if return_moments else None
    var = np.empty((S, nseg, 2, 2), dtype=np.float64) if return_moments else None
    info = np.zeros(4, dtype=np.int32)
    rc = load().spw_wombling_rect_run(code, n, ptr(coords), nseg, ptr(curve), int(rule_len), S, ptr(phi), ptr(sig2), ptr(z),
                                      ptr(eps), ptr(w), ptr(mean), ptr(var), ptr(info))
    check(rc, info)
    d = np.asarray(curve)[1:] - np.asarray(curve)[:-1]
    tval = np.sqrt((d ** 2).sum(axis=1))                                   # R/spatial_wombling.R:42-43
    uval = d / tval[:, None]
    out = {"tval": tval, "uval": uval, "womb1.mcmc": w[0].T, "womb2.mcmc": w[1].T}      # nseg x S

    def inf(ws):                                                           # :213-226 / :228-230
        est = hpd_summary(ws, nbatch)                                      # batch medians, their median and HPD
        sig = np.where((est[:, 1] > 0) & (est[:, 2] > 0), 1.0, np.where((est[:, 1] < 0) & (est[:, 2] < 0), -1.0, 0.0))
        lab = np.ceil(np.arange(1, S + 1) / (S / nbatch)).astype(int)      # the split(...) rule of :212
        bm = np.stack([np.median(ws[lab == b], axis=0) for b in np.unique(lab)])
        tot = bm.sum(axis=1) / tval.sum()
        h = hpd_summary(tot[:, None], tot.shape[0])[0]                     # HPD over the per-batch totals
        return np.column_stack([est, sig]), np.array([est[:, 0].sum() / tval.sum(), np.median(tot), h[1], h[2]])
    g_inf, g_row = inf(w[0])
    c_inf, c_row = inf(w[1])
    out.update({"womb.grad.inf": g_inf, "womb.curv.inf": c_inf, "womb.measure.inf": np.stack([g_row, c_row])})
    if return_moments:
        out["mean.womb"], out["var.womb"] = mean, var
    return out


def hlm_summary(sig2=None, tau2=None, phi=None, beta=None, z=None):
    """Posterior medians and 95 % HPD intervals of the parameters and of every latent z column --
    R/hlm_summary.R:19-50.  The O(S N) order statistics run through the same GPU reducer as the gradient summary
    (``coda::HPDinterval`` on the raw samples = one sample per batch).  Returns ``{"param": rows (median, lower.hpd,
    upper.hpd) for sig2, tau2, phi, beta..., "z": N x 4 (median, lower.hpd, upper.hpd, sig)}``."""
    def one(x):
        x = np.asarray(x, dtype=np.float64)
        x = x.reshape(x.shape[0], -1)
        return hpd_summary(x, x.shape[0])
    rows = [one(sig2)[0], one(tau2)[0], one(phi)[0]]                       # R/hlm_summary.R:26-28
    b = one(beta)
    rows.extend(b)                                                          # :29 / :37
    zest = one(z)                                                           # :41
    sig = np.where(zest[:, 1] > 0, 1.0, np.where(zest[:, 2] < 0, -1.0, 0.0))   # :43-47
    return {"param": np.vstack(rows), "z": np.column_stack([zest, sig])}


SURFACE_NAMES = ("det", "eigen1", "eigen2", "laplace", "div")


def surface_analysis(gradient_est, nbatch=100, prob=0.95, return_mcmc=False):
    """Posterior surface analysis of the README workflow (README.md:123-177): Hessian determinant, eigenvalues,
    Laplacian and divergence per (sample, grid point) from the ``grad.s*.mcmc`` matrices of ``spatial_gradient``, each
    summarised by batch medians + HPD with the significance flag of README.md:143-147.
    Returns ``{name: G x 4 array (estimate, lower, upper, sig)}`` (+ ``name + ".mcmc"``: S x G when ``return_mcmc``)."""
    mats = [np.asarray(gradient_est["grad.%s.mcmc" % k], dtype=np.float64) for k in COMP_NAMES[5]]
    S, g = mats[0].shape
    draws = np.ascontiguousarray(np.stack([m.T for m in mats]))          # [5][g][S] == S x g x 5 column-major
    summ = np.empty((5, 3, g), dtype=np.float64)
    derived = np.empty((5, g, S), dtype=np.float64) if return_mcmc else None
    check(load().spw_surface_analysis(S, g, ptr(draws), int(nbatch), float(prob), ptr(summ), ptr(derived)))
    out = {}
    for k, name in enumerate(SURFACE_NAMES):
        est = summ[k].T
        sig = np.where((est[:, 1] > 0) & (est[:, 2] > 0), 1.0, np.where((est[:, 1] < 0) & (est[:, 2] < 0), -1.0, 0.0))
        out[name] = np.column_stack([est, sig])
        if return_mcmc:
            out[name + ".mcmc"] = derived[k].T
    return out


def hpd_summary(draws, nbatch, prob=0.95):
    """Batch medians + median + HPD for one component: draws [S x G] -> [G x 3] (R/spatial_gradient.R:391-419)."""
    d = f64(draws, "F")
    S, g = d.shape
    out = np.empty((g, 3), dtype=np.float64, order="F")
    check(load().spw_hpd_summary(S, g, int(nbatch), float(prob), ptr(d), ptr(out)))
    return np.ascontiguousarray(out)


def launch_count(reset=False):
    return int(load().spw_launch_count(1 if reset else 0))
