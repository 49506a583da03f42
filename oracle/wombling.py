"""Oracle: ``bayes_cwomb(type = "riemann.sum")`` (test infrastructure only).
Follows R/spatial_wombling.R:37-44 (tval, uval), :114-157 (matern1), :245-306 (matern2; the gaussian branch
:394-455 is the same code): spatial_gradient at the curve vertices, projection on the unit normal of each segment,
batch medians + HPD per segment, length-weighted wombling measures.

``bayes_cwomb_rectilinear`` restates the ``type = "rectilinear"`` branches (:159-244 matern2, :308-393 gaussian) and
their kernels ``Gamma.*`` / ``K.*`` (:464-711).  Those branches do not run upstream as written; the intended semantics
restated here (and documented in DESIGN.md) are:
  * ``ntimes <- 1:length(model$phis)`` (matern1 / matern2; ``phis`` does not exist) and ``ntimes <- 1:length(model$phi)``
    (gaussian; a vector used as ``ncol=`` and in ``1:ntimes``) both mean ntimes = S = length(model$phi);
  * everything else is taken literally, quirks included: the covariance carries sig2 (``Sig.Z = cov(sig2 = sig2.est)``)
    AND the draw is ``mvrnorm(1, mean, sig2 * var)`` (Q20); the quadrature is a left Riemann sum over the first
    rule.len - 1 of rule.len equally spaced nodes with weight tval / rule.len; the double quadrature drops only the
    last of the rule.len^2 node pairs; ``k.g.21 = -k.g.12``; ``K.1212`` of the gaussian kernel lacks phi in two places;
  * ``MASS::mvrnorm`` = eigen(Sigma, symmetric = TRUE) (LOWER triangle of the non-symmetric 2 x 2 ``sig2 * var``),
    eigenvalues in decreasing order, draw = mu + V diag(sqrt(pmax(ev, 0))) eps, error unless ev >= -1e-6 |ev_1|.  The sign
    of each eigenvector is implementation-defined in LAPACK; here: first component >= 0 (second > 0 if the first is 0);
  * matern1 is not restated: ``K.11.m1`` divides 0 by 0 on the rule.len node pairs with t1 = t2, so ``k.gamma`` is NaN
    and ``mvrnorm`` stops in ``eigen`` for every segment."""
import numpy as np

from .gradient import spatial_gradient
from .summary import batch_index, hpd_interval


def bayes_cwomb(coords, model, cov_type, nbatch, curve, eps, rounded=True):
    curve = np.asarray(curve, dtype=np.float64)
    d = curve[1:] - curve[:-1]
    tval = np.sqrt((d ** 2).sum(axis=1))                                   # :42
    uval = d / tval[:, None]                                               # :43
    u_perp = np.column_stack([uval[:, 1], -uval[:, 0]])                    # :115-119 rev(x); second entry negated
    g = spatial_gradient(coords, model, cov_type, curve, nbatch=nbatch, return_mcmc=True, eps=eps)   # :121-126
    s1, s2 = g["grad.s1.mcmc"][:, :-1], g["grad.s2.mcmc"][:, :-1]
    estval = s1 * u_perp[:, 0] + s2 * u_perp[:, 1]                         # :131-133
    S = estval.shape[0]
    lab = batch_index(S, nbatch)                                           # :134

    def seg_inf(mat, strict_pair):
        bm = np.stack([np.median(mat[lab == b], axis=0) for b in range(lab.max() + 1)])   # :135
        lo, hi = hpd_interval(bm)
        inf = np.column_stack([np.median(bm, axis=0), lo, hi])            # :138-141
        if rounded:
            inf = np.round(inf, 3)
        if strict_pair:
            sig = np.where((inf[:, 1] > 0) & (inf[:, 2] > 0), 1.0, np.where((inf[:, 1] < 0) & (inf[:, 2] < 0), -1.0, 0.0))
        else:   # matern1 branch :142-146
            sig = np.where(inf[:, 1] > 0, 1.0, np.where(inf[:, 2] < 0, -1.0, 0.0))
        return np.column_stack([inf, sig])

    def measure(inf, mat):                                                 # :148
        per_sample = (mat * tval).sum(axis=1) / tval.sum()
        lo, hi = hpd_interval(per_sample[:, None])
        return np.array([(inf[:, 0] * tval).sum() / tval.sum(), np.median(per_sample), lo[0], hi[0]])

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


# ---- type = "rectilinear" ------------------------------------------------------------------------------------------
def _u_perp(u):
    return np.array([u[1], -u[0]])                       # rev(u); second entry negated


def _gamma_kernels(cov_type, delta, up, phi):
    """(Gamma.t.1, Gamma.t.2) at delta = -t u + (sj - s0) for an array of deltas [..., 2]
    (R/spatial_wombling.R:464-487 gaussian, :625-647 matern2).  Gamma.1 = -Gamma.t.1 and Gamma.2 = Gamma.t.2 at the
    same nodes (their delta is the negative)."""
    d1, d2 = delta[..., 0], delta[..., 1]
    if cov_type == "gaussian":
        t0 = -2 * phi
        e = np.exp(-phi * (d1 ** 2 + d2 ** 2))
        g1 = t0 * e * d1 * up[0] + t0 * e * d2 * up[1]
        k11 = t0 * e * (1 - 2 * phi * d1 ** 2)
        k22 = t0 * e * (1 - 2 * phi * d2 ** 2)
        k12 = -2 * phi * t0 * e * d1 * d2
    else:
        t0 = -5 * phi ** 2
        t1 = np.sqrt(5) * phi * np.sqrt(d1 ** 2 + d2 ** 2)
        e = np.exp(-t1)
        g1 = t0 * e * (1 + t1) * d1 / 3 * up[0] + t0 * e * (1 + t1) * d2 / 3 * up[1]
        k11 = t0 * e * (1 + t1 - 5 * phi ** 2 * d1 ** 2) / 3
        k22 = t0 * e * (1 + t1 - 5 * phi ** 2 * d2 ** 2) / 3
        k12 = -5 * phi ** 2 * t0 * e * d1 * d2 / 3
    g2 = k11 * up[0] ** 2 + 2 * k12 * up[0] * up[1] + k22 * up[1] ** 2
    return g1, g2


def _k_terms(cov_type, dt, u, up, phi):
    """(K.11, K.12, K.22) at delta = (t2 - t1) u for one scalar dt (R/spatial_wombling.R:515-575, :676-711)."""
    d1, d2 = dt * u[0], dt * u[1]
    cu = np.array([up[0] ** 2, 2 * up[0] * up[1], up[1] ** 2])
    nd = np.sqrt(d1 ** 2 + d2 ** 2)
    zero = (d1 == 0) and (d2 == 0)
    if cov_type == "gaussian":
        t0 = -2 * phi
        e = np.exp(-phi * nd ** 2)
        k11 = t0 * e * (1 - 2 * phi * d1 ** 2)
        k22 = t0 * e * (1 - 2 * phi * d2 ** 2)
        k12 = -2 * phi * t0 * e * d1 * d2
        K11 = k11 * up[0] ** 2 + 2 * k12 * up[0] * up[1] + k22 * up[1] ** 2
        if zero:
            K12 = 0.0
            n4 = 4 * np.diag([phi ** 2, phi ** 2 / 3, phi ** 2])
        else:
            c = 4 * phi ** 2 * e
            k111 = c * (3 - 2 * phi * d1 ** 2) * d1
            k112 = c * (1 - 2 * phi * d1 ** 2) * d2
            k122 = c * (1 - 2 * phi * d2 ** 2) * d1
            k222 = c * (3 - 2 * phi * d2 ** 2) * d2
            n3 = np.array([[k111, k112, k122], [k112, k122, k222]])
            K12 = float(up @ n3 @ cu)
            k1111 = c * (3 - 12 * phi * d1 ** 2 + 4 * phi ** 2 * d1 ** 4)
            k1212 = c * (1 - 2 * d1 ** 2) * (1 - 2 * d2 ** 2)                      # sic: no phi
            k2222 = c * (3 - 12 * phi * d2 ** 2 + 4 * phi ** 2 * d2 ** 4)
            k1112 = -8 * phi ** 3 * e * (3 - 2 * phi * d1 ** 2) * d1 * d2
            k1222 = -8 * phi ** 3 * e * (3 - 2 * phi * d2 ** 2) * d2 * d1
            n4 = np.array([[k1111, k1112, k1212], [k1212, k1212, k1222], [k1212, k1222, k2222]])
        K22 = float(cu @ n4 @ cu)
    else:
        s5 = np.sqrt(5)
        t0 = -5 * phi ** 2
        t1 = s5 * phi * nd
        e = np.exp(-t1)
        k11 = t0 * e * (1 + t1 - 5 * phi ** 2 * d1 ** 2) / 3
        k22 = t0 * e * (1 + t1 - 5 * phi ** 2 * d2 ** 2) / 3
        k12 = -5 * phi ** 2 * t0 * e * d1 * d2 / 3
        K11 = k11 * up[0] ** 2 + 2 * k12 * up[0] * up[1] + k22 * up[1] ** 2
        if zero:
            K12 = 0.0
            n4 = 25 * np.diag([phi ** 4, phi ** 4 / 3, phi ** 4])
        else:
            c = 25 * phi ** 4 * e
            k111 = c * (3 - s5 * phi * d1 ** 2 / nd) * d1 / 3
            k112 = c * (1 - s5 * phi * d1 ** 2 / nd) * d2 / 3
            k122 = c * (1 - s5 * phi * d2 ** 2 / nd) * d1 / 3
            k222 = c * (3 - s5 * phi * d2 ** 2 / nd) * d2 / 3
            n3 = np.array([[k111, k112, k122], [k112, k122, k222]])
            K12 = float(up @ n3 @ cu)
            k1111 = c * (3 - 6 * s5 * phi * d1 ** 2 / nd + s5 * phi * d1 ** 4 / nd ** 3 + 5 * phi ** 2 * d1 ** 4 / nd ** 2) / 3
            k1212 = c * ((1 - s5 * d1 ** 2 / nd) * (1 - s5 * d2 ** 2 / nd) + s5 * phi * d1 ** 2 * d2 ** 2 / nd ** 3) / 3   # sic
            k2222 = c * (3 - 6 * s5 * phi * d2 ** 2 / nd + s5 * phi * d2 ** 4 / nd ** 3 - 5 * phi ** 2 * d2 ** 4 / nd ** 2) / 3   # sic: "-"
            c5 = 25 * s5 * phi ** 5 * e
            k1112 = c5 * (d1 / nd ** 3 - (3 - s5 * phi * d1 ** 2 / nd) / nd) * d1 ** 2 * d2 / 3
            k1222 = c5 * (d2 / nd ** 3 - (3 - s5 * phi * d2 ** 2 / nd) / nd) * d2 ** 2 * d1 / 3
            n4 = np.array([[k1111, k1112, k1212], [k1212, k1212, k1222], [k1212, k1222, k2222]])
        K22 = float(cu @ n4 @ cu)
    return K11, K12, K22


def mvrnorm2(mu, Sigma, eps, tol=1e-6):
    """MASS::mvrnorm(1, mu, Sigma) for a 2 x 2 Sigma with the variates injected (see the module docstring)."""
    ev, V = np.linalg.eigh(np.asarray(Sigma, dtype=np.float64), UPLO="L")
    ev, V = ev[::-1], V[:, ::-1]                          # R: decreasing order
    if not np.all(ev >= -tol * abs(ev[0])):
        raise np.linalg.LinAlgError("'Sigma' is not positive definite")
    for k in range(2):
        if V[0, k] < 0 or (V[0, k] == 0 and V[1, k] < 0):
            V[:, k] = -V[:, k]
    return np.asarray(mu, dtype=np.float64) + V @ (np.sqrt(np.maximum(ev, 0.0)) * np.asarray(eps, dtype=np.float64))


def rectilinear_moments(coords, curve, cov_type, phi, sig2, z, rule_len=10):
    """mean.womb [nseg, 2] and var.womb [nseg, 2, 2] for ONE posterior sample (R/spatial_wombling.R:166-207 / :315-356)."""
    from .cov import cov_by_name, dist_matrix
    from .hlm import r_chol, r_chol2inv
    coords = np.asarray(coords, dtype=np.float64)
    curve = np.asarray(curve, dtype=np.float64)
    ln = int(rule_len)
    rule = np.linspace(0.0, 1.0, ln)
    d = curve[1:] - curve[:-1]
    tval = np.sqrt((d ** 2).sum(axis=1))
    uval = d / tval[:, None]
    K = cov_by_name(cov_type)(dist_matrix(coords), phi, sig2)
    inv = r_chol2inv(r_chol(K))
    nseg = d.shape[0]
    means, vars_ = np.empty((nseg, 2)), np.empty((nseg, 2, 2))
    for i in range(nseg):
        u, up, s0, tv = uval[i], _u_perp(uval[i]), curve[i], tval[i]
        t = rule * tv                                                    # rule.uv * tval[i]
        delta = -t[None, :, None] * u[None, None, :] + (coords - s0)[:, None, :]      # [N, len, 2]
        g1, g2 = _gamma_kernels(cov_type, delta, up, phi)
        gamma = np.column_stack([g1[:, :-1].sum(axis=1), g2[:, :-1].sum(axis=1)]) * (tv / ln)     # gamma.cov
        tgamma = gamma * np.array([-1.0, 1.0])                           # Gamma.1 = -Gamma.t.1, Gamma.2 = Gamma.t.2
        acc = np.zeros(3)
        npairs = ln * ln
        for p in range(npairs - 1):                                      # expand.grid: Var1 fastest; [-len^2] drops the last
            t1, t2 = rule[p % ln] * tv, rule[p // ln] * tv
            K11, K12, K22 = _k_terms(cov_type, t2 - t1, u, up, phi)
            acc += np.array([-K11, K12, K22])
        kg11, kg12, kg22 = acc * tv ** 2 / ln ** 2
        k_gamma = np.array([[kg11, kg12], [-kg12, kg22]])
        tmp = (tgamma.T @ inv).T
        means[i] = tmp.T @ z
        vars_[i] = k_gamma - tmp.T @ gamma
    return means, vars_, tval, uval


def bayes_cwomb_rectilinear(coords, model, cov_type, nbatch, curve, eps, rule_len=10):
    """``bayes_cwomb(type = "rectilinear", approx = NULL)`` for gaussian / matern2.  ``eps`` [S, nseg, 2] are the
    ``rnorm(2)`` variates ``mvrnorm`` draws per (sample, segment)."""
    if cov_type not in ("gaussian", "matern2"):
        raise ValueError("rectilinear wombling: gaussian or matern2 (matern1 divides 0 by 0 upstream)")
    phi = np.asarray(model["phi"], dtype=np.float64).reshape(-1)
    sig2 = np.asarray(model["sig2"], dtype=np.float64).reshape(-1)
    S = phi.shape[0]
    z = np.asarray(model["z"], dtype=np.float64).reshape(S, -1)
    eps = np.asarray(eps, dtype=np.float64)
    w1 = w2 = None
    for s in range(S):
        mean, var, tval, uval = rectilinear_moments(coords, curve, cov_type, phi[s], sig2[s], z[s], rule_len)
        if w1 is None:
            w1, w2 = np.empty((mean.shape[0], S)), np.empty((mean.shape[0], S))
        for i in range(mean.shape[0]):
            wm = mvrnorm2(mean[i], sig2[s] * var[i], eps[s, i])
            w1[i, s], w2[i, s] = wm
    lab = batch_index(S, nbatch)

    def inf(w):
        bm = np.stack([np.median(w[:, lab == b], axis=1) for b in range(lab.max() + 1)])       # nbatch x nseg
        lo, hi = hpd_interval(bm)
        est = np.column_stack([np.median(bm, axis=0), lo, hi])
        sig = np.where((est[:, 1] > 0) & (est[:, 2] > 0), 1.0, np.where((est[:, 1] < 0) & (est[:, 2] < 0), -1.0, 0.0))
        tot = bm.sum(axis=1) / tval.sum()
        tlo, thi = hpd_interval(tot[:, None])
        return np.column_stack([est, sig]), np.array([est[:, 0].sum() / tval.sum(), np.median(tot), tlo[0], thi[0]])
    g_inf, g_row = inf(w1)
    c_inf, c_row = inf(w2)
    return {"tval": tval, "uval": uval, "womb1.mcmc": w1, "womb2.mcmc": w2, "womb.grad.inf": g_inf,
            "womb.curv.inf": c_inf, "womb.measure.inf": np.stack([g_row, c_row])}
