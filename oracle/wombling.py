"""Oracle: ``bayes_cwomb(type = "riemann.sum")`` (test infrastructure only).
Follows R/spatial_wombling.R:37-44 (tval, uval), :114-157 (matern1), :245-306 (matern2; the gaussian branch
:394-455 is the same code): spatial_gradient at the curve vertices, projection on the unit normal of each segment,
batch medians + HPD per segment, length-weighted wombling measures.  rectilinear quadrature is not restated
(broken upstream: ``model$phis``, ``1:ntimes``)."""
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
