"""Oracle: posterior surface analysis of the README workflow (test infrastructure only).
Follows README.md:123-177 of the reference: per (sample i, grid point j) ``hess.mat`` from the s11/s12/s22 draws,
``det``, ``eigen(hess.mat)$values`` (decreasing), ``sum(diag(.))``, ``s1 + s2``; then batch medians + HPD + the
significance flag."""
import numpy as np

from .summary import summarise_draws


def surface_analysis(s1, s2, s11, s12, s22, nbatch=100, prob=0.95):
    S, G = s11.shape
    det = np.empty((S, G)); e1 = np.empty((S, G)); e2 = np.empty((S, G))
    for i in range(S):
        for j in range(G):
            H = np.array([[s11[i, j], s12[i, j]], [s12[i, j], s22[i, j]]])
            det[i, j] = np.linalg.det(H)
            w = np.linalg.eigvalsh(H)[::-1]
            e1[i, j], e2[i, j] = w
    derived = {"det": det, "eigen1": e1, "eigen2": e2, "laplace": s11 + s22, "div": s1 + s2}
    out = {}
    for k, v in derived.items():
        est = summarise_draws(v, nbatch, prob)
        sig = np.where((est[:, 1] > 0) & (est[:, 2] > 0), 1.0, np.where((est[:, 1] < 0) & (est[:, 2] < 0), -1.0, 0.0))
        out[k] = np.column_stack([est, sig])
        out[k + ".mcmc"] = v
    return out
