"""Oracle: ``spatial_gradient`` -- per-posterior-sample gradient / curvature inference over a grid
(test infrastructure only).

Follows the NON-Windows branch of R/spatial_gradient.R (the Windows branch :48-196 uses different
formulas):
  * prelude ``dist.s0`` / ``delta.s0``                         R/spatial_gradient.R:45-46
  * gaussian cross-covariances, V.0, sandwich, draw           R/spatial_gradient.R:215-242
  * matern1                                                   R/spatial_gradient.R:264-278
  * matern2                                                   R/spatial_gradient.R:303-333
  * assembly / batching / HPD / rounding / names              R/spatial_gradient.R:344-442
The normal variates are injected: ``eps[s, g, :]`` is the ``rnorm(5)`` (``rnorm(2)`` for matern1)
the reference draws for sample s, grid point g (g is the inner loop).
"""
import numpy as np

from .cov import cov_by_name, dist_matrix
from .hlm import r_chol, r_chol2inv
from .summary import summarise_draws

I_TILDE = np.array([[3.0, 0.0, 1.0], [0.0, 1.0, 0.0], [1.0, 0.0, 3.0]])   # R/spatial_gradient.R:38-41


def ncomp_of(cov_type):
    return 2 if cov_type == "matern1" else 5


def _adiag(a, b):
    out = np.zeros((a.shape[0] + b.shape[0], a.shape[1] + b.shape[1]))
    out[: a.shape[0], : a.shape[1]] = a
    out[a.shape[0]:, a.shape[1]:] = b
    return out


def v0_matrix(cov_type, phi):
    """``V.0``: R/spatial_gradient.R:226-227 (gaussian, sic 2*phi^2 / 4*phi^4), :271, :317-318."""
    if cov_type == "gaussian":
        return _adiag(2 * phi ** 2 * np.eye(2), 4 * phi ** 4 * I_TILDE)
    if cov_type == "matern1":
        return 3 * phi ** 2 * np.eye(2)
    if cov_type == "matern2":
        return _adiag(5 / 3 * phi ** 2 * np.eye(2), 25 / 3 * phi ** 4 * I_TILDE)
    raise ValueError(cov_type)


def cross_cov_rows(cov_type, phi, d, delta):
    """Return (nabla_K [N x c], nabla_K_t [c x N]) for one grid point.

    ``d`` = dist.s0[, i] (N), ``delta`` = delta.s0[[i]] (N x 2) = coords - grid[i, ].
    """
    d1, d2 = delta[:, 0], delta[:, 1]
    if cov_type == "gaussian":                                     # R/spatial_gradient.R:222-236
        e = np.exp(-phi * d ** 2)
        c11 = -2 * phi * e * (1 - 2 * phi * d1 ** 2)
        c12 = 4 * phi ** 2 * e * d1 * d2
        c22 = -2 * phi * e * (1 - 2 * phi * d2 ** 2)
        g = (2 * phi * e)[:, None] * delta
        nabla_K = np.column_stack([-g, c11, c12, c22])
        nabla_K_t = np.column_stack([g, c11, c12, c22]).T
    elif cov_type == "matern1":                                    # R/spatial_gradient.R:270-272
        e = np.exp(-phi * np.sqrt(3.0) * d)
        nabla_K = (-3 * phi ** 2 * e)[:, None] * delta
        nabla_K_t = -nabla_K.T
    elif cov_type == "matern2":                                    # R/spatial_gradient.R:313-327
        a = np.sqrt(5.0) * phi
        e = np.exp(-a * d)
        g = (phi ** 2 * (1 + a * d) * e)[:, None] * delta
        c11 = -phi ** 2 * e * (1 + a * d - 5 * phi ** 2 * d1 ** 2)
        c12 = 5 * phi ** 4 * e * d1 * d2
        c22 = -phi ** 2 * e * (1 + a * d - 5 * phi ** 2 * d2 ** 2)
        nabla_K = 5 * np.column_stack([-g, c11, c12, c22]) / 3
        nabla_K_t = (5 * np.column_stack([g, c11, c12, c22]) / 3).T
    else:
        raise ValueError(cov_type)
    return nabla_K, nabla_K_t


def gradient_draws(coords, grid, cov_type, phi, sig2, z, eps, want_moments=False, Delta=None):
    """Per-sample loop (R/spatial_gradient.R:208-246 / :257-281 / :296-336).

    phi[S], sig2[S], z[S x N], eps[S x G x c] -> draws[S x G x c]; with ``want_moments`` also the
    deterministic quantities mean[S x G x c] and var[S x G x c x c] (the non-symmetric ``var.grad``).
    """
    coords = np.asarray(coords, dtype=np.float64)
    grid = np.asarray(grid, dtype=np.float64)
    phi = np.asarray(phi, dtype=np.float64).reshape(-1)
    sig2 = np.asarray(sig2, dtype=np.float64).reshape(-1)
    z = np.asarray(z, dtype=np.float64).reshape(phi.shape[0], -1)
    cov = cov_by_name(cov_type)
    c = ncomp_of(cov_type)
    S, G = phi.shape[0], grid.shape[0]
    eps = np.asarray(eps, dtype=np.float64).reshape(S, G, c)
    if Delta is None:
        Delta = dist_matrix(coords)                                # :23
    # prelude :45-46
    delta_s0 = coords[None, :, :] - grid[:, None, :]               # [G][N][2]
    dist_s0 = np.sqrt((delta_s0 ** 2).sum(axis=2))                 # [G][N]
    draws = np.empty((S, G, c))
    means = np.empty((S, G, c)) if want_moments else None
    vars_ = np.empty((S, G, c, c)) if want_moments else None
    for s in range(S):
        K = cov(Delta, phi[s], 1.0)                                # :215 / :264 / :303 (sig2 = 1, no nugget)
        Kinv = r_chol2inv(r_chol(K))                               # :216
        V0 = v0_matrix(cov_type, phi[s])
        for g in range(G):
            nabla_K, nabla_K_t = cross_cov_rows(cov_type, phi[s], dist_s0[g], delta_s0[g])
            tmp = (nabla_K_t @ Kinv).T                             # :238  t(crossprod(t(nabla.K.t), s.grad.in))
            mean = tmp.T @ z[s]                                    # :239
            var = V0 - tmp.T @ nabla_K                             # :240  (non-symmetric)
            U = r_chol(var)                                        # upper triangle only
            draws[s, g] = np.sqrt(sig2[s]) * (U @ eps[s, g]) + mean    # :242  U %*% eps
            if want_moments:
                means[s, g] = mean
                vars_[s, g] = var
    if want_moments:
        return draws, means, vars_
    return draws


COMP_NAMES = {2: ("s1", "s2"), 5: ("s1", "s2", "s11", "s12", "s22")}
EST_NAMES = {2: ("grad1.est", "grad2.est"),
             5: ("grad1.est", "grad2.est", "grad11.est", "grad12.est", "grad22.est")}


def spatial_gradient(coords, model, cov_type, grid_points, nbatch=200, return_mcmc=True, eps=None,
                     rounded=True):
    """Whole function: R/spatial_gradient.R:16-442 (Unix branch).

    ``model`` is a dict with ``phi`` [S], ``sig2`` [S], ``z`` [S x N]; the free variables
    ``niter/nburn/samples`` of the reference reduce to S = len(model['phi']).
    Returns a dict keyed like the reference's list: ``grad1.est`` ... (G x 3 arrays: estimate,
    lower, upper; rounded to 4 digits, 2 for matern1) and ``grad.s1.mcmc`` ... (S x G).
    """
    c = ncomp_of(cov_type)
    draws = gradient_draws(coords, grid_points, cov_type, model["phi"], model["sig2"], model["z"], eps)
    out = {}
    digits = 2 if cov_type == "matern1" else 4                     # :358 vs :409
    for k in range(c):
        est = summarise_draws(draws[:, :, k], nbatch)              # :391-419
        out[EST_NAMES[c][k]] = np.round(est, digits) if rounded else est
    if return_mcmc:
        for k in range(c):
            out["grad.%s.mcmc" % COMP_NAMES[c][k]] = draws[:, :, k]
    return out
