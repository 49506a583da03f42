"""Oracle arbiter: 50-digit mpmath evaluation of the reference formulas for small N
(test infrastructure only).

Used (i) to re-derive the known-answer vectors of SURVEY.md section 9.3 independently of the float64
oracle, and (ii) to decide who is right when the float64 oracle and the GPU path disagree by more
than the tolerance because K is ill-conditioned: the GPU result is accepted if
its error against the arbiter is no larger than the float64 oracle's.
Formulas: R/cov_*.R:14; R/spatial_gradient.R:222-242, :270-278, :313-333; R/hlmBayes_sp.R:99-119.
"""
import mpmath as mp
import numpy as np

mp.mp.dps = 50


def _cov(cov_type, phi, d):
    if cov_type == "gaussian":
        return mp.exp(-phi * d * d)
    if cov_type == "matern1":
        a = phi * mp.sqrt(3)
        return (1 + a * d) * mp.exp(-a * d)
    if cov_type == "matern2":
        a = phi * mp.sqrt(5)
        return (1 + a * d + phi * phi * 5 * d * d / 3) * mp.exp(-a * d)
    raise ValueError(cov_type)


def cov_matrix(coords, cov_type, phi, sig2=1):
    coords = [[mp.mpf(float(v)) for v in row] for row in np.asarray(coords)]
    n = len(coords)
    phi = mp.mpf(float(phi))
    K = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            d = mp.sqrt((coords[i][0] - coords[j][0]) ** 2 + (coords[i][1] - coords[j][1]) ** 2)
            K[i, j] = mp.mpf(float(sig2)) * _cov(cov_type, phi, d)
    return K


def _rows(cov_type, phi, d, d1, d2):
    """Left factor rows nabla.K.t[, j] for one data point (sign convention of the reference)."""
    if cov_type == "gaussian":
        e = mp.exp(-phi * d * d)
        return [2 * phi * e * d1, 2 * phi * e * d2, -2 * phi * e * (1 - 2 * phi * d1 * d1),
                4 * phi * phi * e * d1 * d2, -2 * phi * e * (1 - 2 * phi * d2 * d2)]
    if cov_type == "matern1":
        e = mp.exp(-phi * mp.sqrt(3) * d)
        return [3 * phi * phi * e * d1, 3 * phi * phi * e * d2]
    a = mp.sqrt(5) * phi
    e = mp.exp(-a * d)
    q = 1 + a * d
    f = mp.mpf(5) / 3
    return [f * phi * phi * q * e * d1, f * phi * phi * q * e * d2,
            -f * phi * phi * e * (q - 5 * phi * phi * d1 * d1),
            f * 5 * phi ** 4 * e * d1 * d2,
            -f * phi * phi * e * (q - 5 * phi * phi * d2 * d2)]


def _v0(cov_type, phi):
    if cov_type == "matern1":
        return mp.matrix([[3 * phi * phi, 0], [0, 3 * phi * phi]])
    if cov_type == "gaussian":
        g, h = 2 * phi * phi, 4 * phi ** 4
    else:
        g, h = mp.mpf(5) / 3 * phi * phi, mp.mpf(25) / 3 * phi ** 4
    V = mp.zeros(5, 5)
    V[0, 0] = V[1, 1] = g
    V[2, 2] = V[4, 4] = 3 * h
    V[3, 3] = h
    V[2, 4] = V[4, 2] = h
    return V


def _chol_upper(var):
    """Upper Cholesky factor reading the upper triangle only (what LAPACK dpotrf('U') does)."""
    c = var.rows
    U = mp.zeros(c, c)
    for j in range(c):
        for i in range(j + 1):
            s = var[i, j]
            for k in range(i):
                s -= U[k, i] * U[k, j]
            if i == j:
                if s <= 0:
                    raise ArithmeticError("not positive definite")
                U[j, j] = mp.sqrt(s)
            else:
                U[i, j] = s / U[i, i]
    return U


def gradient_point(coords, cov_type, phi, sig2, z, grid_point, eps):
    """One (sample, grid point): returns dict(rows, mean, var, draw) as numpy float64 arrays."""
    coords_np = np.asarray(coords, dtype=np.float64)
    n = coords_np.shape[0]
    phim = mp.mpf(float(phi))
    K = cov_matrix(coords_np, cov_type, phi, 1)
    Kinv = K ** -1
    c = 2 if cov_type == "matern1" else 5
    A = mp.matrix(c, n)
    for j in range(n):
        d1 = mp.mpf(float(coords_np[j, 0])) - mp.mpf(float(grid_point[0]))
        d2 = mp.mpf(float(coords_np[j, 1])) - mp.mpf(float(grid_point[1]))
        d = mp.sqrt(d1 * d1 + d2 * d2)
        r = _rows(cov_type, phim, d, d1, d2)
        for a in range(c):
            A[a, j] = r[a]
    sgn = [-1, -1, 1, 1, 1][:c]
    B = mp.matrix(n, c)                      # nabla.K: gradient columns negated
    for j in range(n):
        for a in range(c):
            B[j, a] = sgn[a] * A[a, j]
    tmp = (A * Kinv).T
    zv = mp.matrix([mp.mpf(float(v)) for v in z])
    mean = tmp.T * zv
    var = _v0(cov_type, phim) - tmp.T * B
    U = _chol_upper(var)
    ev = mp.matrix([mp.mpf(float(v)) for v in eps[:c]])
    draw = mp.sqrt(mp.mpf(float(sig2))) * (U * ev) + mean
    f = lambda M: np.array([[float(M[i, j]) for j in range(M.cols)] for i in range(M.rows)])
    return {"rows": f(A), "mean": f(mean).ravel(), "var": f(var), "draw": f(draw).ravel(),
            "K": f(K)}


def hlm_quantities(coords, y, cov_type, phi, sig2, tau2):
    """y' Sigma^-1 y and sum(log(diag(chol(Sigma)))) for Sigma = sig2 R(phi) + tau2 I."""
    n = len(y)
    S = cov_matrix(coords, cov_type, phi, sig2)
    for i in range(n):
        S[i, i] += mp.mpf(float(tau2))
    yv = mp.matrix([mp.mpf(float(v)) for v in y])
    quad = (yv.T * (S ** -1) * yv)[0, 0]
    logdet_half = mp.log(mp.det(S)) / 2
    return float(quad), float(logdet_half)
