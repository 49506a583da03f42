"""Oracle: distance matrix and the three covariance kernels (test infrastructure only).

Follows, operation for operation:
  * ``as.matrix(dist(coords))``        R/hlmBayes_sp.R:38, R/spatial_gradient.R:23
  * ``cov_gaussian``                   R/cov_gaussian.R:10-17
  * ``cov_matern1``                    R/cov_matern1.R:10-17
  * ``cov_matern2``                    R/cov_matern2.R:10-17
"""
import numpy as np

COV_TYPES = ("gaussian", "matern1", "matern2")


def dist_matrix(coords):
    """Euclidean distance matrix with an exactly-zero diagonal (R ``dist``: sqrt(sum(dev*dev)))."""
    coords = np.asarray(coords, dtype=np.float64)
    d0 = coords[:, None, 0] - coords[None, :, 0]
    d1 = coords[:, None, 1] - coords[None, :, 1]
    return np.sqrt(d0 * d0 + d1 * d1)


def cov_gaussian(Delta, phi, sig2):
    # R/cov_gaussian.R:14  Sigma = sig2 * exp(-phi * Delta^2)
    return sig2 * np.exp(-phi * (Delta * Delta))


def cov_matern1(Delta, phi, sig2):
    # R/cov_matern1.R:14  sig2 * (1 + phi * sqrt(3) * Delta) * exp(-phi * sqrt(3) * Delta)
    a = phi * np.sqrt(3.0)
    return sig2 * (1.0 + a * Delta) * np.exp(-a * Delta)


def cov_matern2(Delta, phi, sig2):
    # R/cov_matern2.R:14  sig2 * (1 + phi*sqrt(5)*Delta + phi^2*5*Delta^2/3) * exp(-phi*sqrt(5)*Delta)
    a = phi * np.sqrt(5.0)
    return sig2 * (1.0 + a * Delta + (phi * phi) * 5.0 * (Delta * Delta) / 3.0) * np.exp(-a * Delta)


def cov_by_name(cov_type):
    if cov_type == "gaussian":
        return cov_gaussian
    if cov_type == "matern1":
        return cov_matern1
    if cov_type == "matern2":
        return cov_matern2
    # R/hlmBayes_sp.R:52 -- anything else is the broken general-Matern path
    raise ValueError("cov.type must be one of 'gaussian', 'matern1', 'matern2'")
