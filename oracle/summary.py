"""Oracle: batch medians + HPD summary of the gradient draws (test infrastructure only).

Follows R/spatial_gradient.R:391-419 (matern1: :353-358).  Third-party pieces restated here because
they are not under /root/reference (``DESCRIPTION`` lists ``coda`` with no version):
  * ``coda::HPDinterval.mcmc`` (coda 0.19): sort each column; ``gap = max(1, min(n-1, round(n*prob)))``;
    ``init = 1:(n-gap)``; ``ind = which.min(vals[init+gap] - vals[init])`` (first minimum);
    interval = ``(vals[ind], vals[ind+gap])``.
  * base ``median`` (mean of the two middle order statistics for even n);
  * ``split(x, ceiling(seq_along(x) / (length(x)/nbatch)))`` contiguous batching in float arithmetic.
"""
import numpy as np


def batch_index(S, nbatch):
    """0-based batch label per sample: ceiling(seq_along(ngrad) / (length(ngrad)/nbatch)) - 1  (:392)."""
    j = np.arange(1, S + 1, dtype=np.float64)
    lab = np.ceil(j / (S / nbatch)).astype(np.int64)
    # split() orders groups by sorted label; relabel densely
    _, dense = np.unique(lab, return_inverse=True)
    return dense


def r_round_half_even(x):
    return np.rint(x)


def hpd_interval(vals, prob=0.95):
    """coda::HPDinterval for a matrix [n x p] -> (lower[p], upper[p])."""
    vals = np.sort(np.asarray(vals, dtype=np.float64), axis=0)
    n = vals.shape[0]
    if n < 2:
        raise ValueError("obj must have nsamp > 1")
    gap = int(max(1, min(n - 1, r_round_half_even(n * prob))))
    width = vals[gap:, :] - vals[: n - gap, :]
    ind = np.argmin(width, axis=0)                    # first minimum, like which.min
    cols = np.arange(vals.shape[1])
    return vals[ind, cols], vals[ind + gap, cols]


def summarise_draws(draws, nbatch, prob=0.95):
    """draws [S x G] -> [G x 3] (estimate, lower, upper), unrounded.

    batch medians [nbatch x G] (:393), estimate = column medians of the batch medians (:406),
    HPD over the batch medians (:409).
    """
    draws = np.asarray(draws, dtype=np.float64)
    S = draws.shape[0]
    lab = batch_index(S, nbatch)
    nb = int(lab.max()) + 1
    bm = np.empty((nb, draws.shape[1]))
    for b in range(nb):
        bm[b] = np.median(draws[lab == b], axis=0)
    est = np.median(bm, axis=0)
    lo, hi = hpd_interval(bm, prob)
    return np.column_stack([est, lo, hi])
