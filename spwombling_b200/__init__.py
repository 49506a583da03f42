"""spwombling_b200 -- B200-native (sm_100a, fp64) implementation of spWombling's two data-parallel paths:
``hlmBayes_sp`` (covariance assembly + Cholesky + log-likelihood ratio per Metropolis step) and
``spatial_gradient`` (per-posterior-sample gradient / curvature inference over a grid), behind a flat C ABI
(``include/spwombling.h``, ``spwombling_b200/csrc``) with this Python mirror of the R interface on top.
"""
from .api import (COV_TYPES, bayes_cwomb, chol, cov_gaussian, cov_matern1, cov_matern2, cov_matrix, cross_cov,  # noqa: F401
                  hlmBayes_sp, hlm_summary, hpd_summary, launch_count, ncomp_of, spatial_gradient, spsample,
                  surface_analysis)
from ._lib import LIB_PATH, NotPositiveDefinite, SpwError, load  # noqa: F401

__version__ = "0.1.0"
