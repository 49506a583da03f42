"""CPU oracle for the spWombling hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this package.  Nothing under ``spwombling_b200/`` imports it; the product path
fails loudly when the CUDA library is missing.

What it is: a NumPy/SciPy float64 restatement of the reference R package arh926/spWombling for the
two paths named by BASELINE.json (``hlmBayes_sp`` and ``spatial_gradient``), following the reference
operation sequence (explicit ``chol2inv`` inverse, per-grid-point sandwich, upper-triangle-only 5x5
``chol``), each function citing the R file:line it follows.  The reference is pure R and R is not
installed in this image, so the reference itself cannot be executed here or on the GPU box.

PARITY UNPINNED by the reference: arh926/spWombling has no tests, no golden vectors, no seeds and no
examples.  The pins this repo holds instead are
  * 50-digit mpmath known-answer vectors for the reference formulas, checked in
    ``tests/test_oracle_kat.py`` and re-derived independently by ``oracle/arbiter.py``;
  * finite-difference identities tying the cross-covariance rows to derivatives of the kernels;
  * the package's own datasets (meuse, boston.c) converted to fixtures under ``tests/golden`` by
    ``tools/make_golden.py``.
Third-party arithmetic the reference reaches that is not under /root/reference (all unversioned in
its DESCRIPTION): LAPACK dpotrf/dpotri via base R ``chol``/``chol2inv`` (restated with SciPy's
LAPACK bindings), ``coda::HPDinterval`` (restated from coda 0.19's published algorithm in
``oracle/summary.py``), ``magic::adiag`` (block diagonal), base ``median``.
"""
from .cov import COV_TYPES, cov_gaussian, cov_matern1, cov_matern2, cov_by_name, dist_matrix  # noqa: F401
from .hlm import hlm_bayes_sp  # noqa: F401
from .gradient import spatial_gradient, gradient_draws, cross_cov_rows, v0_matrix  # noqa: F401
from .summary import hpd_interval, batch_index, summarise_draws  # noqa: F401
