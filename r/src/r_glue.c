/*
 * r_glue.c -- .Call shim between the unchanged R functions of spWombling and libspwombling (include/spwombling.h).
 * Marshalling only: every wrapper checks types, allocates the R result vectors, calls one C-ABI entry point with
 * plain pointers, and turns a non-zero status into an R error (the reference stops in chol() there).
 *
 * NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no R headers (Rinternals.h); tests/test_abi.py compiles it
 * against stub headers with R CMD SHLIB's flags.  Build inside the package with src/Makevars (r/src/Makevars) and
 * useDynLib(spWombling, .registration = TRUE) (r/patches/).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include "spwombling.h"

static void spw_stop(int rc) {
  if (rc != SPW_OK) Rf_error("spWombling (GPU): %s", spw_last_error());
}

/* cov_*(Delta, phi, sig2)  --  R/cov_gaussian.R:10, R/cov_matern1.R:10, R/cov_matern2.R:10 */
SEXP spw_r_cov_delta(SEXP type, SEXP delta, SEXP phi, SEXP sig2) {
  SEXP out = PROTECT(Rf_duplicate(delta));
  spw_stop(spw_cov_delta(Rf_asInteger(type), (long long)XLENGTH(delta), REAL(delta), Rf_asReal(phi),
                         Rf_asReal(sig2), REAL(out)));
  UNPROTECT(1);
  return out;
}

/* hlmBayes_sp hot loop  --  R/hlmBayes_sp.R:99-302.  norm/unif are drawn by the R host in the reference's order. */
SEXP spw_r_hlm_run(SEXP coords, SEXP y, SEXP type, SEXP niter, SEXP nburn, SEXP report, SEXP hyper /* a.s,b.s,a.t,b.t,lo,hi */,
                   SEXP norm, SEXP unif, SEXP want_trgt) {
  const int n = Rf_length(y), it = Rf_asInteger(niter), rep = Rf_asInteger(report), nrep = it / rep;
  const double* h = REAL(hyper);
  SEXP ans = PROTECT(Rf_allocVector(VECSXP, 6));
  SEXP sig2 = PROTECT(Rf_allocVector(REALSXP, it)), tau2 = PROTECT(Rf_allocVector(REALSXP, it));
  SEXP phi = PROTECT(Rf_allocVector(REALSXP, it)), trgt = PROTECT(Rf_allocVector(REALSXP, it + 1));
  SEXP acc = PROTECT(Rf_allocMatrix(REALSXP, nrep > 0 ? nrep : 1, 3)), tun = PROTECT(Rf_allocMatrix(REALSXP, nrep > 0 ? nrep : 1, 3));
  int info[4];
  spw_stop(spw_hlm_run(n, REAL(coords), REAL(y), Rf_asInteger(type), it, Rf_asInteger(nburn), rep, h[0], h[1], h[2], h[3],
                       h[4], h[5], REAL(norm), REAL(unif), Rf_asLogical(want_trgt), REAL(sig2), REAL(tau2), REAL(phi),
                       REAL(trgt), REAL(acc), REAL(tun), info));
  SET_VECTOR_ELT(ans, 0, sig2); SET_VECTOR_ELT(ans, 1, tau2); SET_VECTOR_ELT(ans, 2, phi);
  SET_VECTOR_ELT(ans, 3, trgt); SET_VECTOR_ELT(ans, 4, acc);  SET_VECTOR_ELT(ans, 5, tun);
  UNPROTECT(7);
  return ans;
}

/* spatial_gradient hot loop + summary  --  R/spatial_gradient.R:199-419 */
SEXP spw_r_gradient_run(SEXP type, SEXP coords, SEXP grid, SEXP phi, SEXP sig2, SEXP z, SEXP eps, SEXP ngpu, SEXP nbatch,
                        SEXP want_mcmc) {
  const int ty = Rf_asInteger(type), n = Rf_nrows(coords), g = Rf_nrows(grid), S = Rf_length(phi);
  const int c = (ty == SPW_COV_MATERN1) ? 2 : 5, mc = Rf_asLogical(want_mcmc);
  SEXP ans = PROTECT(Rf_allocVector(VECSXP, 2));
  SEXP summ = PROTECT(Rf_alloc3DArray(REALSXP, g, 3, c));
  SEXP draws = PROTECT(mc ? Rf_alloc3DArray(REALSXP, S, g, c) : Rf_allocVector(REALSXP, 0));
  int info[4];
  spw_stop(spw_gradient_run(ty, n, REAL(coords), g, REAL(grid), S, REAL(phi), REAL(sig2), REAL(z), REAL(eps),
                            Rf_asInteger(ngpu), Rf_asInteger(nbatch), 0.95, mc ? REAL(draws) : NULL, REAL(summ), NULL, NULL,
                            info));
  SET_VECTOR_ELT(ans, 0, summ); SET_VECTOR_ELT(ans, 1, draws);
  UNPROTECT(3);
  return ans;
}

/* spsample, N x N part  --  R/spsample_beta.R:50-100 (w = (y - X beta)/tau2 per sample and eps are n x S) */
SEXP spw_r_spsample_run(SEXP type, SEXP coords, SEXP phi, SEXP sig2, SEXP tau2, SEXP w, SEXP eps) {
  const int n = Rf_nrows(coords), S = Rf_length(phi);
  SEXP ans = PROTECT(Rf_allocVector(VECSXP, 2));
  SEXP z = PROTECT(Rf_allocMatrix(REALSXP, S, n));
  SEXP bad = PROTECT(Rf_allocVector(INTSXP, S));
  int info[4];
  spw_stop(spw_spsample_run(Rf_asInteger(type), n, REAL(coords), S, REAL(phi), REAL(sig2), REAL(tau2), REAL(w), REAL(eps),
                            REAL(z), INTEGER(bad), info));
  SET_VECTOR_ELT(ans, 0, z); SET_VECTOR_ELT(ans, 1, bad);
  UNPROTECT(3);
  return ans;
}

/* posterior surface analysis  --  README.md:123-177; draws is the S x g x 5 array of grad.s*.mcmc */
SEXP spw_r_surface_analysis(SEXP draws, SEXP S, SEXP g, SEXP nbatch) {
  const int s = Rf_asInteger(S), gg = Rf_asInteger(g);
  SEXP summ = PROTECT(Rf_alloc3DArray(REALSXP, gg, 3, 5));
  spw_stop(spw_surface_analysis(s, gg, REAL(draws), Rf_asInteger(nbatch), 0.95, REAL(summ), NULL));
  UNPROTECT(1);
  return summ;
}

/* batch medians + median + coda::HPDinterval for the columns of an S x g matrix  --  R/spatial_gradient.R:391-419,
 * R/hlm_summary.R:26-40 (nbatch = S: HPDinterval on the raw samples) */
SEXP spw_r_hpd_summary(SEXP draws, SEXP nbatch, SEXP prob) {
  const int S = Rf_nrows(draws), g = Rf_length(draws) / (S > 0 ? S : 1);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, g, 3));
  spw_stop(spw_hpd_summary(S, g, Rf_asInteger(nbatch), Rf_asReal(prob), REAL(draws), REAL(out)));
  UNPROTECT(1);
  return out;
}

/* bayes_cwomb(type = "rectilinear")  --  R/spatial_wombling.R:159-244, :308-393; eps is the 2 x nseg x S array of the
 * rnorm() variates mvrnorm consumes; returns the 2 x (nseg x S) measures as an nseg x S x 2 array */
SEXP spw_r_wombling_rect(SEXP type, SEXP coords, SEXP curve, SEXP rule_len, SEXP phi, SEXP sig2, SEXP z, SEXP eps) {
  const int n = Rf_nrows(coords), nseg = Rf_nrows(curve) - 1, S = Rf_length(phi);
  SEXP w = PROTECT(Rf_alloc3DArray(REALSXP, nseg, S, 2));
  int info[4];
  spw_stop(spw_wombling_rect_run(Rf_asInteger(type), n, REAL(coords), nseg, REAL(curve), Rf_asInteger(rule_len), S,
                                 REAL(phi), REAL(sig2), REAL(z), REAL(eps), REAL(w), NULL, NULL, info));
  UNPROTECT(1);
  return w;
}

/* .onUnload */
SEXP spw_r_shutdown(void) {
  spw_shutdown();
  return Rf_allocVector(INTSXP, 0);
}

static const R_CallMethodDef call_methods[] = {
  {"spw_r_cov_delta", (DL_FUNC)&spw_r_cov_delta, 4},
  {"spw_r_hlm_run", (DL_FUNC)&spw_r_hlm_run, 10},
  {"spw_r_gradient_run", (DL_FUNC)&spw_r_gradient_run, 10},
  {"spw_r_spsample_run", (DL_FUNC)&spw_r_spsample_run, 7},
  {"spw_r_surface_analysis", (DL_FUNC)&spw_r_surface_analysis, 4},
  {"spw_r_hpd_summary", (DL_FUNC)&spw_r_hpd_summary, 3},
  {"spw_r_wombling_rect", (DL_FUNC)&spw_r_wombling_rect, 8},
  {"spw_r_shutdown", (DL_FUNC)&spw_r_shutdown, 0},
  {NULL, NULL, 0}
};

void R_init_spWombling(DllInfo* dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
