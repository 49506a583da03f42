/*
 * spwombling.h -- flat C ABI of libspwombling (B200 / sm_100a, fp64).
 *
 * The reference (arh926/spWombling) is a pure-R package with no native layer: its "operator API" is
 * the R function signature and everything below it is base-R linear algebra.
 * This header is the boundary a maintainer inserts between those R signatures and the arithmetic:
 * the R host functions keep their arguments and return lists and reach these entry points through
 * .Call / .C (see INTEGRATION.md for the glue).  Every entry point cites the reference code it
 * replaces (file:line relative to the reference tree).
 *
 * Conventions
 *   - all arrays are fp64, COLUMN-MAJOR exactly as R stores them, sizes are int32;
 *   - pointers are HOST pointers unless the function name ends in _dev (device pointers on the
 *     current CUDA device; used by bench.py for the HBM-resident timing and by the multi-rank host);
 *   - cov type: 0 = "gaussian", 1 = "matern1", 2 = "matern2"; ncomp = 2 for matern1, else 5;
 *   - return value: SPW_OK, or an error code; spw_last_error() gives the message.  A non-positive
 *     definite matrix is SPW_ERR_NOT_PD (the reference raises R's chol() error there; no jitter);
 *   - random variates are INPUTS (the R host draws them in the reference's order), so results are
 *     reproducible against the reference with the same variates.
 *   - there is no CPU fallback: every function needs a CUDA device and fails with SPW_ERR_CUDA.
 *   - threading: R calls from one thread and so should other hosts.  Each device has ONE context (stream,
 *     workspaces, plans): calls that use the same device must not overlap; calls on different devices may
 *     (spw_gradient_run itself drives its ngpu devices from internal threads that join before it returns).
 */
#ifndef SPWOMBLING_H
#define SPWOMBLING_H

#ifdef __cplusplus
extern "C" {
#endif

#define SPW_OK 0
#define SPW_ERR_NOT_PD 1   /* chol() failed: info[0] = sample / step index, info[1] = failing minor (1-based) */
#define SPW_ERR_ARG 2
#define SPW_ERR_CUDA 3

#define SPW_COV_GAUSSIAN 0
#define SPW_COV_MATERN1 1
#define SPW_COV_MATERN2 2

/* library / device management ------------------------------------------------------------------ */
int spw_version(void);
const char* spw_last_error(void);
int spw_device_count(void);
int spw_set_device(int device);
/* frees cached workspaces, streams and plans of the calling process (all devices) */
int spw_shutdown(void);
/* cap for the per-call device workspace in bytes (0 = default: 60 % of free HBM) */
int spw_set_workspace_limit(long long bytes);
/* number of CUDA kernel launches issued by this library since the last call with reset != 0 */
long long spw_launch_count(int reset);

/* 1 if the last multi-GPU spw_gradient_run gathered its shards with NCCL (grouped send/recv over NVLink; libnccl is
 * dlopen()ed, SPW_NCCL=0 disables), 0 if it used peer copies, -1 if no multi-GPU gather has run yet */
int spw_last_gather_used_nccl(void);

/* Phase timing of the spatial_gradient pipeline with CUDA events on the launching stream (off by default).
 * ms / calls have 6 entries: 0 covariance assembly, 1 Cholesky, 2 triangular inverse, 3 Linv z,
 * 4 cross-covariance assembly, 5 fused TRMM + Gram + draw kernel; main_flops = algorithmic flops of phase 5
 * (per launch: samples x grid points x (c n (n+1) + (c (c+1) + 2 c) n)). */
int spw_profile_enable(int on);
int spw_profile_read(double* ms, long long* calls, double* main_flops, int reset);

/* diagnostic: first call enables cycle counters in the diagonal-block kernel, later calls read 16 of them */
int spw_debug_diag_timing(long long* out16);
/* diagnostic (SPW_DAG_TRACE=1): per-task trace of the last persistent single-matrix Cholesky of order n (+ extra
 * appended rows) on the current device; 4 words per task (packed id, start ns, end ns, 2 * SM + sub-worker).
 * Returns the number of words written (<= cap_words). */
long long spw_debug_dag_trace(int n, int extra, unsigned long long* out, long long cap_words);
/* diagnostic: mean CUDA-event time (ms) of one single-matrix factorisation of order n under the schedule the environment selects */
int spw_debug_potrf_time(int n, int extra, int reps, double* ms_out, int* info);
long long spw_debug_la_trace(int n, int extra, unsigned long long* out, long long cap_words);

/* covariance kernels ----------------------------------------------------------------------------
 * spw_cov: as.matrix(dist(coords)) fused with cov_gaussian / cov_matern1 / cov_matern2
 *   replaces R/hlmBayes_sp.R:38 + R/cov_gaussian.R:14, R/cov_matern1.R:14, R/cov_matern2.R:14.
 *   coords [n x 2], out [n x n] = sig2 * rho(phi, ||s_i - s_j||), exact sig2 on the diagonal.
 * spw_cov_delta: the same kernels applied elementwise to a caller-supplied distance array (the
 *   signature cov_*(Delta, phi, sig2) of the reference), len elements. */
int spw_cov(int type, int n, const double* coords, double phi, double sig2, double* out);
int spw_cov_delta(int type, long long len, const double* delta, double phi, double sig2, double* out);

/* spw_cross_cov: cross-covariance of (d1, d2, d11, d12, d22) Z(s0) with Z(data), i.e. the reference's
 *   nabla.K.t [ncomp x n] for every grid point -- R/spatial_gradient.R:233-236 (gaussian), :272
 *   (matern1), :324-327 (matern2), including the prelude dist.s0 / delta.s0 (:45-46).
 *   out is [n x ncomp x g] column-major: out[k + n*(a + ncomp*j)] = nabla.K.t[a, k] at grid point j.
 *   (nabla.K is the same with the first two components negated, R/spatial_gradient.R:222-225.) */
int spw_cross_cov(int type, int n, const double* coords, int g, const double* grid, double phi, double* out);

/* spw_chol: upper Cholesky factor as R's chol(): reads the upper triangle of a [n x n], writes the
 *   upper factor U (zeros below the diagonal) -- R/hlmBayes_sp.R:99.  logdet_half = sum(log(diag U)). */
int spw_chol(int n, const double* a, double* u, double* logdet_half, int* info);

/* hlmBayes_sp ------------------------------------------------------------------------------------
 * The whole collapsed Metropolis-within-Gibbs chain of R/hlmBayes_sp.R:99-302 on the device:
 * covariance assembly, Cholesky, quadratic form and log-determinant per proposal, accept / reject,
 * adaptive scaling every `report` iterations.  norm / unif are the rnorm(1) / runif(1) variates in the
 * reference's order: index 3*(i-1) + {0: sigma2, 1: tau2, 2: phi} for iteration i (1-based).
 * Outputs: out_sig2 / out_tau2 / out_phi [niter] (full chains; the R host slices (nburn+1):niter,
 * R/hlmBayes_sp.R:334-336), out_trgt [niter+1] when want_trgt (R/hlmBayes_sp.R:102-106,234-238,325),
 * out_accept / out_tuning [nrep x 3] column-major with nrep = niter / report (acceptance rates and the
 * scales in force during each batch, R/hlmBayes_sp.R:245-246,268), info[4]. */
int spw_hlm_run(int n, const double* coords, const double* y, int type,
                int niter, int nburn, int report,
                double a_sigma, double b_sigma, double a_tau, double b_tau,
                double lower_phi, double upper_phi,
                const double* norm, const double* unif, int want_trgt,
                double* out_sig2, double* out_tau2, double* out_phi, double* out_trgt,
                double* out_accept, double* out_tuning, int* info);

/* spatial_gradient -------------------------------------------------------------------------------
 * spw_gradient_run: the per-posterior-sample loop of R/spatial_gradient.R:199-338 for samples
 *   1..S: K = cov(phi_s, sig2 = 1) (no nugget), Cholesky, cross-covariances at every grid point,
 *   conditional mean and (non-symmetric) conditional covariance, upper-triangle 5x5 chol, draw
 *   sqrt(sig2_s) * U eps + mean.
 *   phi, sig2 [S]; z [S x n] column-major (model$z); eps [ncomp x g x S] column-major:
 *   eps[a + ncomp*(j + g*s)] = a-th normal variate for (sample s, grid point j) -- the order in
 *   which the reference's inner loops call rnorm(ncomp).
 *   out_draws [S x g x ncomp] column-major = the ncomp matrices grad.s1.mcmc ... (each S x g),
 *   R/spatial_gradient.R:378-390.  out_summary [g x 3 x ncomp] column-major = per component the
 *   (estimate, lower, upper) columns of grad1.est ... computed from nbatch batch medians with a
 *   prob-HPD interval (R/spatial_gradient.R:391-419), unrounded; NULL to skip.
 *   out_mean [ncomp x g x S] and out_var [ncomp x ncomp x g x S] (mean.grad and var.grad,
 *   R/spatial_gradient.R:239-240) are optional (NULL to skip).
 *   ngpu >= 1: samples are sharded contiguously over the first ngpu devices of this process
 *   (R/spatial_gradient.R:29-32,200: mclapply over sample chunks), the shards' draws are gathered
 *   once on device 0 -- one grouped NCCL send/recv over NVLink (peer copies if libnccl cannot be loaded) --
 *   and summarised there. */
int spw_gradient_run(int type, int n, const double* coords, int g, const double* grid,
                     int S, const double* phi, const double* sig2, const double* z, const double* eps,
                     int ngpu, int nbatch, double prob,
                     double* out_draws, double* out_summary, double* out_mean, double* out_var, int* info);

/* Device-resident shard entry point (current device).  Layouts are the library's internal ones:
 *   z_dev [S x ldz] row-major per sample (ldz >= n), eps_dev [S][g][ncomp],
 *   out_draws_dev [S][ncomp][g] (sample-major so that shards of different ranks concatenate),
 *   out_mean_dev [S][g][ncomp] / out_var_dev [S][g][ncomp*ncomp] optional.  info is a host pointer. */
int spw_gradient_shard_dev(int type, int n, const double* coords_dev, int g, const double* grid_dev,
                           int S, const double* phi_dev, const double* sig2_dev,
                           const double* z_dev, int ldz, const double* eps_dev,
                           double* out_draws_dev, double* out_mean_dev, double* out_var_dev,
                           int* info, void* stream);

/* spsample (the producer of model$z between the two hot paths) --------------------
 * spw_spsample_run: the per-sample Gibbs draw of the latent process, R/spsample_beta.R:50-100, without the p x p
 *   regression block (:76-78), which the host evaluates: w [n x S] column-major holds (y - X beta_s)/tau2_s,
 *   eps [n x S] the rnorm(N) of :93.  out_z [S x n] column-major (z.mcmc).  bad[s] != 0 marks a "Bad Iteration"
 *   (:85: chol(Sig.Z.in) failed); the host applies the fallback rule of :86-89 to those rows.  A sample whose
 *   chol(R.Z) fails is redone with R.Z + 1e-4 I (:70-71). */
int spw_spsample_run(int type, int n, const double* coords, int S, const double* phi, const double* sig2,
                     const double* tau2, const double* w, const double* eps, double* out_z, int* bad, int* info);

/* bayes_cwomb(type = "rectilinear", approx = NULL) -- R/spatial_wombling.R:159-244
 * (matern2) and :308-393 (gaussian), kernels Gamma.* / K.* :464-711.  The branch does not run upstream as written
 * (`ntimes <- 1:length(model$phis)`); the semantics implemented are ntimes = S = length(model$phi), everything else
 * literally (see oracle/wombling.py and DESIGN.md).  type: SPW_COV_GAUSSIAN or SPW_COV_MATERN2 (the matern1 kernels
 * divide 0 by 0 upstream).  curve: (nseg + 1) x 2 column-major polyline; rule_len: rule.len; z: S x n column-major;
 * eps [S][nseg][2]: the two rnorm() variates mvrnorm consumes per (sample, segment).
 * out_w [2][S][nseg] = womb1.mcmc, womb2.mcmc (each nseg x S column-major); optional out_mean [S][nseg][2] (mean.womb)
 * and out_var [S][nseg][4] (var.womb, row-major 2 x 2).  SPW_ERR_NOT_PD with info = {sample, minor, 1} when
 * chol(Sig.Z) fails, {sample, segment, 3} when mvrnorm's "'Sigma' is not positive definite" test fails. */
int spw_wombling_rect_run(int type, int n, const double* coords, int nseg, const double* curve, int rule_len, int S,
                          const double* phi, const double* sig2, const double* z, const double* eps, double* out_w,
                          double* out_mean, double* out_var, int* info);

/* spw_hpd_summary: batch medians, median of batch medians and coda::HPDinterval over the batch
 *   medians -- R/spatial_gradient.R:391-419 (:353-358 for matern1).  draws [S x g] column-major,
 *   out [g x 3] column-major (estimate, lower, upper), unrounded. */
int spw_hpd_summary(int S, int g, int nbatch, double prob, const double* draws, double* out);
/* device variant on the internal layout: draws_dev [S][ncomp][g], out_dev [ncomp][3][g] */
int spw_hpd_summary_dev(int S, int g, int ncomp, int nbatch, double prob, const double* draws_dev,
                        double* out_dev, void* stream);
/* [S][ncomp][g] (device) -> R layout [ncomp][g][S] (device): the S x g column-major matrices */
int spw_draws_to_r_layout_dev(int S, int g, int ncomp, const double* draws_dev, double* out_dev, void* stream);

/* Posterior surface analysis (README.md:123-177): from the five S x g draw matrices
 * (draws [S x g x 5] column-major, components s1, s2, s11, s12, s22) compute per (sample, grid point) the Hessian
 * determinant, its two eigenvalues (decreasing), the Laplacian and the divergence (out_derived [S x g x 5], optional)
 * and their batch-median / HPD summaries (out_summary [g x 3 x 5] column-major, optional). */
int spw_surface_analysis(int S, int g, const double* draws, int nbatch, double prob, double* out_summary,
                         double* out_derived);
/* device variant on the internal layout: draws_dev / out_dev [S][5][g] */
int spw_surface_analysis_dev(int S, int g, const double* draws_dev, double* out_dev, void* stream);

/* .C()-compatible variants (every argument a pointer, no R headers needed) ------------------------ */
int spw_cov_p(const int* type, const int* n, const double* coords, const double* phi, const double* sig2,
              double* out, int* status);
int spw_hpd_summary_p(const int* S, const int* g, const int* nbatch, const double* prob,
                      const double* draws, double* out, int* status);
/* hyper = c(a.sigma, b.sigma, a.tau, b.tau, lower.phi, upper.phi) */
int spw_hlm_run_p(const int* n, const double* coords, const double* y, const int* type, const int* niter,
                  const int* nburn, const int* report, const double* hyper, const double* norm, const double* unif,
                  const int* want_trgt, double* out_sig2, double* out_tau2, double* out_phi, double* out_trgt,
                  double* out_accept, double* out_tuning, int* info, int* status);
int spw_gradient_run_p(const int* type, const int* n, const double* coords, const int* g, const double* grid,
                       const int* S, const double* phi, const double* sig2, const double* z, const double* eps,
                       const int* ngpu, const int* nbatch, const double* prob, const int* want_draws,
                       double* out_draws, double* out_summary, int* info, int* status);

#ifdef __cplusplus
}
#endif
#endif /* SPWOMBLING_H */
