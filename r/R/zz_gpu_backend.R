## Drop-in bodies for the hot functions: same formals and return lists as R/hlmBayes_sp.R:27-36,
## R/spatial_gradient.R:16-21, R/spsample_beta.R:19-26, R/hlm_summary.R:19-23 and R/spatial_wombling.R:28-36; the
## arithmetic goes through .Call into libspwombling.  The file name sorts after every file of the reference, so with
## R's default (alphabetical) collation these definitions replace the originals at load time.  (Not runnable in the
## build image -- no R there; see INTEGRATION.md.)

.spw_type <- function(cov.type) {
  if (is.null(cov.type) || !(cov.type %in% c("gaussian", "matern1", "matern2")))
    stop("cov.type must be one of 'gaussian', 'matern1', 'matern2'")
  match(cov.type, c("gaussian", "matern1", "matern2")) - 1L
}

cov_gaussian <- function(Delta = NULL, phi = NULL, sig2 = NULL) .Call(spw_r_cov_delta, 0L, Delta, phi, sig2)
cov_matern1  <- function(Delta = NULL, phi = NULL, sig2 = NULL) .Call(spw_r_cov_delta, 1L, Delta, phi, sig2)
cov_matern2  <- function(Delta = NULL, phi = NULL, sig2 = NULL) .Call(spw_r_cov_delta, 2L, Delta, phi, sig2)

hlmBayes_sp <- function(y = NULL, coords = NULL, niter = NULL, nburn = NULL, report = NULL,
                        a.sigma = NULL, b.sigma = NULL, a.tau = NULL, b.tau = NULL,
                        lower.phi = NULL, upper.phi = NULL, cov.type = NULL,
                        verbose = TRUE, trgtFn.compute = FALSE, digits = 3) {
  if (is.null(niter)) { niter <- 5e3; nburn <- niter / 2; report <- 1e2 }        # R/hlmBayes_sp.R:43-47
  if (is.null(a.sigma)) a.sigma <- 2; if (is.null(a.tau)) a.tau <- 2             # :59-64
  if (is.null(b.sigma)) b.sigma <- 1; if (is.null(b.tau)) b.tau <- 1
  if (is.null(lower.phi)) lower.phi <- 1e-3; if (is.null(upper.phi)) upper.phi <- 30
  ## the reference consumes rnorm(1), runif(1) three times per iteration (:112,123,138,149,162,192):
  ## draw them in exactly that interleaved order so set.seed() reproduces the reference chain
  norm <- unif <- numeric(3 * niter)
  for (k in seq_len(3 * niter)) { norm[k] <- rnorm(1); unif[k] <- runif(1) }
  res <- .Call(spw_r_hlm_run, as.matrix(coords) + 0, as.numeric(y), .spw_type(cov.type), as.integer(niter),
               as.integer(nburn), as.integer(report), c(a.sigma, b.sigma, a.tau, b.tau, lower.phi, upper.phi),
               norm, unif, isTRUE(trgtFn.compute))
  if (verbose) {                                                                     # the report block of :266-281, printed
    bar <- "=+++++++++++++++++++++++++++++++++++++++++++++++++="                      # after the device chain returns
    q3 <- function(x) paste(round(median(x), digits = (digits - 1)), "(", round(quantile(x, probs = 0.025), digits = (digits - 1)),
                            ",", round(quantile(x, probs = 0.975), digits = (digits - 1)), ")")
    for (b in seq_len(nrow(res[[5]]))) {
      i <- b * report
      idx <- if (i > nburn) nburn:i else (i - report + 1):i                           # :269 vs :276
      cat("Iteration::", i, "Acceptance:", res[[5]][b, ], "Tuning:", round(res[[6]][b, ], digits), "\n",
          bar, "\n",
          "Sigma2::", q3(res[[1]][idx]), "\t", "Tau2::", q3(res[[2]][idx]), "\t", "Phi::", q3(res[[3]][idx]), "\n",
          bar, "\n")
    }
  }
  nu <- c(Inf, 1.5, 2.5)[.spw_type(cov.type) + 1L]                                  # :70-72
  keep <- (nburn + 1):niter                                                          # :334-336
  out <- list(y = y, coords = coords)
  if (trgtFn.compute) out$trgt_fn <- res[[4]]                                        # :325
  c(out, list(nu = nu, sig2 = res[[1]][keep], tau2 = res[[2]][keep], phi = res[[3]][keep]))
}

spatial_gradient <- function(coords = NULL, model = NULL, cov.type = c("matern1", "matern2", "gaussian"),
                             grid.points = NULL, nbatch = 200, return.mcmc = TRUE) {
  ty <- .spw_type(cov.type); nc <- if (ty == 1L) 2L else 5L
  S <- length(model$phi); G <- nrow(grid.points)
  ## rnorm(nc) per (sample, grid point), grid point inner -- the order of R/spatial_gradient.R:242,278,333
  eps <- array(rnorm(nc * G * S), c(nc, G, S))
  res <- .Call(spw_r_gradient_run, ty, as.matrix(coords) + 0, as.matrix(grid.points) + 0, as.numeric(model$phi),
               as.numeric(model$sig2), as.matrix(model$z) + 0, eps, getOption("spWombling.ngpu", 1L),
               as.integer(nbatch), isTRUE(return.mcmc))
  comp <- c("s1", "s2", "s11", "s12", "s22")[seq_len(nc)]
  est.names <- c("grad1.est", "grad2.est", "grad11.est", "grad12.est", "grad22.est")[seq_len(nc)]
  out <- list()
  for (k in seq_len(nc)) {
    m <- res[[1]][, , k]; colnames(m) <- c(paste0("grad.", comp[k], ".est"), "lower", "upper")
    out[[est.names[k]]] <- if (nc == 2L) round(m, 2) else data.frame(round(m, 4))   # :358 / :409
  }
  if (return.mcmc) for (k in seq_len(nc)) out[[paste0("grad.", comp[k], ".mcmc")]] <- res[[2]][, , k]
  out
}

spsample <- function(y = NULL, X = NULL, coords = NULL, phi = NULL, sig2 = NULL, tau2 = NULL,
                     cov.type = c("matern1", "matern2", "gaussian"), silent = TRUE) {
  if (is.null(cov.type)) stop("Please provide a covariance kernel!")                 # R/spsample_beta.R:27
  N <- length(y); S <- length(phi); p <- ncol(X)
  XtX <- t(X) %*% X                                                                   # :30
  beta <- matrix(0, S, p); w <- matrix(0, N, S); eps <- matrix(0, N, S)
  for (i in seq_len(S)) {                                                             # p x p block, :76-78, and the variates in
    post_sd_beta <- chol2inv(chol(1e-4 * diag(p) + XtX / tau2[i]))                    # the reference's order: rnorm(p), rnorm(N)
    post_mean_beta <- post_sd_beta %*% crossprod(X, y - mean(y)) / tau2[i]
    beta[i, ] <- as.vector(crossprod(chol(post_sd_beta), rnorm(p)) + post_mean_beta)
    w[, i] <- (y - X %*% beta[i, ]) / tau2[i]                                         # :92
    eps[, i] <- rnorm(N)                                                              # :93
  }
  res <- .Call(spw_r_spsample_run, .spw_type(cov.type), as.matrix(coords) + 0, as.numeric(phi), as.numeric(sig2),
               as.numeric(tau2), w, eps)
  z <- res[[1]]; bad <- res[[2]]; beta0 <- numeric(S); z.last <- NULL
  for (i in seq_len(S)) {                                                             # "Bad Iteration" rule, :85-89 (one chunk)
    if (bad[i] != 0) {
      if (!silent) cat("Bad Iteration::", i, "\n")
      z[i, ] <- if (i == 1) 0 else if (i == 2) z[1, ] else apply(z[1:(i - 1), , drop = FALSE], 2, median)
    } else z.last <- z[i, ]
    beta0[i] <- mean(z.last)                                                          # :99
  }
  list(z = z, beta0 = beta0, beta = as.vector(beta))                                  # :110-112
}

## hlm_summary: medians + HPD intervals, R/hlm_summary.R:19-50; the O(S N) order statistics of the latent z columns go
## through the GPU reducer (one batch per sample == coda::HPDinterval on the raw samples)
.spw_hpd <- function(x) {
  x <- as.matrix(x) + 0
  .Call(spw_r_hpd_summary, x, nrow(x), 0.95)
}
hlm_summary <- function(sig2 = NULL, tau2 = NULL, phi = NULL, beta = NULL, z = NULL) {
  one <- function(x) as.numeric(.spw_hpd(as.numeric(x)))
  if (is.null(nrow(beta))) {
    param_est <- rbind(sig2 = one(sig2), tau2 = one(tau2), phi = one(phi), beta = one(beta))          # :26-30
  } else {
    param_est <- rbind(sig2 = one(sig2), tau2 = one(tau2), phi = one(phi), .spw_hpd(beta))            # :33-37
  }
  colnames(param_est) <- c("median", "lower.hpd", "upper.hpd")
  z_est <- data.frame(.spw_hpd(z))                                                                    # :41
  colnames(z_est) <- c("median", "lower.hpd", "upper.hpd")
  z_est$sig <- ifelse(z_est[, 2] > 0, 1, ifelse(z_est[, 3] < 0, -1, 0))                               # :43-47
  list(param = param_est, z = z_est)
}

## bayes_cwomb: the riemann.sum branches ride on the new spatial_gradient unchanged (the original function body can
## stay); the rectilinear branch (R/spatial_wombling.R:159-244, :308-393) goes to the GPU.  ntimes = length(model$phi)
## (the upstream `model$phis` / `1:ntimes` do not run), matern1 is refused (K.11.m1 is NaN at t1 = t2).
.spw_cwomb_rect <- function(coords, model, cov.type, nbatch, curve, rule.len) {
  if (cov.type == "matern1") stop("bayes_cwomb(type = 'rectilinear'): the matern1 kernels divide 0 by 0 (K.11.m1)")
  S <- length(model$phi); nseg <- nrow(curve) - 1
  eps <- array(rnorm(2 * nseg * S), c(2, nseg, S))                     # mvrnorm: rnorm(2) per (sample, segment), segment inner
  w <- .Call(spw_r_wombling_rect, .spw_type(cov.type), as.matrix(coords) + 0, as.matrix(curve) + 0, as.integer(rule.len),
             as.numeric(model$phi), as.numeric(model$sig2), as.matrix(model$z) + 0, eps)
  tval <- sapply(1:nseg, function(x) sqrt(sum((curve[(x + 1), ] - curve[x, ])^2)))                    # :42-43
  uval <- t(sapply(1:nseg, function(x) (curve[(x + 1), ] - curve[x, ]))) / tval
  slist <- split(1:S, ceiling(seq_along(1:S) / (S / nbatch)))                                         # :212
  inf <- function(wm) {
    w.batch <- do.call(rbind, lapply(slist, function(x) apply(wm[, x, drop = FALSE], 1, median)))
    est <- data.frame(.Call(spw_r_hpd_summary, t(wm) + 0, as.integer(nbatch), 0.95))                 # :214-216 on the GPU
    colnames(est) <- c("median", "lower", "upper")
    est$sig <- ifelse(est[, 2] > 0 & est[, 3] > 0, 1, ifelse(est[, 2] < 0 & est[, 3] < 0, -1, 0))
    tot <- rowSums(w.batch) / sum(tval)
    list(est = est, row = c(sum(est[, "median"]) / sum(tval), median(tot), HPDinterval(as.mcmc(tot))))
  }
  g <- inf(w[, , 1]); cc <- inf(w[, , 2])
  womb.measure.inf <- rbind(g$row, cc$row)
  colnames(womb.measure.inf) <- c("line-segment-level", "batch-median", "batch-lhpd", "batch-uhpd")
  rownames(womb.measure.inf) <- c("Gamma1(C)", "Gamma2(C)")
  list(tval = tval, uval = uval, womb1.mcmc = w[, , 1], womb2.mcmc = w[, , 2], womb.grad.inf = g$est,
       womb.curv.inf = cc$est, womb.measure.inf = womb.measure.inf)
}

.bayes_cwomb_reference <- bayes_cwomb      # the reference body: its riemann.sum branches call the GPU spatial_gradient
bayes_cwomb <- function(coords = NULL, model = NULL, cov.type = c("gaussian", "matern1", "matern2"), nbatch = NULL,
                        curve = NULL, type = c("rectilinear", "riemann.sum"), approx = NULL, verbose = T, rule.len = 10) {
  if (type[1] == "rectilinear" && is.null(approx))
    return(.spw_cwomb_rect(coords, model, cov.type[1], nbatch, curve, rule.len))
  .bayes_cwomb_reference(coords = coords, model = model, cov.type = cov.type, nbatch = nbatch, curve = curve, type = type,
                         approx = approx, verbose = verbose, rule.len = rule.len)
}

## Posterior surface analysis of README.md:123-177 (Hessian determinant, eigenvalues, Laplacian, divergence per
## (sample, grid point), then batch medians + HPD): one call instead of the S * G eigen() loop.
surface_analysis <- function(gradient.est, nbatch = 100) {
  comp <- c("s1", "s2", "s11", "s12", "s22")
  m <- lapply(comp, function(k) gradient.est[[paste0("grad.", k, ".mcmc")]])
  S <- nrow(m[[1]]); G <- ncol(m[[1]])
  draws <- array(unlist(m), c(S, G, 5))
  res <- .Call(spw_r_surface_analysis, draws, as.integer(S), as.integer(G), as.integer(nbatch))
  out <- list()
  for (k in 1:5) {
    est <- data.frame(res[, , k]); colnames(est) <- c("estimate", "lower", "upper")
    est$sig <- ifelse(est[, 2] > 0 & est[, 3] > 0, 1, ifelse(est[, 2] < 0 & est[, 3] < 0, -1, 0))    # README.md:143-147
    out[[c("det", "eigen1", "eigen2", "laplace", "div")[k]]] <- est
  }
  out
}
