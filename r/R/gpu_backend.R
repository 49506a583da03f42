## Drop-in bodies for the two hot functions: same formals and return lists as R/hlmBayes_sp.R:27-36 and
## R/spatial_gradient.R:16-21; the arithmetic goes through .Call into libspwombling.  (Not runnable in the
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
  if (verbose) for (b in seq_len(nrow(res[[5]])))
    cat("Iteration::", b * report, "Acceptance:", res[[5]][b, ], "Tuning:", round(res[[6]][b, ], digits), "\n")
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
