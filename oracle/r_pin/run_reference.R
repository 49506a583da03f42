# Pin the CPU oracle (and through it the CUDA path) to the UNMODIFIED reference package.
#
#   Rscript oracle/r_pin/run_reference.R /path/to/spWombling [repo root]
#
# Needs R (>= 3.6) with the package's own imports installed (coda, magic, MASS, parallel); it does NOT install or
# load spWombling -- it source()s the reference's R files as they are into a private environment in which
#   * rnorm / runif replay the committed variates of tests/golden/r_pin/in/*.csv in the order the code consumes them
#     (hlmBayes_sp: rnorm(1), runif(1) per update, R/hlmBayes_sp.R:112,123,138,149,162,192;
#      spatial_gradient: rnorm(ncomp) per (sample, grid point), R/spatial_gradient.R:242,278,333;
#      spsample: rnorm(ncol(X)) then rnorm(N) per sample, R/spsample_beta.R:78,93),
#   * mclapply is lapply and detectCores() returns 2 (one worker: the variates are then consumed in sample order),
#   * the free variables niter / nburn / samples that spatial_gradient reads from its caller's environment
#     (R/spatial_gradient.R:24-26) are defined as S, 0, 1:S.
# Nothing else is changed.  Outputs go to tests/golden/r_pin/out/*.csv (17 significant digits);
# tests/test_r_pin.py compares the oracle (CPU) and the CUDA path (-m gpu) with them when they exist and reports
# "parity unpinned" when they do not.  This image has no R, so the outputs are NOT committed yet (DESIGN.md section 2).

args <- commandArgs(trailingOnly = TRUE)
if (length(args) < 1) stop("usage: Rscript run_reference.R /path/to/spWombling [repo root]")
ref <- normalizePath(args[1])
root <- if (length(args) >= 2) normalizePath(args[2]) else normalizePath(".")
indir <- file.path(root, "tests", "golden", "r_pin", "in")
outdir <- file.path(root, "tests", "golden", "r_pin", "out")
dir.create(outdir, recursive = TRUE, showWarnings = FALSE)

rd <- function(case, key) as.matrix(read.csv(file.path(indir, paste0(case, ".", key, ".csv")), header = FALSE))
rv <- function(case, key) as.numeric(rd(case, key))
wr <- function(x, case, key) {
  x <- as.matrix(x)
  s <- matrix(sprintf("%.17g", x), nrow = nrow(x), ncol = ncol(x))
  write(t(s), file.path(outdir, paste0(case, ".", key, ".csv")), ncolumns = ncol(x), sep = ",")
}

replay <- function(values, what) {
  pos <- 0
  function(n, ...) {
    if (n < 1) return(numeric(0))
    idx <- pos + seq_len(n)
    if (idx[n] > length(values)) stop(paste("replay of", what, "exhausted"))
    pos <<- idx[n]
    values[idx]
  }
}

# a fresh environment per run: the reference's files, sourced unmodified, see the overrides defined in it
fresh_env <- function(norm = numeric(0), unif = numeric(0), S = NULL) {
  e <- new.env(parent = globalenv())
  for (pkg in c("coda", "magic", "MASS", "parallel")) suppressPackageStartupMessages(library(pkg, character.only = TRUE))
  e$rnorm <- replay(norm, "rnorm")
  e$runif <- replay(unif, "runif")
  e$mclapply <- function(X, FUN, ..., mc.cores = 1) lapply(X, FUN, ...)
  e$detectCores <- function(...) 2
  if (!is.null(S)) {
    e$niter <- S
    e$nburn <- 0
    e$samples <- 1:S
  }
  for (f in c("cov_gaussian.R", "cov_matern1.R", "cov_matern2.R", "matern.R", "hlmBayes_sp.R", "spatial_gradient.R",
              "spsample_beta.R"))
    sys.source(file.path(ref, "R", f), envir = e)
  e
}

covs <- c("gaussian", "matern1", "matern2")

# ---- covariance kernels (R/cov_*.R:10-17) ----
e <- fresh_env()
delta <- rv("cov", "delta")
par <- rv("cov", "par")
wr(cbind(e$cov_gaussian(Delta = delta, phi = par[1], sig2 = par[2]),
         e$cov_matern1(Delta = delta, phi = par[1], sig2 = par[2]),
         e$cov_matern2(Delta = delta, phi = par[1], sig2 = par[2])), "cov", "out")

# ---- hlmBayes_sp (R/hlmBayes_sp.R:27-339) ----
for (cv in covs) {
  case <- paste0("hlm_", cv)
  par <- rv(case, "par")
  e <- fresh_env(norm = rv(case, "norm"), unif = rv(case, "unif"))
  res <- e$hlmBayes_sp(y = rv(case, "y"), coords = rd(case, "coords"), niter = par[1], nburn = par[2], report = par[3],
                       cov.type = cv, verbose = FALSE, trgtFn.compute = TRUE)
  wr(cbind(res$sig2, res$tau2, res$phi), case, "chain")            # post-burn-in slices (nburn+1):niter
  wr(res$trgt_fn, case, "trgt_fn")                                  # length niter + 1
}

# ---- spatial_gradient, non-Windows branch (R/spatial_gradient.R:198-442) ----
if (Sys.info()[["sysname"]] == "Windows") stop("run this on Linux / macOS: the Windows branch computes different formulas")
for (cv in covs) {
  case <- paste0("grad_", cv)
  phi <- rv(case, "phi")
  S <- length(phi)
  eps <- rd(case, "eps")                                            # (S * G) x ncomp, row = one rnorm(ncomp) call
  e <- fresh_env(norm = as.numeric(t(eps)), S = S)
  model <- list(phi = phi, sig2 = rv(case, "sig2"), z = rd(case, "z"))
  res <- e$spatial_gradient(coords = rd(case, "coords"), model = model, cov.type = cv, grid.points = rd(case, "grid"),
                            nbatch = rv(case, "par")[1], return.mcmc = TRUE)
  for (nm in names(res)) wr(res[[nm]], case, nm)                    # grad*.est (rounded) and grad.s*.mcmc (S x G)
}

# ---- spsample, the X-taking definition that wins at load time (R/spsample_beta.R:19-113) ----
case <- "spsample_matern2"
eb <- rd(case, "eps_beta")
ez <- rd(case, "eps_z")
norm <- as.numeric(unlist(lapply(seq_len(nrow(eb)), function(i) c(eb[i, ], ez[i, ]))))
e <- fresh_env(norm = norm)
res <- e$spsample(y = rv(case, "y"), X = rd(case, "X"), coords = rd(case, "coords"), phi = rv(case, "phi"),
                  sig2 = rv(case, "sig2"), tau2 = rv(case, "tau2"), cov.type = "matern2")
wr(res$z, case, "z")
wr(res$beta0, case, "beta0")
wr(matrix(res$beta, nrow = nrow(eb), byrow = TRUE), case, "beta")

writeLines(c(paste("reference:", ref), paste("R:", R.version.string), paste("date:", Sys.time())),
           file.path(outdir, "PROVENANCE.txt"))
cat("wrote", length(list.files(outdir)), "files to", outdir, "\n")
