// Covariance / cross-covariance assembly kernels (see cov.cu).
#pragma once
#include <algorithm>
#include "common.cuh"
#include "linalg.cuh"

namespace spw {

// out_b = sig2_b * rho(phi_b, dist(coords)) + nug_b * I for b < batch.  phi / sig2 / nug are device arrays
// read at [b * param_stride] (sig2 == nullptr -> 1, nug == nullptr -> 0).  lower_only skips the tiles
// strictly above the diagonal.
int cov_assemble_batched(int type, int n, const double* cx, const double* cy, const double* phi, const double* sig2,
                         const double* nug, int param_stride, MatView out, int batch, int lower_only, cudaStream_t st);

int cov_delta(int type, long long len, const double* delta, double phi, double sig2, double* out, cudaStream_t st);

// at[b][(gl * ncomp + a)][k] = nabla.K.t[a, k] at grid point g_begin + gl for sample b (phi[b]).
int cross_cov_batched(int type, int n, const double* cx, const double* cy, int g_begin, int g_count,
                      const double* gx, const double* gy, const double* phi, MatView at, int batch, cudaStream_t st);

}  // namespace spw
