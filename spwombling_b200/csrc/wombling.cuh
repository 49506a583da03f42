// Line-integral ("rectilinear") wombling measures (see wombling.cu).
#pragma once
#include "common.cuh"
#include "linalg.cuh"

namespace spw {

// at[b][(seg * 2 + a)][k]: the quadrature-integrated cross-covariance rows gamma.cov of every segment (type 0 / 2)
int womb_gamma_batched(int type, int n, const double* cx, const double* cy, const double* curve, int nseg, int len,
                       const double* phi, MatView at, int batch, cudaStream_t st);
// gram [S][nseg][5] (from the fused kernel) -> draws out_w [2][S][nseg], optional mean [S][nseg][2] / var [S][nseg][4]
// (row-major 2 x 2); status[b] = 1-based segment whose sig2 * var failed mvrnorm's positive-definiteness test
int womb_finish_batched(int type, int nb, int s_begin, int S, int nseg, int len, const double* curve, const double* phi,
                        const double* sig2, const double* gram, const double* eps, double* out_w, double* out_mean,
                        double* out_var, int* status, cudaStream_t st);

}  // namespace spw
