// Fused TRMM + Gram + draw kernel of the spatial_gradient path (see gradient_tma.cu).
#pragma once
#include "common.cuh"
#include "linalg.cuh"

namespace spw {

struct GradParams {
  MatView linv;            // [batch][n][ld]: lower triangle = chol(K)^-1 (upper triangle holds the transpose)
  MatView at;              // [batch][g_count * ncomp][ld]: nabla.K.t rows of the grid chunk, k contiguous
  const double* v;         // [batch][ldv]: Linv z
  long long ldv;
  const double* phi;       // [batch]
  const double* sig2;      // [batch]
  const double* eps;       // [S][g_total][ncomp]   (indexed by s_begin + b)
  double* out_draws;       // [S][ncomp][g_total]
  double* out_mean;        // [S][g_total][ncomp] or null
  double* out_var;         // [S][g_total][ncomp*ncomp] (each block column-major) or null
  double* out_gram = nullptr;   // [S][g_total][ncomp(ncomp+1)/2 + ncomp] or null: the raw sums -- upper triangle of
                                // M = A K^-1 A^T row by row, then A K^-1 z (TMA kernel only; used by the wombling path)
  int* status;             // [batch]: first grid point (1-based) whose 5x5 chol failed, 0 if none
  int n;
  int g_begin, g_count, g_total;
  int s_begin;
};

// TMA + mbarrier + warp-specialised kernel (gradient_tma.cu)
int gradient_tma_launch(int type, const GradParams& P, int batch, cudaStream_t st);

}  // namespace spw
