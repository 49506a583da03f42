// Summary / layout kernels (see summary.cu).
#pragma once
#include <algorithm>
#include <vector>
#include "common.cuh"
#include "../../include/spwombling.h"

namespace spw {

std::vector<int> batch_starts(int S, int nbatch);
int hpd_gap(int nb, double prob);

// draws_dev [S][ncomp][G] -> out_dev [ncomp][3][G] (estimate, lower, upper)
// scratch: device buffer of at least hpd_scratch_bytes(...) bytes
size_t hpd_scratch_bytes(int S, int G, int ncomp, int nbatch);
int hpd_summary_dev(int S, int G, int ncomp, int nbatch, double prob, const double* draws_dev, double* out_dev,
                    void* scratch, cudaStream_t st);

// draws_dev [S][5][G] -> out_dev [S][5][G]: det, eigenvalue 1, eigenvalue 2, Laplacian, divergence (README.md:123-137)
int surface_analysis_dev(int S, int G, const double* draws_dev, double* out_dev, cudaStream_t st);

// batch of row-major [rows][cols] -> [cols][rows]
int transpose_batched(int rows, int cols, const double* in, long long in_ld, long long in_bstride, double* out,
                      long long out_ld, long long out_bstride, int batch, cudaStream_t st);

}  // namespace spw
