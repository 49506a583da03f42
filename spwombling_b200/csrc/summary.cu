// Batch medians, median of batch medians and HPD interval of the gradient draws; layout transposes.
//   ngrad / slist / batch medians / estimates / HPDinterval   R/spatial_gradient.R:391-419 (:353-358 matern1)
//   coda::HPDinterval.mcmc (coda 0.19): sort; gap = max(1, min(n-1, round(n*prob))); first minimal-width window.
// All kernels put consecutive grid points on consecutive threads (coalesced over the [S][ncomp][G] layout) and
// select order statistics by rank counting (ties broken by index), which is exact and order-independent.
#include <cmath>
#include <vector>
#include "summary.cuh"

namespace spw {

// one thread per (batch, comp, grid point): median of draws[s][comp][g], s in [bstart[b], bstart[b+1])
__global__ void __launch_bounds__(128)
batch_median_kernel(int G, int ncomp, int nb, const int* __restrict__ bstart, const double* __restrict__ draws,
                    double* __restrict__ med /* [ncomp][nb][G] */) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y, a = blockIdx.z;
  if (g >= G) return;
  const int s0 = bstart[b], L = bstart[b + 1] - s0;
  const long long stride = (long long)ncomp * G;
  const double* x = draws + ((long long)s0 * ncomp + a) * G + g;
  const int lo = (L - 1) / 2, hi = L / 2;
  double vlo = 0.0, vhi = 0.0;
  for (int i = 0; i < L; ++i) {
    const double xi = x[i * stride];
    int rank = 0;
    for (int j = 0; j < L; ++j) {
      const double xj = x[j * stride];
      rank += (xj < xi) || (xj == xi && j < i);
    }
    if (rank == lo) vlo = xi;
    if (rank == hi) vhi = xi;
  }
  med[((long long)a * nb + b) * G + g] = (lo == hi) ? vlo : (vlo + vhi) / 2.0;   // R median: mean of the two middle values
}

// one thread per (comp, grid point): sort the nb batch medians, estimate = their median, HPD window
__global__ void __launch_bounds__(128)
hpd_kernel(int G, int nb, int gap, const double* __restrict__ med, double* __restrict__ sorted,
           double* __restrict__ out /* [ncomp][3][G] */) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int a = blockIdx.y;
  if (g >= G) return;
  const double* x = med + (long long)a * nb * G + g;
  double* srt = sorted + (long long)a * nb * G + g;
  for (int i = 0; i < nb; ++i) {
    const double xi = x[(long long)i * G];
    int rank = 0;
    for (int j = 0; j < nb; ++j) {
      const double xj = x[(long long)j * G];
      rank += (xj < xi) || (xj == xi && j < i);
    }
    srt[(long long)rank * G] = xi;
  }
  const int lo = (nb - 1) / 2, hi = nb / 2;
  const double est = (lo == hi) ? srt[(long long)lo * G] : (srt[(long long)lo * G] + srt[(long long)hi * G]) / 2.0;
  int best = 0;
  double bw = srt[(long long)gap * G] - srt[0];
  for (int i = 1; i + gap < nb; ++i) {
    const double w = srt[(long long)(i + gap) * G] - srt[(long long)i * G];
    if (w < bw) {   // strict: first minimum, like which.min
      bw = w;
      best = i;
    }
  }
  double* o = out + (long long)a * 3 * G + g;
  o[0] = est;
  o[G] = srt[(long long)best * G];
  o[2 * (long long)G] = srt[(long long)(best + gap) * G];
}

// tiled transpose of a batch of row-major [rows][cols] matrices -> [cols][rows]
__global__ void __launch_bounds__(256)
transpose_kernel(int rows, int cols, const double* __restrict__ in, long long in_ld, long long in_bstride,
                 double* __restrict__ out, long long out_ld, long long out_bstride) {
  __shared__ double tile[32][33];
  const int bz = blockIdx.z;
  const double* ip = in + (long long)bz * in_bstride;
  double* op = out + (long long)bz * out_bstride;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    if (r < rows && c < cols) tile[i][tx] = ip[(long long)r * in_ld + c];
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < rows && c < cols) op[(long long)c * out_ld + r] = tile[tx][i];
  }
}

int transpose_batched(int rows, int cols, const double* in, long long in_ld, long long in_bstride, double* out,
                      long long out_ld, long long out_bstride, int batch, cudaStream_t st) {
  if (rows <= 0 || cols <= 0 || batch <= 0) return SPW_OK;
  const int gx = ceil_div(cols, 32), gy_total = ceil_div(rows, 32);
  for (int y0 = 0; y0 < gy_total; y0 += 65535) {
    const int gy = std::min(65535, gy_total - y0);
    transpose_kernel<<<dim3(gx, gy, batch), 256, 0, st>>>(rows - y0 * 32, cols, in + (long long)y0 * 32 * in_ld, in_ld,
                                                          in_bstride, out + (long long)y0 * 32, out_ld, out_bstride);
    SPW_LAUNCHED();
  }
  return SPW_OK;
}

// Posterior surface analysis of README.md:123-137: per (sample, grid point) the Hessian [[s11, s12], [s12, s22]] of the
// drawn curvature gives det, the two eigenvalues (decreasing, as eigen() returns them), the Laplacian, and the
// divergence s1 + s2.  draws / out: [S][5][G].
__global__ void __launch_bounds__(256)
surface_kernel(long long S, int G, const double* __restrict__ draws, double* __restrict__ out) {
  const long long total = S * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / G;
    const int g = (int)(i % G);
    const double* d = draws + s * 5 * G + g;
    const double s1 = d[0], s2 = d[G], a = d[2 * (long long)G], b = d[3 * (long long)G], c = d[4 * (long long)G];
    const double t = 0.5 * (a + c), h = hypot(0.5 * (a - c), b);
    double* o = out + s * 5 * G + g;
    o[0] = a * c - b * b;            // det(hess.mat)
    o[G] = t + h;                    // eigen(hess.mat)$values[1]
    o[2 * (long long)G] = t - h;     // eigen(hess.mat)$values[2]
    o[3 * (long long)G] = a + c;     // sum(diag(hess.mat))
    o[4 * (long long)G] = s1 + s2;   // divergence
  }
}

int surface_analysis_dev(int S, int G, const double* draws_dev, double* out_dev, cudaStream_t st) {
  if (S <= 0 || G <= 0) return SPW_OK;
  const long long total = (long long)S * G;
  const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 8);
  surface_kernel<<<grid, 256, 0, st>>>(S, G, draws_dev, out_dev);
  SPW_LAUNCHED();
  return SPW_OK;
}

std::vector<int> batch_starts(int S, int nbatch) {
  // split(ngrad, ceiling(seq_along(ngrad) / (length(ngrad)/nbatch)))  -- R/spatial_gradient.R:392, in double arithmetic
  std::vector<int> starts;
  const double per = (double)S / (double)nbatch;
  double prev = -1.0;
  for (int j = 1; j <= S; ++j) {
    const double lab = std::ceil((double)j / per);
    if (lab != prev) {
      starts.push_back(j - 1);
      prev = lab;
    }
  }
  starts.push_back(S);
  return starts;
}

int hpd_gap(int nb, double prob) {
  const int r = (int)std::nearbyint((double)nb * prob);   // R round(): half to even
  return std::max(1, std::min(nb - 1, r));
}

size_t hpd_scratch_bytes(int S, int G, int ncomp, int nbatch) {
  const size_t nb = (size_t)std::min(std::max(nbatch, 1), std::max(S, 1)) + 2;
  return sizeof(int) * (nb + 2) + 256 + 2 * sizeof(double) * (size_t)ncomp * nb * G + 512;
}

int hpd_summary_dev(int S, int G, int ncomp, int nbatch, double prob, const double* draws_dev, double* out_dev,
                    void* scratch, cudaStream_t st) {
  if (S < 1 || G < 1 || nbatch < 1) { set_error("hpd_summary: bad sizes"); return SPW_ERR_ARG; }
  std::vector<int> starts = batch_starts(S, nbatch);
  const int nb = (int)starts.size() - 1;
  if (nb < 2) { set_error("hpd_summary: obj must have nsamp > 1 (need at least two batches)"); return SPW_ERR_ARG; }
  if (nb > 65535) { set_error("hpd_summary: too many batches"); return SPW_ERR_ARG; }
  char* base = (char*)scratch;
  int* d_starts = (int*)base;
  size_t off = (sizeof(int) * starts.size() + 255) / 256 * 256;
  double* d_med = (double*)(base + off);
  double* d_sorted = d_med + (size_t)ncomp * nb * G;
  SPW_CUDA(cudaMemcpyAsync(d_starts, starts.data(), sizeof(int) * starts.size(), cudaMemcpyHostToDevice, st));
  SPW_CUDA(cudaStreamSynchronize(st));   // the host vector must outlive the copy (pageable source)
  batch_median_kernel<<<dim3(ceil_div(G, 128), nb, ncomp), 128, 0, st>>>(G, ncomp, nb, d_starts, draws_dev, d_med);
  SPW_LAUNCHED();
  hpd_kernel<<<dim3(ceil_div(G, 128), ncomp), 128, 0, st>>>(G, nb, hpd_gap(nb, prob), d_med, d_sorted, out_dev);
  SPW_LAUNCHED();
  return SPW_OK;
}

}  // namespace spw
