// Batch medians, median of batch medians and HPD interval of the gradient draws; layout transposes.
//   ngrad / slist / batch medians / estimates / HPDinterval   R/spatial_gradient.R:391-419 (:353-358 matern1)
//   coda::HPDinterval.mcmc (coda 0.19): sort; gap = max(1, min(n-1, round(n*prob))); first minimal-width window.
// The kernels for small batches put consecutive grid points on consecutive threads (coalesced over the
// [S][ncomp][G] layout) and select order statistics by rank counting (ties broken by index), which is exact and
// order-independent; long batches and long chains of batch medians are sorted instead (see below).
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
  bool bad = false;                       // a NaN draw: R's median returns NA
  for (int i = 0; i < L; ++i) {
    const double xi = x[i * stride];
    bad |= (xi != xi);
    int rank = 0;
    for (int j = 0; j < L; ++j) {
      const double xj = x[j * stride];
      rank += (xj < xi) || (xj == xi && j < i);
    }
    if (rank == lo) vlo = xi;
    if (rank == hi) vhi = xi;
  }
  med[((long long)a * nb + b) * G + g] = bad ? NAN : ((lo == hi) ? vlo : (vlo + vhi) / 2.0);   // R median: mean of the two middle values
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
  bool bad = false;
  for (int i = 0; i < nb; ++i) bad |= (x[(long long)i * G] != x[(long long)i * G]);
  if (bad) {                              // NaN batch medians propagate (R: NA)
    double* o = out + (long long)a * 3 * G + g;
    o[0] = o[G] = o[2 * (long long)G] = NAN;
    return;
  }
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

// ---------------------------------------------------------------------------------------------------
// Large batches / many batches: order statistics by sorting.  The rank-counting kernels above are O(L^2) per thread,
// fine for the gradient summary (25-50 samples per batch, 100-200 batches) but not for hlm_summary or the wombling
// measure, where every sample is its own batch (nb = S): there each (component, grid point) column is copied into
// a row of a [rows][P] buffer (P = next power of two, padded with +inf) and the rows are sorted by a bitonic network --
// in shared memory up to 16384 elements per row, with global-memory passes for the strides above that.
// NaN draws are replaced by +inf for the sort and flagged per row: the row's results are NaN (R's median / HPDinterval
// return NA).
// ---------------------------------------------------------------------------------------------------
constexpr int SORT_CHUNK = 16384;            // doubles of a row sorted in shared memory by one CTA (128 KB)
constexpr int SORT_THREADS = 1024;

__device__ __forceinline__ void cmp_exchange(double& a, double& b, bool ascending) {
  if ((a > b) == ascending) {
    const double t = a;
    a = b;
    b = t;
  }
}

// steps j = j_hi, j_hi/2, ..., 1 of stage k on every SORT_CHUNK-sized chunk of every row (chunk = min(P, SORT_CHUNK));
// k_lo: first stage to run (2 for a full sort of the chunk), k_hi: last stage
__global__ void __launch_bounds__(SORT_THREADS)
bitonic_smem_kernel(double* __restrict__ buf, int P, int chunk, int k_lo, int k_hi) {
  extern __shared__ double sv[];
  const long long row = blockIdx.y;
  const int c0 = blockIdx.x * chunk;
  double* g = buf + row * P + c0;
  for (int i = threadIdx.x; i < chunk; i += SORT_THREADS) sv[i] = g[i];
  __syncthreads();
  for (int k = k_lo; k <= k_hi; k <<= 1) {
    for (int j = min(k >> 1, chunk >> 1); j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (chunk >> 1); t += SORT_THREADS) {
        const int i = 2 * t - (t & (j - 1));          // index with bit j clear
        const bool asc = (((c0 + i) & k) == 0);
        cmp_exchange(sv[i], sv[i + j], asc);
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < chunk; i += SORT_THREADS) g[i] = sv[i];
}

// one step (k, j) with j >= SORT_CHUNK in global memory
__global__ void __launch_bounds__(256)
bitonic_global_kernel(double* __restrict__ buf, int P, int k, int j) {
  const long long row = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (P >> 1)) return;
  const int i = 2 * t - (t & (j - 1));
  double* g = buf + row * P;
  double a = g[i], b = g[i + j];
  const bool asc = ((i & k) == 0);
  if ((a > b) == asc) {
    g[i] = b;
    g[i + j] = a;
  }
}

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// sort every row of buf [rows][P] ascending (P a power of two)
static int sort_rows_dev(double* buf, long long rows, int P, cudaStream_t st) {
  if (rows <= 0 || P <= 1) return SPW_OK;
  const int chunk = std::min(P, SORT_CHUNK);
  static bool attr_set[64] = {false};
  int dev = 0;
  SPW_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    SPW_CUDA(cudaFuncSetAttribute(bitonic_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(SORT_CHUNK * sizeof(double))));
    attr_set[dev & 63] = true;
  }
  const size_t smem = (size_t)chunk * sizeof(double);
  for (long long r0 = 0; r0 < rows; r0 += 65535) {
    const int nr = (int)std::min<long long>(65535, rows - r0);
    double* b = buf + r0 * P;
    bitonic_smem_kernel<<<dim3(P / chunk, nr), SORT_THREADS, smem, st>>>(b, P, chunk, 2, chunk);
    SPW_LAUNCHED();
    for (int k = 2 * chunk; k <= P; k <<= 1) {
      for (int j = k >> 1; j >= chunk; j >>= 1) {
        bitonic_global_kernel<<<dim3(ceil_div(P >> 1, 256), nr), 256, 0, st>>>(b, P, k, j);
        SPW_LAUNCHED();
      }
      bitonic_smem_kernel<<<dim3(P / chunk, nr), SORT_THREADS, smem, st>>>(b, P, chunk, k, k);
      SPW_LAUNCHED();
    }
  }
  return SPW_OK;
}

// rows of the batch-median sort: row = (a * G + g) * nb + b holds the draws of batch b, padded with +inf
__global__ void __launch_bounds__(256)
fill_batch_rows_kernel(int G, int ncomp, int nb, int P, const int* __restrict__ bstart, const double* __restrict__ draws,
                       double* __restrict__ buf, int* __restrict__ nan_flag) {
  const long long row = blockIdx.x;
  const int b = (int)(row % nb);
  const long long ag = row / nb;
  const int a = (int)(ag / G), g = (int)(ag % G);
  const int s0 = bstart[b], L = bstart[b + 1] - s0;
  bool bad = false;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    double v = INFINITY;
    if (i < L) {
      v = draws[((long long)(s0 + i) * ncomp + a) * G + g];
      if (v != v) { bad = true; v = INFINITY; }
    }
    buf[row * P + i] = v;
  }
  if (bad) atomicOr(&nan_flag[row], 1);
}

__global__ void __launch_bounds__(256)
median_from_rows_kernel(long long rows, int G, int nb, int P, const int* __restrict__ bstart, const double* __restrict__ buf,
                        const int* __restrict__ nan_flag, double* __restrict__ med /* [ncomp][nb][G] */) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const int b = (int)(row % nb);
  const long long ag = row / nb;
  const int a = (int)(ag / G), g = (int)(ag % G);
  const int L = bstart[b + 1] - bstart[b];
  const double* v = buf + row * P;
  const int lo = (L - 1) / 2, hi = L / 2;
  double m = (lo == hi) ? v[lo] : (v[lo] + v[hi]) / 2.0;
  if (nan_flag[row]) m = NAN;
  med[((long long)a * nb + b) * G + g] = m;
}

// rows of the HPD sort: row = a * G + g holds the nb batch medians, padded with +inf
__global__ void __launch_bounds__(256)
fill_hpd_rows_kernel(int G, int nb, int P, const double* __restrict__ med, double* __restrict__ buf,
                     int* __restrict__ nan_flag) {
  const long long row = blockIdx.x;
  const int a = (int)(row / G), g = (int)(row % G);
  bool bad = false;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    double v = INFINITY;
    if (i < nb) {
      v = med[((long long)a * nb + i) * G + g];
      if (v != v) { bad = true; v = INFINITY; }
    }
    buf[row * P + i] = v;
  }
  if (bad) atomicOr(&nan_flag[row], 1);
}

// one CTA per sorted row: estimate = median, HPD = first window of `gap` order statistics with minimal width
__global__ void __launch_bounds__(256)
hpd_rows_kernel(int G, int nb, int P, int gap, const double* __restrict__ buf, const int* __restrict__ nan_flag,
                double* __restrict__ out /* [ncomp][3][G] */) {
  __shared__ double sw[256];
  __shared__ int si[256];
  const long long row = blockIdx.x;
  const int a = (int)(row / G), g = (int)(row % G);
  const double* v = buf + row * P;
  double bw = INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i + gap < nb; i += blockDim.x) {
    const double w = v[i + gap] - v[i];
    if (w < bw || (w == bw && i < bi)) { bw = w; bi = i; }       // first minimum, like which.min
  }
  sw[threadIdx.x] = bw;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const double w2 = sw[threadIdx.x + o];
      const int i2 = si[threadIdx.x + o];
      if (w2 < sw[threadIdx.x] || (w2 == sw[threadIdx.x] && i2 < si[threadIdx.x])) { sw[threadIdx.x] = w2; si[threadIdx.x] = i2; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int lo = (nb - 1) / 2, hi = nb / 2, best = si[0];
    double est = (lo == hi) ? v[lo] : (v[lo] + v[hi]) / 2.0;
    double l = v[best], u = v[best + gap];
    if (nan_flag[row]) est = l = u = NAN;
    double* o = out + (long long)a * 3 * G + g;
    o[0] = est;
    o[G] = l;
    o[2 * (long long)G] = u;
  }
}

// small problems: rank counting (exact, coalesced); anything else: sorting
static bool summary_small(int lmax, int nb) { return lmax <= 64 && nb <= 1024; }

size_t hpd_scratch_bytes(int S, int G, int ncomp, int nbatch) {
  std::vector<int> starts = batch_starts(std::max(S, 1), std::max(nbatch, 1));
  const size_t nb = starts.size() - 1;
  int lmax = 1;
  for (size_t i = 0; i + 1 < starts.size(); ++i) lmax = std::max(lmax, starts[i + 1] - starts[i]);
  size_t bytes = sizeof(int) * (nb + 2) + 256 + 2 * sizeof(double) * (size_t)ncomp * nb * G + 1024;
  if (!summary_small(lmax, (int)nb)) {
    const size_t cols = (size_t)ncomp * G;
    size_t sort_bytes = 0;
    if (lmax > 64) sort_bytes = std::max(sort_bytes, (sizeof(double) * next_pow2(lmax) + sizeof(int)) * cols * nb);
    sort_bytes = std::max(sort_bytes, (sizeof(double) * next_pow2((int)nb) + sizeof(int)) * cols);
    bytes += sort_bytes + 1024;
  }
  return bytes;
}

int hpd_summary_dev(int S, int G, int ncomp, int nbatch, double prob, const double* draws_dev, double* out_dev,
                    void* scratch, cudaStream_t st) {
  if (S < 1 || G < 1 || nbatch < 1) { set_error("hpd_summary: bad sizes"); return SPW_ERR_ARG; }
  std::vector<int> starts = batch_starts(S, nbatch);
  const int nb = (int)starts.size() - 1;
  if (nb < 2) { set_error("hpd_summary: obj must have nsamp > 1 (need at least two batches)"); return SPW_ERR_ARG; }
  int lmax = 1;
  for (int i = 0; i < nb; ++i) lmax = std::max(lmax, starts[i + 1] - starts[i]);
  char* base = (char*)scratch;
  int* d_starts = (int*)base;
  size_t off = (sizeof(int) * starts.size() + 255) / 256 * 256;
  double* d_med = (double*)(base + off);
  double* d_sorted = d_med + (size_t)ncomp * nb * G;
  SPW_CUDA(cudaMemcpyAsync(d_starts, starts.data(), sizeof(int) * starts.size(), cudaMemcpyHostToDevice, st));
  SPW_CUDA(cudaStreamSynchronize(st));   // the host vector must outlive the copy (pageable source)
  const int gap = hpd_gap(nb, prob);
  if (summary_small(lmax, nb)) {
    batch_median_kernel<<<dim3(ceil_div(G, 128), nb, ncomp), 128, 0, st>>>(G, ncomp, nb, d_starts, draws_dev, d_med);
    SPW_LAUNCHED();
    hpd_kernel<<<dim3(ceil_div(G, 128), ncomp), 128, 0, st>>>(G, nb, gap, d_med, d_sorted, out_dev);
    SPW_LAUNCHED();
    return SPW_OK;
  }
  const long long cols = (long long)ncomp * G;
  double* sbuf = d_sorted + (size_t)ncomp * nb * G;       // [rows][P] followed by the per-row NaN flags
  if (lmax <= 64) {
    for (int b0 = 0; b0 < nb; b0 += 65535) {              // grid.y limit of the rank-counting kernel
      const int cnt = std::min(65535, nb - b0);
      batch_median_kernel<<<dim3(ceil_div(G, 128), cnt, ncomp), 128, 0, st>>>(G, ncomp, nb, d_starts + b0, draws_dev,
                                                                               d_med + (size_t)b0 * G);
      SPW_LAUNCHED();
    }
  } else {
    const int P = next_pow2(lmax);
    const long long rows = cols * nb;
    if (rows > 0x7fffffffLL) { set_error("hpd_summary: too many (column, batch) pairs"); return SPW_ERR_ARG; }
    int* flags = (int*)(sbuf + (size_t)rows * P);
    SPW_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * (size_t)rows, st));
    fill_batch_rows_kernel<<<(unsigned)rows, 256, 0, st>>>(G, ncomp, nb, P, d_starts, draws_dev, sbuf, flags);
    SPW_LAUNCHED();
    SPW_TRY(sort_rows_dev(sbuf, rows, P, st));
    median_from_rows_kernel<<<(unsigned)ceil_div((int)std::min<long long>(rows, 0x7fffffff), 256), 256, 0, st>>>(
        rows, G, nb, P, d_starts, sbuf, flags, d_med);
    SPW_LAUNCHED();
  }
  {
    const int P = next_pow2(nb);
    if (cols > 0x7fffffffLL) { set_error("hpd_summary: too many columns"); return SPW_ERR_ARG; }
    int* flags = (int*)(sbuf + (size_t)cols * P);
    SPW_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * (size_t)cols, st));
    fill_hpd_rows_kernel<<<(unsigned)cols, 256, 0, st>>>(G, nb, P, d_med, sbuf, flags);
    SPW_LAUNCHED();
    SPW_TRY(sort_rows_dev(sbuf, cols, P, st));
    hpd_rows_kernel<<<(unsigned)cols, 256, 0, st>>>(G, nb, P, gap, sbuf, flags, out_dev);
    SPW_LAUNCHED();
  }
  return SPW_OK;
}

}  // namespace spw
