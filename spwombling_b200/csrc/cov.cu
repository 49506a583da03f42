// HBM-bound assembly kernels: covariance matrices and derivative cross-covariances, computed straight
// from coordinate tiles (the N x N distance matrix and the N x G dist.s0 / delta.s0 prelude of the
// reference are never materialised).
//   cov_*            R/cov_gaussian.R:14, R/cov_matern1.R:14, R/cov_matern2.R:14
//   dist             R/hlmBayes_sp.R:38, R/spatial_gradient.R:23
//   nabla.K.t rows   R/spatial_gradient.R:233-236 (gaussian), :272 (matern1), :324-327 (matern2)
//   dist.s0/delta.s0 R/spatial_gradient.R:45-46
#include "cov.cuh"

namespace spw {

// --- K1: Sigma_b = sig2_b * rho(phi_b, dist) + nug_b * I, 64 x 64 tiles, 16-byte coalesced stores ------
template <int TYPE>
__global__ void __launch_bounds__(256)
cov_assemble_kernel(int n, const double* __restrict__ cx, const double* __restrict__ cy,
                    const double* __restrict__ phi, const double* __restrict__ sig2, const double* __restrict__ nug,
                    int param_stride, MatView out, int lower_only) {
  const int tr = blockIdx.y, tc = blockIdx.x, b = blockIdx.z;
  if (lower_only && tc > tr) return;
  __shared__ double sx_r[64], sy_r[64], sx_c[64], sy_c[64];
  const int tid = threadIdx.x;
  const int r0 = tr * 64, c0 = tc * 64;
  if (tid < 64) {
    const int r = r0 + tid;
    sx_r[tid] = r < n ? cx[r] : 0.0;
    sy_r[tid] = r < n ? cy[r] : 0.0;
  } else if (tid < 128) {
    const int c = c0 + tid - 64;
    sx_c[tid - 64] = c < n ? cx[c] : 0.0;
    sy_c[tid - 64] = c < n ? cy[c] : 0.0;
  }
  __syncthreads();
  const double ph = phi[(long long)b * param_stride];
  const double s2 = sig2 ? sig2[(long long)b * param_stride] : 1.0;
  const double ng = nug ? nug[(long long)b * param_stride] : 0.0;
  const int cp = tid & 31, rg = tid >> 5;
  const int lc = 2 * cp, c = c0 + lc;
  double* op = out.p + (long long)b * out.stride;
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) {
    const int lr = rg + 8 * rr, r = r0 + lr;
    if (r >= n || c >= n) continue;
    double v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const double dx = sx_r[lr] - sx_c[lc + e], dy = sy_r[lr] - sy_c[lc + e];
      // (dx*dx) + (dy*dy), each product rounded: bit-identical to R's dist() and symmetric in (r, c)
      const double d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
      double val = s2 * cov_rho<TYPE>(ph, d);
      if (r == c + e) val = s2 + ng;   // exact zero distance on the diagonal
      v[e] = val;
    }
    double* dst = op + (long long)r * out.ld + c;
    if (c + 1 < n) *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
    else dst[0] = v[0];
  }
}

int cov_assemble_batched(int type, int n, const double* cx, const double* cy, const double* phi, const double* sig2,
                         const double* nug, int param_stride, MatView out, int batch, int lower_only, cudaStream_t st) {
  if (batch <= 0 || n <= 0) return SPW_OK;
  const int nt = ceil_div(n, 64);
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = std::min(65535, batch - b0);
    dim3 grid(nt, nt, nb);
    MatView o = out;
    o.p += (long long)b0 * out.stride;
    const double* p = phi + (long long)b0 * param_stride;
    const double* s = sig2 ? sig2 + (long long)b0 * param_stride : nullptr;
    const double* g = nug ? nug + (long long)b0 * param_stride : nullptr;
    switch (type) {
      case 0: cov_assemble_kernel<0><<<grid, 256, 0, st>>>(n, cx, cy, p, s, g, param_stride, o, lower_only); break;
      case 1: cov_assemble_kernel<1><<<grid, 256, 0, st>>>(n, cx, cy, p, s, g, param_stride, o, lower_only); break;
      case 2: cov_assemble_kernel<2><<<grid, 256, 0, st>>>(n, cx, cy, p, s, g, param_stride, o, lower_only); break;
      default: set_error("bad cov type %d", type); return SPW_ERR_ARG;
    }
    SPW_LAUNCHED();
  }
  return SPW_OK;
}

// --- elementwise cov_*(Delta, phi, sig2) on a caller-supplied distance array -----------------------------
template <int TYPE>
__global__ void __launch_bounds__(256)
cov_delta_kernel(long long len, const double* __restrict__ delta, double phi, double sig2, double* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride)
    out[i] = sig2 * cov_rho<TYPE>(phi, delta[i]);
}

int cov_delta(int type, long long len, const double* delta, double phi, double sig2, double* out, cudaStream_t st) {
  if (len <= 0) return SPW_OK;
  const int grid = (int)std::min<long long>((len + 255) / 256, 148 * 16);
  switch (type) {
    case 0: cov_delta_kernel<0><<<grid, 256, 0, st>>>(len, delta, phi, sig2, out); break;
    case 1: cov_delta_kernel<1><<<grid, 256, 0, st>>>(len, delta, phi, sig2, out); break;
    case 2: cov_delta_kernel<2><<<grid, 256, 0, st>>>(len, delta, phi, sig2, out); break;
    default: set_error("bad cov type %d", type); return SPW_ERR_ARG;
  }
  SPW_LAUNCHED();
  return SPW_OK;
}

// --- K2: rows of nabla.K.t for every (sample, grid point): at[b][(gl*c + a)][k], k contiguous -------------
template <int TYPE>
__global__ void __launch_bounds__(128)
cross_cov_kernel(int n, const double* __restrict__ cx, const double* __restrict__ cy, int g_begin,
                 const double* __restrict__ gx, const double* __restrict__ gy, const double* __restrict__ phi,
                 MatView at) {
  constexpr int C = (TYPE == 1) ? 2 : 5;
  const int gl = blockIdx.y, b = blockIdx.z;
  const int k = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (k >= n) return;
  const double ph = phi[b];
  const double g1 = gx[g_begin + gl], g2 = gy[g_begin + gl];
  double r[C][2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int kk = (k + e < n) ? k + e : k;
    const double d1 = cx[kk] - g1, d2 = cy[kk] - g2;          // delta.s0 = coords - grid (R/spatial_gradient.R:46)
    const double d = sqrt(__dadd_rn(__dmul_rn(d1, d1), __dmul_rn(d2, d2)));   // dist.s0 (R/spatial_gradient.R:45)
    if constexpr (TYPE == 0) {
      const double e0 = exp(-ph * (d * d));
      const double t = 2.0 * ph * e0;
      r[0][e] = t * d1;
      r[1][e] = t * d2;
      r[2][e] = -t * (1.0 - 2.0 * ph * d1 * d1);
      r[3][e] = 4.0 * ph * ph * e0 * d1 * d2;
      r[4][e] = -t * (1.0 - 2.0 * ph * d2 * d2);
    } else if constexpr (TYPE == 1) {
      const double e0 = exp(-ph * 1.7320508075688772 * d);
      const double t = 3.0 * ph * ph * e0;
      r[0][e] = t * d1;
      r[1][e] = t * d2;
    } else {
      const double a = 2.2360679774997898 * ph;
      const double e0 = exp(-a * d);
      const double q = 1.0 + a * d;
      const double p2 = ph * ph;
      r[0][e] = 5.0 * (p2 * q * e0 * d1) / 3.0;
      r[1][e] = 5.0 * (p2 * q * e0 * d2) / 3.0;
      r[2][e] = 5.0 * (-p2 * e0 * (q - 5.0 * p2 * d1 * d1)) / 3.0;
      r[3][e] = 5.0 * (5.0 * p2 * p2 * e0 * d1 * d2) / 3.0;
      r[4][e] = 5.0 * (-p2 * e0 * (q - 5.0 * p2 * d2 * d2)) / 3.0;
    }
  }
  double* base = at.p + (long long)b * at.stride + (long long)gl * C * at.ld + k;
#pragma unroll
  for (int a = 0; a < C; ++a) {
    double* dst = base + (long long)a * at.ld;
    if (k + 1 < n) *reinterpret_cast<double2*>(dst) = make_double2(r[a][0], r[a][1]);
    else dst[0] = r[a][0];
  }
}

int cross_cov_batched(int type, int n, const double* cx, const double* cy, int g_begin, int g_count,
                      const double* gx, const double* gy, const double* phi, MatView at, int batch, cudaStream_t st) {
  if (batch <= 0 || g_count <= 0) return SPW_OK;
  if (g_count > 65535) { set_error("grid chunk too large"); return SPW_ERR_ARG; }
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = std::min(65535, batch - b0);
    dim3 grid(ceil_div(ceil_div(n, 2), 128), g_count, nb);
    MatView o = at;
    o.p += (long long)b0 * at.stride;
    switch (type) {
      case 0: cross_cov_kernel<0><<<grid, 128, 0, st>>>(n, cx, cy, g_begin, gx, gy, phi + b0, o); break;
      case 1: cross_cov_kernel<1><<<grid, 128, 0, st>>>(n, cx, cy, g_begin, gx, gy, phi + b0, o); break;
      case 2: cross_cov_kernel<2><<<grid, 128, 0, st>>>(n, cx, cy, g_begin, gx, gy, phi + b0, o); break;
      default: set_error("bad cov type %d", type); return SPW_ERR_ARG;
    }
    SPW_LAUNCHED();
  }
  return SPW_OK;
}

}  // namespace spw
