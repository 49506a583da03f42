// bayes_cwomb(type = "rectilinear") -- the line-integral wombling measures of R/spatial_wombling.R:159-244 (matern2)
// and :308-393 (gaussian), kernels Gamma.* / K.* (:464-711).  See oracle/wombling.py for the semantics chosen where the
// upstream code does not run as written (ntimes = S) and for what is reproduced on purpose (sig2 inside the covariance
// AND in the draw, the left Riemann rule with weight tval / rule.len, the dropped last node pair, k.g.21 = -k.g.12,
// mvrnorm's eigen construction on the lower triangle).
//
// Per posterior sample the N x N work is the spatial_gradient machinery: K = cov(phi, sig2 = 1), Linv = chol(K)^-1,
// v = Linv z, then for every curve segment the two quadrature-integrated cross-covariance rows (womb_gamma_kernel)
// go through the fused TRMM + Gram kernel (two components per "grid point" = segment), which returns
// M = G K^-1 G^T and G K^-1 z.  womb_finish_kernel adds the double quadrature k.gamma, the signs and 1 / sig2 factors
// of the reference's sandwich, and the eigen-based draw.
#include "wombling.cuh"

namespace spw {

namespace {

// (Gamma.t.1, Gamma.t.2) at delta = -t u + (sj - s0); Gamma.1 = -Gamma.t.1 and Gamma.2 = Gamma.t.2 at the same node
template <int TYPE>
__device__ __forceinline__ void gamma_kernels(double d1, double d2, double up1, double up2, double phi, double& g1,
                                              double& g2) {
  double k11, k22, k12;
  if constexpr (TYPE == 0) {                                  // R/spatial_wombling.R:464-487
    const double t0 = -2.0 * phi;
    const double e = exp(-(phi * (d1 * d1 + d2 * d2)));
    g1 = t0 * e * d1 * up1 + t0 * e * d2 * up2;
    k11 = t0 * e * (1.0 - 2.0 * phi * (d1 * d1));
    k22 = t0 * e * (1.0 - 2.0 * phi * (d2 * d2));
    k12 = -2.0 * phi * t0 * e * d1 * d2;
  } else {                                                    // :625-647
    const double t0 = -5.0 * (phi * phi);
    const double t1 = 2.2360679774997898 * phi * sqrt(d1 * d1 + d2 * d2);
    const double e = exp(-t1);
    g1 = t0 * e * (1.0 + t1) * d1 / 3.0 * up1 + t0 * e * (1.0 + t1) * d2 / 3.0 * up2;
    k11 = t0 * e * (1.0 + t1 - 5.0 * (phi * phi) * (d1 * d1)) / 3.0;
    k22 = t0 * e * (1.0 + t1 - 5.0 * (phi * phi) * (d2 * d2)) / 3.0;
    k12 = -5.0 * (phi * phi) * t0 * e * d1 * d2 / 3.0;
  }
  g2 = k11 * (up1 * up1) + 2.0 * k12 * up1 * up2 + k22 * (up2 * up2);
}

// (K.11, K.12, K.22) at delta = (t2 - t1) u     R/spatial_wombling.R:515-575 (gaussian), :676-711 (matern2)
template <int TYPE>
__device__ void k_terms(double dt, double u1, double u2, double up1, double up2, double phi, double& K11, double& K12,
                        double& K22) {
  const double d1 = dt * u1, d2 = dt * u2;
  const double cu0 = up1 * up1, cu1 = 2.0 * up1 * up2, cu2 = up2 * up2;
  const double nd = sqrt(d1 * d1 + d2 * d2);
  const bool zero = (d1 == 0.0) && (d2 == 0.0);
  double k11, k22, k12, n4[3][3], n3[2][3];
  if constexpr (TYPE == 0) {
    const double t0 = -2.0 * phi;
    const double e = exp(-(phi * (nd * nd)));
    k11 = t0 * e * (1.0 - 2.0 * phi * (d1 * d1));
    k22 = t0 * e * (1.0 - 2.0 * phi * (d2 * d2));
    k12 = -2.0 * phi * t0 * e * d1 * d2;
    if (zero) {
      K12 = 0.0;
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) n4[a][b] = 0.0;
      n4[0][0] = 4.0 * (phi * phi); n4[1][1] = 4.0 * ((phi * phi) / 3.0); n4[2][2] = 4.0 * (phi * phi);
    } else {
      const double c = 4.0 * (phi * phi) * e;
      const double k111 = c * (3.0 - 2.0 * phi * (d1 * d1)) * d1;
      const double k112 = c * (1.0 - 2.0 * phi * (d1 * d1)) * d2;
      const double k122 = c * (1.0 - 2.0 * phi * (d2 * d2)) * d1;
      const double k222 = c * (3.0 - 2.0 * phi * (d2 * d2)) * d2;
      n3[0][0] = k111; n3[0][1] = k112; n3[0][2] = k122;
      n3[1][0] = k112; n3[1][1] = k122; n3[1][2] = k222;
      const double d14 = (d1 * d1) * (d1 * d1), d24 = (d2 * d2) * (d2 * d2);
      const double k1111 = c * (3.0 - 12.0 * phi * (d1 * d1) + 4.0 * (phi * phi) * d14);
      const double k1212 = c * (1.0 - 2.0 * (d1 * d1)) * (1.0 - 2.0 * (d2 * d2));             // sic: no phi
      const double k2222 = c * (3.0 - 12.0 * phi * (d2 * d2) + 4.0 * (phi * phi) * d24);
      const double p3 = (phi * phi) * phi;
      const double k1112 = -8.0 * p3 * e * (3.0 - 2.0 * phi * (d1 * d1)) * d1 * d2;
      const double k1222 = -8.0 * p3 * e * (3.0 - 2.0 * phi * (d2 * d2)) * d2 * d1;
      n4[0][0] = k1111; n4[0][1] = k1112; n4[0][2] = k1212;
      n4[1][0] = k1212; n4[1][1] = k1212; n4[1][2] = k1222;
      n4[2][0] = k1212; n4[2][1] = k1222; n4[2][2] = k2222;
      const double r0 = up1 * n3[0][0] + up2 * n3[1][0], r1 = up1 * n3[0][1] + up2 * n3[1][1], r2 = up1 * n3[0][2] + up2 * n3[1][2];
      K12 = r0 * cu0 + r1 * cu1 + r2 * cu2;
    }
  } else {
    const double s5 = 2.2360679774997898;
    const double t0 = -5.0 * (phi * phi);
    const double t1 = s5 * phi * nd;
    const double e = exp(-t1);
    const double p2 = phi * phi, p4 = p2 * p2;
    k11 = t0 * e * (1.0 + t1 - 5.0 * p2 * (d1 * d1)) / 3.0;
    k22 = t0 * e * (1.0 + t1 - 5.0 * p2 * (d2 * d2)) / 3.0;
    k12 = -5.0 * p2 * t0 * e * d1 * d2 / 3.0;
    if (zero) {
      K12 = 0.0;
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) n4[a][b] = 0.0;
      n4[0][0] = 25.0 * p4; n4[1][1] = 25.0 * (p4 / 3.0); n4[2][2] = 25.0 * p4;
    } else {
      const double c = 25.0 * p4 * e;
      const double nd3 = (nd * nd) * nd, nd2 = nd * nd;
      const double k111 = c * (3.0 - s5 * phi * (d1 * d1) / nd) * d1 / 3.0;
      const double k112 = c * (1.0 - s5 * phi * (d1 * d1) / nd) * d2 / 3.0;
      const double k122 = c * (1.0 - s5 * phi * (d2 * d2) / nd) * d1 / 3.0;
      const double k222 = c * (3.0 - s5 * phi * (d2 * d2) / nd) * d2 / 3.0;
      n3[0][0] = k111; n3[0][1] = k112; n3[0][2] = k122;
      n3[1][0] = k112; n3[1][1] = k122; n3[1][2] = k222;
      const double d14 = (d1 * d1) * (d1 * d1), d24 = (d2 * d2) * (d2 * d2);
      const double k1111 = c * (3.0 - 6.0 * s5 * phi * (d1 * d1) / nd + s5 * phi * d14 / nd3 + 5.0 * p2 * d14 / nd2) / 3.0;
      const double k1212 = c * ((1.0 - s5 * (d1 * d1) / nd) * (1.0 - s5 * (d2 * d2) / nd) + s5 * phi * (d1 * d1) * (d2 * d2) / nd3) / 3.0;   // sic
      const double k2222 = c * (3.0 - 6.0 * s5 * phi * (d2 * d2) / nd + s5 * phi * d24 / nd3 - 5.0 * p2 * d24 / nd2) / 3.0;                  // sic
      const double c5 = 25.0 * s5 * (p4 * phi) * e;
      const double k1112 = c5 * (d1 / nd3 - (3.0 - s5 * phi * (d1 * d1) / nd) / nd) * (d1 * d1) * d2 / 3.0;
      const double k1222 = c5 * (d2 / nd3 - (3.0 - s5 * phi * (d2 * d2) / nd) / nd) * (d2 * d2) * d1 / 3.0;
      n4[0][0] = k1111; n4[0][1] = k1112; n4[0][2] = k1212;
      n4[1][0] = k1212; n4[1][1] = k1212; n4[1][2] = k1222;
      n4[2][0] = k1212; n4[2][1] = k1222; n4[2][2] = k2222;
      const double r0 = up1 * n3[0][0] + up2 * n3[1][0], r1 = up1 * n3[0][1] + up2 * n3[1][1], r2 = up1 * n3[0][2] + up2 * n3[1][2];
      K12 = r0 * cu0 + r1 * cu1 + r2 * cu2;
    }
  }
  K11 = k11 * cu0 + 2.0 * k12 * up1 * up2 + k22 * cu2;
  const double q0 = cu0 * n4[0][0] + cu1 * n4[1][0] + cu2 * n4[2][0];
  const double q1 = cu0 * n4[0][1] + cu1 * n4[1][1] + cu2 * n4[2][1];
  const double q2 = cu0 * n4[0][2] + cu1 * n4[1][2] + cu2 * n4[2][2];
  K22 = q0 * cu0 + q1 * cu1 + q2 * cu2;
}

// gamma.cov rows of every (sample, segment): at[b][(seg * 2 + a)][k] = sum over the first len-1 nodes of Gamma.t.a
// at data point k, times tval / len                                       R/spatial_wombling.R:180-185 / :329-334
template <int TYPE>
__global__ void __launch_bounds__(128)
womb_gamma_kernel(int n, const double* __restrict__ cx, const double* __restrict__ cy, const double* __restrict__ curve,
                  int nseg, int len, const double* __restrict__ phi, MatView at) {
  const int seg = blockIdx.y, b = blockIdx.z;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  // curve is (nseg + 1) x 2 column-major (R layout)
  const double s01 = curve[seg], s02 = curve[nseg + 1 + seg];
  const double e1 = curve[seg + 1] - s01, e2 = curve[nseg + 1 + seg + 1] - s02;
  const double tval = sqrt(e1 * e1 + e2 * e2);
  const double u1 = e1 / tval, u2 = e2 / tval;
  const double up1 = u2, up2 = -u1;
  const double ph = phi[b];
  const double dj1 = cx[k] - s01, dj2 = cy[k] - s02;
  double a1 = 0.0, a2 = 0.0;
  for (int m = 0; m + 1 < len; ++m) {
    const double t = ((double)m * (1.0 / (double)(len - 1))) * tval;      // seq(0, 1, length.out = len) * tval
    double g1, g2;
    gamma_kernels<TYPE>(-t * u1 + dj1, -t * u2 + dj2, up1, up2, ph, g1, g2);
    a1 += g1;
    a2 += g2;
  }
  const double w = tval / (double)len;
  double* base = at.p + (long long)b * at.stride + (long long)(seg * 2) * at.ld + k;
  base[0] = a1 * w;
  base[at.ld] = a2 * w;
}

// per (sample, segment): k.gamma by the double quadrature, the reference's sandwich signs, mvrnorm's draw
template <int TYPE>
__global__ void __launch_bounds__(128)
womb_finish_kernel(int nb, int s_begin, int S, int nseg, int len, const double* __restrict__ curve,
                   const double* __restrict__ phi, const double* __restrict__ sig2, const double* __restrict__ gram,
                   const double* __restrict__ eps, double* __restrict__ out_w, double* __restrict__ out_mean,
                   double* __restrict__ out_var, int* __restrict__ status) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nb * nseg) return;
  const int b = idx / nseg, seg = idx % nseg;
  const long long sidx = s_begin + b;
  const double s01 = curve[seg], s02 = curve[nseg + 1 + seg];
  const double e1 = curve[seg + 1] - s01, e2 = curve[nseg + 1 + seg + 1] - s02;
  const double tval = sqrt(e1 * e1 + e2 * e2);
  const double u1 = e1 / tval, u2 = e2 / tval;
  const double up1 = u2, up2 = -u1;
  const double ph = phi[b], s2 = sig2[b];
  double a11 = 0.0, a12 = 0.0, a22 = 0.0;
  const int npairs = len * len;
  for (int p = 0; p + 1 < npairs; ++p) {                 // expand.grid: the first variable runs fastest; [-len^2]
    const double t1 = ((double)(p % len) * (1.0 / (double)(len - 1))) * tval;
    const double t2 = ((double)(p / len) * (1.0 / (double)(len - 1))) * tval;
    double K11, K12, K22;
    k_terms<TYPE>(t2 - t1, u1, u2, up1, up2, ph, K11, K12, K22);
    a11 += -K11;
    a12 += K12;
    a22 += K22;
  }
  const double sc = (tval * tval) / (double)(len * len);
  const double kg11 = a11 * sc, kg12 = a12 * sc, kg22 = a22 * sc;
  // gram: M00, M01, M11, (G K1^-1 z)_0, (G K1^-1 z)_1 with K1 = cov(sig2 = 1); the reference's K carries sig2, its left
  // factor is tgamma = (-g1, g2):  tmp' gamma = diag(-1, 1) M / sig2,  mean = diag(-1, 1) G K1^-1 z / sig2
  const double* gm = gram + (sidx * nseg + seg) * 5;
  const double m0 = -gm[3] / s2, m1 = gm[4] / s2;
  const double v00 = kg11 + gm[0] / s2, v01 = kg12 + gm[1] / s2;
  const double v10 = -kg12 - gm[1] / s2, v11 = kg22 - gm[2] / s2;
  if (out_mean) {
    out_mean[(sidx * nseg + seg) * 2] = m0;
    out_mean[(sidx * nseg + seg) * 2 + 1] = m1;
  }
  if (out_var) {
    double* o = out_var + (sidx * nseg + seg) * 4;         // row-major 2 x 2
    o[0] = v00; o[1] = v01; o[2] = v10; o[3] = v11;
  }
  // MASS::mvrnorm(1, mean, sig2 * var): eigen(Sigma, symmetric = TRUE) reads the lower triangle
  const double A = s2 * v00, Bv = s2 * v10, C = s2 * v11;
  const double half = 0.5 * (A - C), mid = 0.5 * (A + C);
  const double r = sqrt(half * half + Bv * Bv);
  const double l1 = mid + r, l2 = mid - r;                 // decreasing order
  if (!(l2 >= -1e-6 * fabs(l1)) || !(l1 == l1)) atomicCAS(&status[b], 0, seg + 1);
  // eigenvector of l1: (Bv, l1 - A) or (l1 - C, Bv), whichever is better conditioned; l2: orthogonal
  double x, y;
  if (Bv == 0.0) {
    if (A >= C) { x = 1.0; y = 0.0; } else { x = 0.0; y = 1.0; }
  } else if (half >= 0.0) {
    x = l1 - C; y = Bv;
  } else {
    x = Bv; y = l1 - A;
  }
  const double nrm = sqrt(x * x + y * y);
  x /= nrm; y /= nrm;
  double x2 = -y, y2 = x;
  if (x < 0.0 || (x == 0.0 && y < 0.0)) { x = -x; y = -y; }          // sign convention: first component >= 0
  if (x2 < 0.0 || (x2 == 0.0 && y2 < 0.0)) { x2 = -x2; y2 = -y2; }
  const double* e = eps + (sidx * nseg + seg) * 2;
  const double c1 = sqrt(fmax(l1, 0.0)) * e[0], c2 = sqrt(fmax(l2, 0.0)) * e[1];
  out_w[sidx * nseg + seg] = m0 + x * c1 + x2 * c2;
  out_w[(long long)S * nseg + sidx * nseg + seg] = m1 + y * c1 + y2 * c2;
}

}  // namespace

int womb_gamma_batched(int type, int n, const double* cx, const double* cy, const double* curve, int nseg, int len,
                       const double* phi, MatView at, int batch, cudaStream_t st) {
  if (batch <= 0 || nseg <= 0) return SPW_OK;
  if (nseg > 65535 || batch > 65535) { set_error("wombling: too many segments / samples per batch"); return SPW_ERR_ARG; }
  dim3 grid(ceil_div(n, 128), nseg, batch);
  if (type == 0) womb_gamma_kernel<0><<<grid, 128, 0, st>>>(n, cx, cy, curve, nseg, len, phi, at);
  else womb_gamma_kernel<2><<<grid, 128, 0, st>>>(n, cx, cy, curve, nseg, len, phi, at);
  SPW_LAUNCHED();
  return SPW_OK;
}

int womb_finish_batched(int type, int nb, int s_begin, int S, int nseg, int len, const double* curve, const double* phi,
                        const double* sig2, const double* gram, const double* eps, double* out_w, double* out_mean,
                        double* out_var, int* status, cudaStream_t st) {
  if (nb <= 0 || nseg <= 0) return SPW_OK;
  const int total = nb * nseg;
  if (type == 0)
    womb_finish_kernel<0><<<ceil_div(total, 128), 128, 0, st>>>(nb, s_begin, S, nseg, len, curve, phi, sig2, gram, eps, out_w,
                                                               out_mean, out_var, status);
  else
    womb_finish_kernel<2><<<ceil_div(total, 128), 128, 0, st>>>(nb, s_begin, S, nseg, len, curve, phi, sig2, gram, eps, out_w,
                                                               out_mean, out_var, status);
  SPW_LAUNCHED();
  return SPW_OK;
}

}  // namespace spw
