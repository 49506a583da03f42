// hlmBayes_sp: the collapsed Metropolis-within-Gibbs chain of R/hlmBayes_sp.R:99-302, resident on the device.
// Per proposal: Sigma' = sig2' R(phi') + tau2' I assembled straight from the coordinates (R is never stored),
// blocked Cholesky of the system augmented with the row y^T (so row n of the factor is (L^-1 y)^T),
// log-determinant and quadratic form by reduction, accept / reject and adaptation in a one-thread kernel.
// The reference's explicit inverse (chol2inv, :100,115,141,182) is not formed: only y' Sigma^-1 y = |L^-1 y|^2
// is consumed (:118,144,184).  One iteration is a fixed launch sequence and is replayed as a CUDA graph.
#include <cstdlib>
#include "hlm.cuh"
#include "cov.cuh"

namespace spw {

struct HlmState {
  double par[3];        // sig2, tau2, phi (current)
  double prop[3];       // parameters the next factorisation is built from
  double lprop;         // log of the proposed parameter
  double e[3];          // proposal scales e.sig2, e.tau2, e.phi        R/hlmBayes_sp.R:50
  double acc[3];        // acceptance counters                          R/hlmBayes_sp.R:76
  double quad, logdet;  // y' Sigma^-1 y and sum(log(diag(chol(Sigma)))) of the current state
  double quad_p, logdet_p;
  int oob;              // phi proposal outside [lower.phi, upper.phi]  R/hlmBayes_sp.R:164
  int iter;             // 0-based iteration counter
  int nrep;
  int fail_iter;        // 1-based iteration whose chol() failed first (0: none); the reference stops there
};

struct HlmConst {
  double a_sigma, b_sigma, a_tau, b_tau, lower_phi, upper_phi;
  int niter, report, want_trgt;
};

__device__ double hlm_target(const HlmState& s, const HlmConst& c) {
  // R/hlmBayes_sp.R:103-105, 235-237 (sic: "- -(a.tau + 1) * log(tau2)" and "/4")
  return -(c.a_sigma + 1.0) * log(s.par[0]) - c.b_sigma / s.par[0] - -(c.a_tau + 1.0) * log(s.par[1]) -
         c.b_tau / s.par[1] - s.logdet / 4.0 - s.quad / 2.0;
}

// p = 0, 1, 2: which parameter is proposed.  p = -1: initial state (no proposal).
__device__ void hlm_propose(HlmState& s, const HlmConst& c, int p, const double* __restrict__ norm) {
  s.prop[0] = s.par[0];
  s.prop[1] = s.par[1];
  s.prop[2] = s.par[2];
  s.oob = 0;
  if (p < 0) return;
  const double l = log(s.par[p]) + s.e[p] * norm[3 * (long long)s.iter + p];   // :112, :138, :162
  s.lprop = l;
  const double v = exp(l);
  if (p == 2 && (v < c.lower_phi || v > c.upper_phi)) {
    s.oob = 1;          // no linear algebra in the reference; here the factorisation of the current state is redone
  } else {
    s.prop[p] = v;
  }
}

struct HlmOut {
  double *sig2, *tau2, *phi, *trgt, *accept, *tuning;
  int nrep;
};

// failed: chol() of the proposal hit a non-positive pivot (the reference's chol() raises an R error at that point,
// R/hlmBayes_sp.R:114/140/181).  The proposal is rejected -- a NaN log-ratio must not be "accepted" through
// fmin(NaN, 0) = 0 -- and the iteration is recorded; the caller returns SPW_ERR_NOT_PD with it.
__device__ void hlm_accept(HlmState& s, const HlmConst& c, int p, const double* __restrict__ unif, const HlmOut& o,
                           bool failed) {
  double* out_sig2 = o.sig2; double* out_tau2 = o.tau2; double* out_phi = o.phi; double* out_trgt = o.trgt;
  double* out_accept = o.accept; double* out_tuning = o.tuning;
  const int nrep = o.nrep;
  if (p < 0) {          // initial factorisation: R/hlmBayes_sp.R:99-106
    s.quad = s.quad_p;
    s.logdet = s.logdet_p;
    if (c.want_trgt) out_trgt[0] = hlm_target(s, c);
    return;
  }
  const long long i = s.iter;
  double r;
  const double l = s.lprop, lcur = log(s.par[p]);
  const double lik = -(s.quad_p - s.quad) / 2.0 - (s.logdet_p - s.logdet);
  if (p == 0) r = -c.b_sigma * (exp(-l) - 1.0 / s.par[0]) - (c.a_sigma + 1.0) * (l - lcur) + lik;   // :117-119
  else if (p == 1) r = -c.b_tau * (exp(-l) - 1.0 / s.par[1]) - (c.a_tau + 1.0) * (l - lcur) + lik;  // :143-145
  else r = -(l - lcur) + lik;                                                                      // :184-185
  double accept_prob = fmin(r, 0.0);
  if (s.oob) accept_prob = -INFINITY;                                                              // :164
  if (failed || !(r == r)) {
    accept_prob = -INFINITY;
    if (s.fail_iter == 0) s.fail_iter = (int)i + 1;
  }
  if (log(unif[3 * i + p]) < accept_prob) {                                                        // :123, :149, :192
    s.par[p] = exp(l);
    s.quad = s.quad_p;
    s.logdet = s.logdet_p;
    s.acc[p] += 1.0;
  }
  if (p == 0) out_sig2[i] = s.par[0];
  if (p == 1) out_tau2[i] = s.par[1];
  if (p == 2) {
    out_phi[i] = s.par[2];
    if (c.want_trgt) out_trgt[i + 1] = hlm_target(s, c);                                           // :234-238
    if ((i + 1) % c.report == 0) {                                                                 // :244
      const int row = (int)((i + 1) / c.report) - 1;
      for (int k = 0; k < 3; ++k) {
        double x = s.acc[k] / (double)c.report;                                                    // :245
        if (row < nrep) {
          out_accept[row + (long long)nrep * k] = x;
          out_tuning[row + (long long)nrep * k] = s.e[k];
        }
        x = fmax(0.17, fmin(x, 0.75));                                                             // :289
        double step = 1.0;
        if (x > 0.50) step = x / 0.50;                                                             // :290
        else if (x < 0.25) step = x / 0.25;                                                        // :291
        s.e[k] *= step;                                                                            // :294-296
        s.acc[k] = 0.0;                                                                            // :299
      }
    }
    s.iter += 1;
  }
}

__global__ void hlm_propose_kernel(HlmState* S, HlmConst c, int p, const double* __restrict__ norm) {
  hlm_propose(*S, c, p, norm);
}

__global__ void hlm_accept_kernel(HlmState* S, HlmConst c, int p, const double* __restrict__ unif, HlmOut o,
                                  const int* __restrict__ status) {
  hlm_accept(*S, c, p, unif, o, *status != 0);
}

__device__ void hlm_init_state(HlmState& s) {
  for (int k = 0; k < 3; ++k) {
    s.par[k] = 1.0;     // phi = sig2 = tau2 = 1   R/hlmBayes_sp.R:67
    s.prop[k] = 1.0;
    s.e[k] = 0.1;       // R/hlmBayes_sp.R:50
    s.acc[k] = 0.0;
  }
  s.lprop = 0.0;
  s.quad = s.logdet = s.quad_p = s.logdet_p = 0.0;
  s.oob = 0;
  s.iter = 0;
  s.fail_iter = 0;
}

__global__ void hlm_init_state_kernel(HlmState* S) { hlm_init_state(*S); }

// ---------------------------------------------------------------------------------------------------------------
// Small systems (n <= HLM_SMALL_MAX): the whole chain in ONE persistent CTA, the matrix resident in shared memory.
// The augmented lower matrix [Sigma'; y^T] is stored as packed 4 x 4 tiles (tile (ti, tk), tk <= ti, at tile index
// ti (ti + 1) / 2 + tk, 18 doubles apart so that 128-bit accesses of neighbouring threads hit different banks); the
// columns are padded to a multiple of 4 with identity and the row y^T starts a tile row of its own, so it is never a
// pivot row.  Each proposal re-assembles the tiles from the coordinates and factors them right-looking, four pivots
// at a time with one barrier per tile column: one thread per tile of the column factors the 4 x 4 pivot tile
// (redundantly, reciprocal square roots) and solves its own tile X = C Ld^-T in place -- final entries of L -- while
// the remaining threads give every tile further right its rank-4 update X_i X_k^T of the previous column.
// y' Sigma^-1 y is the squared norm of the y tile row, sum(log(diag(chol))) the sum of the logs of the pivots.
// ---------------------------------------------------------------------------------------------------------------
constexpr int HLM_SMALL_THREADS = 512;
constexpr int HLM_TS = 18;                              // doubles between consecutive tiles

__device__ __forceinline__ void tile_load(const double* t, double (&a)[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const double2 u = *reinterpret_cast<const double2*>(t + 4 * r);
    const double2 v = *reinterpret_cast<const double2*>(t + 4 * r + 2);
    a[r][0] = u.x; a[r][1] = u.y; a[r][2] = v.x; a[r][3] = v.y;
  }
}
__device__ __forceinline__ void tile_store(double* t, const double (&a)[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    *reinterpret_cast<double2*>(t + 4 * r) = make_double2(a[r][0], a[r][1]);
    *reinterpret_cast<double2*>(t + 4 * r + 2) = make_double2(a[r][2], a[r][3]);
  }
}

template <int TYPE>
__global__ void __launch_bounds__(HLM_SMALL_THREADS, 1)
hlm_small_kernel(int n, const double* __restrict__ cx, const double* __restrict__ cy, const double* __restrict__ y,
                 HlmState* Sg, HlmConst c, const double* __restrict__ norm, const double* __restrict__ unif, HlmOut o,
                 int* __restrict__ status) {
  extern __shared__ __align__(16) double hs[];
  const int NC = (n + 3) / 4;                           // tile columns (pivot groups)
  const int NR = NC + 1;                                // tile rows: NC matrix tile rows + the y tile row
  const int ntiles = NC * (NC + 1) / 2 + NC;            // lower tiles of the matrix + the y tile row
  double* T = hs;                                       // ntiles * HLM_TS
  double* sx = T + (size_t)ntiles * HLM_TS;
  double* sy = sx + 4 * NC;
  double* syv = sy + 4 * NC;
  double* diagv = syv + 4 * NC;                         // pivots l_jj
  double* red = diagv + 4 * NC;                         // 2 x 16 partial sums
  unsigned short* tl = reinterpret_cast<unsigned short*>(red + 32);   // tile list sorted by tile column: (ti << 8) | tk
  int* tstart = reinterpret_cast<int*>(tl + ((ntiles + 3) & ~3));     // first list entry of each tile column (+ end)
  __shared__ HlmState s;
  __shared__ int bad;
  const int tid = threadIdx.x;
  for (int i = tid; i < 4 * NC; i += HLM_SMALL_THREADS) {
    sx[i] = (i < n) ? cx[i] : 0.0;
    sy[i] = (i < n) ? cy[i] : 0.0;
    syv[i] = (i < n) ? y[i] : 0.0;
  }
  if (tid == 0) {
    hlm_init_state(s);
    bad = 0;
    int e = 0;
    for (int tk = 0; tk < NC; ++tk) {
      tstart[tk] = e;
      for (int ti = tk; ti < NR; ++ti) tl[e++] = (unsigned short)((ti << 8) | tk);
    }
    tstart[NC] = e;
  }
  __syncthreads();
  const int total_steps = 1 + 3 * c.niter;              // initial factorisation + 3 proposals per iteration
  if (tid == 0) hlm_propose(s, c, -1, norm);
  __syncthreads();
  for (int step = 0; step < total_steps; ++step) {
    const int p = (step == 0) ? -1 : (step - 1) % 3;
    const double ph = s.prop[2], s2 = s.prop[0], t2 = s.prop[1];
    // ---- assemble all tiles, one element per thread and pass (tiles are stored in list order: column by column) ----
    for (int idx = tid; idx < ntiles * 16; idx += HLM_SMALL_THREADS) {
      const int e = idx >> 4, r = (idx >> 2) & 3, cc = idx & 3;
      const int ti = tl[e] >> 8, tk = tl[e] & 255;
      const int i = 4 * ti + r, k = 4 * tk + cc;
      double v = 0.0;
      if (ti == NC) {                                   // the y tile row: y^T in its first row
        if (r == 0 && k < n) v = syv[k];
      } else if (i >= n || k >= n) {
        v = (i == k) ? 1.0 : 0.0;                       // identity padding
      } else if (i == k) {
        v = s2 + t2;
      } else if (k < i) {
        const double dx = sx[i] - sx[k], dy = sy[i] - sy[k];
        const double d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        v = s2 * cov_rho<TYPE>(ph, d);
      }
      T[(size_t)e * HLM_TS + 4 * r + cc] = v;
    }
    __syncthreads();
    // ---- factorisation: NC tile columns, ONE barrier each.  In step jb the "owner" lanes (one per tile of tile
    //      column jb: warp 0, and warp 1 while more than 32 tile rows remain) first apply the pending update of
    //      column jb - 1 to the pivot tile and to their own tile, then factor / solve (their tiles are final after
    //      this step); meanwhile the worker warps apply the update of column jb - 1 to all tiles right of column jb.
    //      The step time is the owners' dependent FP64 chain, so their path is uniform and the warps that share a
    //      scheduler with an owner warp (warp & 3 == 0, resp. <= 1) stay idle. ----
    int base_prev = 0, base_cur = 0;                      // first tile of tile columns jb - 1 and jb (== tstart[])
    for (int jb = 0; jb < NC; base_prev = base_cur, base_cur += NR - jb, ++jb) {
      const int nown = (NR - jb > 32) ? 2 : 1;
      const int warp = tid >> 5;
      if (warp < nown) {
        const int ti = jb + tid;
        if (ti < NR) {
          const double* colp = T + (size_t)base_cur * HLM_TS;                   // tile (jb + i, jb) at colp + i * TS
          double D[4][4], C[4][4];
          tile_load(colp, D);
          tile_load(colp + (size_t)tid * HLM_TS, C);
          if (jb > 0) {
            const double* prevp = T + (size_t)(base_prev + 1) * HLM_TS;        // tile (jb + i, jb - 1) at prevp + i * TS
            double xk[4][4], xi[4][4];
            tile_load(prevp, xk);
            tile_load(prevp + (size_t)tid * HLM_TS, xi);
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
              for (int cc = 0; cc <= r; ++cc) {
                double acc = D[r][cc];
#pragma unroll
                for (int k = 0; k < 4; ++k) acc = fma(-xk[r][k], xk[cc][k], acc);
                D[r][cc] = acc;
              }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) {
                double acc = C[r][cc];
#pragma unroll
                for (int k = 0; k < 4; ++k) acc = fma(-xi[r][k], xk[cc][k], acc);
                C[r][cc] = acc;
              }
          }
          const double q0 = D[0][0], rs0 = fast_rsqrt(q0);
          const double l10 = D[1][0] * rs0, l20 = D[2][0] * rs0, l30 = D[3][0] * rs0;
          const double q1 = fma(-l10, l10, D[1][1]), rs1 = fast_rsqrt(q1);
          const double l21 = fma(-l20, l10, D[2][1]) * rs1, l31 = fma(-l30, l10, D[3][1]) * rs1;
          const double q2 = fma(-l21, l21, fma(-l20, l20, D[2][2])), rs2 = fast_rsqrt(q2);
          const double l32 = fma(-l31, l21, fma(-l30, l20, D[3][2])) * rs2;
          const double q3 = fma(-l32, l32, fma(-l31, l31, fma(-l30, l30, D[3][3]))), rs3 = fast_rsqrt(q3);
#pragma unroll
          for (int r = 0; r < 4; ++r) {        // X = C Ld^-T (idle work on the pivot tile's own lane)
            const double x0 = C[r][0] * rs0;
            const double x1 = fma(-x0, l10, C[r][1]) * rs1;
            const double x2 = fma(-x1, l21, fma(-x0, l20, C[r][2])) * rs2;
            const double x3 = fma(-x2, l32, fma(-x1, l31, fma(-x0, l30, C[r][3]))) * rs3;
            C[r][0] = x0; C[r][1] = x1; C[r][2] = x2; C[r][3] = x3;
          }
          if (tid == 0) {
            int first_bad = 0;
            if (!(q3 > 0.0)) first_bad = 4;
            if (!(q2 > 0.0)) first_bad = 3;
            if (!(q1 > 0.0)) first_bad = 2;
            if (!(q0 > 0.0)) first_bad = 1;
            if (first_bad && bad == 0) bad = 4 * jb + first_bad;
            *reinterpret_cast<double2*>(&diagv[4 * jb]) = make_double2(q0 * rs0, q1 * rs1);
            *reinterpret_cast<double2*>(&diagv[4 * jb + 2]) = make_double2(q2 * rs2, q3 * rs3);
          } else {
            tile_store(T + (size_t)(base_cur + tid) * HLM_TS, C);
          }
        }
      } else if ((warp & 3) >= nown && jb > 0) {
        // workers: rank-4 update with column jb - 1 of every tile (ti, tk), tk > jb
        const int nwork = (4 - nown) * 4 * 32;
        const int widx = ((warp >> 2) * (4 - nown) + (warp & 3) - nown) * 32 + (tid & 31);
        const double* prevp = T + (size_t)(base_prev - (jb - 1)) * HLM_TS;     // tile (i, jb - 1) at prevp + i * TS
        for (int e = base_cur + NR - jb + widx; e < ntiles; e += nwork) {
          const int ti = tl[e] >> 8, tk = tl[e] & 255;
          double xi[4][4], xk[4][4], a[4][4];
          tile_load(prevp + (size_t)ti * HLM_TS, xi);
          tile_load(prevp + (size_t)tk * HLM_TS, xk);
          double* tp = T + (size_t)e * HLM_TS;
          tile_load(tp, a);
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              double acc = a[r][cc];
#pragma unroll
              for (int k = 0; k < 4; ++k) acc = fma(-xi[r][k], xk[cc][k], acc);
              a[r][cc] = acc;
            }
          tile_store(tp, a);
        }
      }
      __syncthreads();
    }
    // ---- sum_j log l_jj and |L[y][.]|^2 ----
    double sl = 0.0, sq = 0.0;
    for (int j = tid; j < n; j += HLM_SMALL_THREADS) {
      sl += log(diagv[j]);
      const double w = T[(size_t)(tstart[j >> 2] + NC - (j >> 2)) * HLM_TS + (j & 3)];   // row 0 of the y tile row
      sq = fma(w, w, sq);
    }
    sl = warp_sum(sl);
    sq = warp_sum(sq);
    if ((tid & 31) == 0) {
      red[tid >> 5] = sl;
      red[16 + (tid >> 5)] = sq;
    }
    __syncthreads();
    if (tid == 0) {
      double al = 0.0, aq = 0.0;
      for (int w = 0; w < HLM_SMALL_THREADS / 32; ++w) {
        al += red[w];
        aq += red[16 + w];
      }
      s.logdet_p = al;
      s.quad_p = aq;
      hlm_accept(s, c, p, unif, o, bad != 0);
      if (step + 1 < total_steps) hlm_propose(s, c, step % 3, norm);      // the next proposal, in the same serial section
    }
    __syncthreads();
    if (bad) break;                        // the reference stops in chol() here
  }
  if (tid == 0) {
    *Sg = s;
    if (bad) atomicCAS(status, 0, bad);
  }
}

static size_t hlm_small_smem(int n) {
  const size_t NC = (n + 3) / 4, ntiles = NC * (NC + 1) / 2 + NC;
  return (ntiles * HLM_TS + 4 * 4 * NC + 32) * sizeof(double) + ((ntiles + 3) & ~(size_t)3) * sizeof(unsigned short) +
         (NC + 2) * sizeof(int) + 64;
}

template <int TYPE>
static int hlm_small_launch(int n, const double* cx, const double* cy, const double* y, HlmState* S, const HlmConst& c,
                            const double* norm, const double* unif, const HlmOut& ho, int* status, cudaStream_t st) {
  const size_t smem = hlm_small_smem(n);
  SPW_CUDA(cudaFuncSetAttribute(hlm_small_kernel<TYPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  hlm_small_kernel<TYPE><<<1, HLM_SMALL_THREADS, smem, st>>>(n, cx, cy, y, S, c, norm, unif, ho, status);
  SPW_LAUNCHED();
  return SPW_OK;
}

// status word and the iteration of the first failed chol() back to the host
static int hlm_finish(HlmWorkspace& ws, HlmState* S, int* info_host, cudaStream_t st) {
  int status = 0;
  HlmState hs;
  SPW_CUDA(cudaMemcpyAsync(&status, ws.status, sizeof(int), cudaMemcpyDeviceToHost, st));
  SPW_CUDA(cudaMemcpyAsync(&hs, S, sizeof(HlmState), cudaMemcpyDeviceToHost, st));
  SPW_CUDA(cudaStreamSynchronize(st));
  if (info_host) {
    info_host[0] = hs.fail_iter;          // 1-based iteration (0: the initial factorisation, or no failure)
    info_host[1] = status;
  }
  if (status == LA_STALLED) {
    set_error("hlmBayes_sp: iteration %d: the kernels of the look-ahead Cholesky did not run at the same time (a profiler "
              "that replays kernels one by one, or a time-sliced GPU); set SPW_LA=0", hs.fail_iter);
    return SPW_ERR_CUDA;
  }
  if (status != 0) {
    set_error("hlmBayes_sp: iteration %d: the leading minor of order %d is not positive (chol of sig2*R + tau2*I)",
              hs.fail_iter, status);
    return SPW_ERR_NOT_PD;
  }
  if (hs.fail_iter != 0) {                // no failed pivot, but a log-ratio that is not a number (NaN / Inf in y or coords):
    set_error("hlmBayes_sp: iteration %d: the acceptance ratio is not a number (R: 'missing value where TRUE/FALSE "
              "needed')", hs.fail_iter);  // the reference's `if(log(runif(1)) < accept.prob)` stops there
    return SPW_ERR_NOT_PD;
  }
  return SPW_OK;
}

int hlm_run_dev(HlmWorkspace& ws, int type, int n, const double* cx, const double* cy, const double* y_dev,
                int niter, int report, double a_sigma, double b_sigma, double a_tau, double b_tau,
                double lower_phi, double upper_phi, const double* norm_dev, const double* unif_dev, int want_trgt,
                double* out_sig2, double* out_tau2, double* out_phi, double* out_trgt, double* out_accept,
                double* out_tuning, int* info_host, cudaStream_t st) {
  HlmConst c{a_sigma, b_sigma, a_tau, b_tau, lower_phi, upper_phi, niter, report, want_trgt};
  const int nrep = niter / report;
  HlmState* S = ws.state;
  MatView L{ws.L, ws.ld, (long long)ws.ld * (n + 1)};
  const HlmOut ho{out_sig2, out_tau2, out_phi, out_trgt, out_accept, out_tuning, nrep};
  SPW_CUDA(cudaMemsetAsync(ws.status, 0, sizeof(int), st));
  {
    // default for n <= HLM_SMALL_DEFAULT_MAX, where the single-CTA chain beats the blocked multi-kernel path replayed as a
    // CUDA graph (15.2k vs 5.1k it/s at n = 100, 6.8k vs 3.7k at n = 155, 4.0k vs 2.8k at n = 200); SPW_HLM_SMALL=0 / 1
    // forces the choice (1: up to HLM_SMALL_MAX)
    const char* env_small = getenv("SPW_HLM_SMALL");
    const bool allow_small = env_small ? (env_small[0] == '1') : (n <= HLM_SMALL_DEFAULT_MAX);
    if (allow_small && n <= HLM_SMALL_MAX) {
      switch (type) {
        case 0: SPW_TRY(hlm_small_launch<0>(n, cx, cy, y_dev, S, c, norm_dev, unif_dev, ho, ws.status, st)); break;
        case 1: SPW_TRY(hlm_small_launch<1>(n, cx, cy, y_dev, S, c, norm_dev, unif_dev, ho, ws.status, st)); break;
        default: SPW_TRY(hlm_small_launch<2>(n, cx, cy, y_dev, S, c, norm_dev, unif_dev, ho, ws.status, st)); break;
      }
      return hlm_finish(ws, S, info_host, st);
    }
  }
  hlm_init_state_kernel<<<1, 1, 0, st>>>(S);
  SPW_LAUNCHED();

  auto substep = [&](int p) -> int {
    hlm_propose_kernel<<<1, 1, 0, st>>>(S, c, p, norm_dev);
    SPW_LAUNCHED();
    // Sigma' (lower triangle) from the proposed parameters held in device memory
    SPW_TRY(cov_assemble_batched(type, n, cx, cy, &S->prop[2], &S->prop[0], &S->prop[1], 0, L, 1, 1, st));
    SPW_CUDA(cudaMemcpyAsync(ws.L + (long long)n * ws.ld, y_dev, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    if (ws.dag) SPW_TRY(potrf_dag(*ws.dag, L, ws.dinv, ws.status, st));
    else if (ws.lap) SPW_TRY(potrf_lookahead(*ws.lap, *ws.las, L, ws.dinv, ws.status, st));
    else SPW_TRY(potrf_batched(*ws.plan, L, ws.dinv, nullptr, 1, ws.status, st, ws.la));
    SPW_TRY(logdet_quad_batched(n, L, &S->logdet_p, &S->quad_p, 1, st));
    hlm_accept_kernel<<<1, 1, 0, st>>>(S, c, p, unif_dev, ho, ws.status);
    SPW_LAUNCHED();
    return SPW_OK;
  };
  auto iteration = [&]() -> int {
    for (int p = 0; p < 3; ++p) SPW_TRY(substep(p));
    return SPW_OK;
  };

  SPW_TRY(substep(-1));
  const char* env = getenv("SPW_HLM_GRAPH");
  const bool use_graph = !(env && env[0] == '0');
  int done = 0;
  if (niter > 0) {      // first iteration eagerly (also warms up function attributes outside of capture)
    SPW_TRY(iteration());
    done = 1;
  }
  if (use_graph && niter > 2) {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    SPW_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const long long before = g_launches;
    int rc = iteration();
    const long long per_iter = g_launches - before;
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc != SPW_OK) return rc;
    if (ce != cudaSuccess) { set_error("graph capture failed: %s", cudaGetErrorString(ce)); return SPW_ERR_CUDA; }
    if (ws.exec_cache && *ws.exec_cache) {          // same order, type and schedule as the last call: only parameters differ
      cudaGraphExecUpdateResultInfo upd;
      const cudaError_t ue = cudaGraphExecUpdate(*ws.exec_cache, graph, &upd);
      if (ue == cudaSuccess && upd.result == cudaGraphExecUpdateSuccess) {
        exec = *ws.exec_cache;
      } else {
        cudaGetLastError();
        cudaGraphExecDestroy(*ws.exec_cache);
        *ws.exec_cache = nullptr;
      }
    }
    if (!exec) {
      SPW_CUDA(cudaGraphInstantiate(&exec, graph, 0));
      if (ws.exec_cache) *ws.exec_cache = exec;
    }
    g_launches += per_iter * (long long)(niter - done - 1);   // the captured pass was counted once
    for (; done < niter; ++done) SPW_CUDA(cudaGraphLaunch(exec, st));
    SPW_CUDA(cudaStreamSynchronize(st));
    if (!ws.exec_cache) cudaGraphExecDestroy(exec);
    cudaGraphDestroy(graph);
  } else {
    for (; done < niter; ++done) SPW_TRY(iteration());
  }
  return hlm_finish(ws, S, info_host, st);
}

size_t hlm_state_bytes() { return sizeof(HlmState); }

}  // namespace spw
