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

__device__ void hlm_accept(HlmState& s, const HlmConst& c, int p, const double* __restrict__ unif, const HlmOut& o) {
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

__global__ void hlm_accept_kernel(HlmState* S, HlmConst c, int p, const double* __restrict__ unif, HlmOut o) {
  hlm_accept(*S, c, p, unif, o);
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
}

__global__ void hlm_init_state_kernel(HlmState* S) { hlm_init_state(*S); }

// ---------------------------------------------------------------------------------------------------------------
// Small systems (n <= HLM_SMALL_MAX): the whole chain in ONE persistent CTA with the matrix in REGISTERS.
// Thread (ta, tb) of a 16 x 16 grid owns the elements (i, k) = (ta + 16 r, tb + 16 c), c <= r, of the augmented
// lower matrix [Sigma'; y^T]; each proposal re-assembles them from the coordinates (no shared-memory matrix at all)
// and factors in LDL^T form, right-looking: per column the owners publish the raw column to shared memory, ONE
// barrier, every thread updates its registers.  Column j stays "raw" (L[i][j] = raw[i][j] / sqrt(d_j)), so
// y' Sigma^-1 y = sum_j raw[n][j]^2 / d_j and sum(log(diag(chol))) = 1/2 sum_j log d_j need no square roots.
// Launch-bound configurations (README N = 100, meuse N = 155) avoid ~27 kernel launches per iteration this way.
// ---------------------------------------------------------------------------------------------------------------
constexpr int HLM_SMALL_THREADS = 256;

__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;\n" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

template <int TYPE, int NT>
__global__ void __launch_bounds__(HLM_SMALL_THREADS, 1)
hlm_small_kernel(int n, const double* __restrict__ cx, const double* __restrict__ cy, const double* __restrict__ y,
                 HlmState* Sg, HlmConst c, const double* __restrict__ norm, const double* __restrict__ unif, HlmOut o,
                 int* __restrict__ status) {
  __shared__ double sx[16 * NT], sy[16 * NT], syv[16 * NT];
  __shared__ double colbuf[2][16 * NT];
  __shared__ double dvec[16 * NT], wvec[16 * NT];      // pivots d_j and raw y-row entries, reduced after the sweep
  __shared__ double red[2][8];
  __shared__ HlmState s;
  __shared__ int bad;
  const int tid = threadIdx.x;
  const int ta = tid >> 4, tb = tid & 15;
  for (int i = tid; i < 16 * NT; i += HLM_SMALL_THREADS) {
    sx[i] = (i < n) ? cx[i] : 0.0;
    sy[i] = (i < n) ? cy[i] : 0.0;
    syv[i] = (i < n) ? y[i] : 0.0;
  }
  if (tid == 0) {
    hlm_init_state(s);
    bad = 0;
  }
  __syncthreads();
  double a[NT][NT];                                    // a[r][cc] for cc <= r (the rest is never touched)
  const int total_steps = 1 + 3 * c.niter;             // initial factorisation + 3 proposals per iteration
  for (int step = 0; step < total_steps; ++step) {
    const int p = (step == 0) ? -1 : (step - 1) % 3;
    if (tid == 0) hlm_propose(s, c, p, norm);
    __syncthreads();
    const double ph = s.prop[2], s2 = s.prop[0], t2 = s.prop[1];
    // ---- assemble this thread's elements of Sigma' = sig2 R(phi) + tau2 I and of the row y^T ----
#pragma unroll
    for (int r = 0; r < NT; ++r)
#pragma unroll
      for (int cc = 0; cc <= r; ++cc) {
        const int i = ta + 16 * r, k = tb + 16 * cc;
        double v = 0.0;
        if (k < n && i <= n && i >= k) {
          if (i == n) {
            v = syv[k];
          } else if (i == k) {
            v = s2 + t2;
          } else {
            const double dx = sx[i] - sx[k], dy = sy[i] - sy[k];
            const double d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            v = s2 * cov_rho<TYPE>(ph, d);
          }
        }
        a[r][cc] = v;
      }
    // ---- right-looking LDL^T, one barrier per column; thread 0 accumulates log d_j and raw[n][j]^2 / d_j ----
    for (int j = 0; j < n; ++j) {
      const int jr = j >> 4, jb = j & 15;
      double* cb = colbuf[j & 1];
      if (tb == jb) {                                  // owners of column j publish it (rows i = ta + 16 r >= j)
#pragma unroll
        for (int r = 0; r < NT; ++r)
#pragma unroll
          for (int cc = 0; cc <= r; ++cc)
            if (cc == jr) cb[ta + 16 * r] = a[r][cc];
      }
      __syncthreads();
      const double dj = cb[j];
      const double inv = fast_rcp(dj);
      if (tid == HLM_SMALL_THREADS - 1) {              // bookkeeping off the critical path (last warp owns few elements)
        if (!(dj > 0.0) && bad == 0) bad = j + 1;
        dvec[j] = dj;
        wvec[j] = cb[n];
      }
      double ci[NT], ck[NT];
#pragma unroll
      for (int r = 0; r < NT; ++r) {
        ci[r] = cb[ta + 16 * r];
        ck[r] = cb[tb + 16 * r] * inv;
      }
#pragma unroll
      for (int r = 0; r < NT; ++r)
#pragma unroll
        for (int cc = 0; cc <= r; ++cc) {
          const int i = ta + 16 * r, k = tb + 16 * cc;
          if (k > j && i >= k) a[r][cc] = fma(-ci[r], ck[cc], a[r][cc]);   // rows beyond n / columns >= n hold zeros
        }
    }
    __syncthreads();
    double sl = 0.0, sq = 0.0;
    for (int j = tid; j < n; j += HLM_SMALL_THREADS) {
      sl += log(dvec[j]);
      sq += wvec[j] * wvec[j] / dvec[j];
    }
    sl = warp_sum(sl);
    sq = warp_sum(sq);
    if ((tid & 31) == 0) {
      red[0][tid >> 5] = sl;
      red[1][tid >> 5] = sq;
    }
    __syncthreads();
    if (tid == 0) {
      double al = 0.0, aq = 0.0;
      for (int w = 0; w < HLM_SMALL_THREADS / 32; ++w) {
        al += red[0][w];
        aq += red[1][w];
      }
      s.logdet_p = 0.5 * al;
      s.quad_p = aq;
      hlm_accept(s, c, p, unif, o);
    }
    __syncthreads();
  }
  if (tid == 0) {
    *Sg = s;
    if (bad) atomicCAS(status, 0, bad);
  }
}

template <int TYPE>
static int hlm_small_launch(int n, const double* cx, const double* cy, const double* y, HlmState* S, const HlmConst& c,
                            const double* norm, const double* unif, const HlmOut& ho, int* status, cudaStream_t st) {
  const int nt = (n + 1 + 15) / 16;
  if (nt <= 4) hlm_small_kernel<TYPE, 4><<<1, HLM_SMALL_THREADS, 0, st>>>(n, cx, cy, y, S, c, norm, unif, ho, status);
  else if (nt <= 7) hlm_small_kernel<TYPE, 7><<<1, HLM_SMALL_THREADS, 0, st>>>(n, cx, cy, y, S, c, norm, unif, ho, status);
  else hlm_small_kernel<TYPE, 11><<<1, HLM_SMALL_THREADS, 0, st>>>(n, cx, cy, y, S, c, norm, unif, ho, status);
  SPW_LAUNCHED();
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
    // opt-in: measured on B200 the single-CTA chain is instruction-issue bound (~70 us per factorisation at n = 100)
    // and no faster than the blocked multi-kernel path replayed as a CUDA graph, so it is not the default
    const char* env_small = getenv("SPW_HLM_SMALL");
    const bool allow_small = (env_small && env_small[0] == '1');
    if (allow_small && n <= HLM_SMALL_MAX) {
      switch (type) {
        case 0: SPW_TRY(hlm_small_launch<0>(n, cx, cy, y_dev, S, c, norm_dev, unif_dev, ho, ws.status, st)); break;
        case 1: SPW_TRY(hlm_small_launch<1>(n, cx, cy, y_dev, S, c, norm_dev, unif_dev, ho, ws.status, st)); break;
        default: SPW_TRY(hlm_small_launch<2>(n, cx, cy, y_dev, S, c, norm_dev, unif_dev, ho, ws.status, st)); break;
      }
      int status = 0;
      SPW_CUDA(cudaMemcpyAsync(&status, ws.status, sizeof(int), cudaMemcpyDeviceToHost, st));
      SPW_CUDA(cudaStreamSynchronize(st));
      if (info_host) {
        info_host[0] = 0;
        info_host[1] = status;
      }
      if (status != 0) {
        set_error("hlmBayes_sp: the leading minor of order %d is not positive (chol of sig2*R + tau2*I)", status);
        return SPW_ERR_NOT_PD;
      }
      return SPW_OK;
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
    SPW_TRY(potrf_batched(*ws.plan, L, ws.dinv, nullptr, 1, ws.status, st, ws.la));
    SPW_TRY(logdet_quad_batched(n, L, &S->logdet_p, &S->quad_p, 1, st));
    hlm_accept_kernel<<<1, 1, 0, st>>>(S, c, p, unif_dev, ho);
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
    SPW_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    g_launches += per_iter * (long long)(niter - done - 1);   // the captured pass was counted once
    for (; done < niter; ++done) SPW_CUDA(cudaGraphLaunch(exec, st));
    SPW_CUDA(cudaStreamSynchronize(st));
    cudaGraphExecDestroy(exec);
    cudaGraphDestroy(graph);
  } else {
    for (; done < niter; ++done) SPW_TRY(iteration());
  }
  int status = 0;
  SPW_CUDA(cudaMemcpyAsync(&status, ws.status, sizeof(int), cudaMemcpyDeviceToHost, st));
  SPW_CUDA(cudaStreamSynchronize(st));
  if (info_host) {
    info_host[0] = 0;
    info_host[1] = status;
  }
  if (status != 0) {
    set_error("hlmBayes_sp: the leading minor of order %d is not positive (chol of sig2*R + tau2*I)", status);
    return SPW_ERR_NOT_PD;
  }
  return SPW_OK;
}

size_t hlm_state_bytes() { return sizeof(HlmState); }

}  // namespace spw
