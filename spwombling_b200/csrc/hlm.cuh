// hlmBayes_sp device chain (see hlm.cu).
#pragma once
#include "common.cuh"
#include "linalg.cuh"
#include "potrf_dag.cuh"
#include "potrf_la.cuh"

namespace spw {

struct HlmState;

struct HlmWorkspace {
  double* L = nullptr;        // (n + 1) x ld work matrix: Sigma' lower triangle, row n = y^T
  long long ld = 0;
  double* dinv = nullptr;     // inverted diagonal blocks
  int* status = nullptr;
  HlmState* state = nullptr;
  const LinalgPlan* plan = nullptr;   // built with extra_rows = 1
  Lookahead* la = nullptr;            // optional lookahead schedule of the Cholesky
  const LaPlan* lap = nullptr;        // look-ahead schedule (potrf_la.cu); overrides plan / la
  LaStreams* las = nullptr;
  const DagPlan* dag = nullptr;       // persistent tile-DAG Cholesky (one kernel per factorisation); overrides plan / la
  cudaGraphExec_t* exec_cache = nullptr;   // optional: the instantiated iteration graph of the previous call on this device,
                                           // updated in place when the new capture has the same topology (~35 ms less per call)
};

size_t hlm_state_bytes();
constexpr int HLM_SMALL_DEFAULT_MAX = 200;   // the single-CTA chain is the default up to this order
constexpr int HLM_SMALL_MAX = 200;   // largest n handled by the single-CTA chain (packed 4 x 4 tiles in shared memory)

// All pointers are device pointers except info_host.  Outputs as documented for spw_hlm_run.
int hlm_run_dev(HlmWorkspace& ws, int type, int n, const double* cx, const double* cy, const double* y_dev,
                int niter, int report, double a_sigma, double b_sigma, double a_tau, double b_tau,
                double lower_phi, double upper_phi, const double* norm_dev, const double* unif_dev, int want_trgt,
                double* out_sig2, double* out_tau2, double* out_phi, double* out_trgt, double* out_accept,
                double* out_tuning, int* info_host, cudaStream_t st);

}  // namespace spw
