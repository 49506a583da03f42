// Single-matrix Cholesky as ONE persistent kernel: a tile DAG executed with in-kernel dependency flags
// (see potrf_dag.cu).  Replaces the ~400 stream-ordered launches of potrf_batched for batch == 1
// (chol() of R/hlmBayes_sp.R:99,114,140,181).
#pragma once
#include <vector>
#include "common.cuh"
#include "linalg.cuh"

namespace spw {

// One tile task of the DAG.  Row blocks are 64 rows (the appended rows of the augmented hlm system form one extra
// block after the last matrix block); a task covers row blocks [rb, rb + nrb), nrb <= 2, of column block cb.
//   type 0: panel      L[rows][cb] = A[rows][cb] * Dinv_cb^T                (in place)
//   type 1: update     A[rows][cb] -= L[rows][k0..k1) * L[cb][k0..k1)^T     (K = 64 (k1 - k0) <= 256)
//   type 2: row task   panel product of the rows against column cb, then the K = 64 updates by column cb of the
//                      near columns k0 .. k1-1 (queued form; expanded into types 0 / 1 by the sub-worker that runs it)
struct DagTask {
  short type, rb, nrb, cb, k0, k1, pad0, pad1;
};

struct DagPlan {
  int n = 0, rows = 0, nblk = 0, nrb = 0;
  int n_urgent = 0, n_bulk = 0;
  int active_cap = 0;             // row tasks that may run at the same time
  DagTask* d_utasks = nullptr;    // row tasks, grouped by source column
  int* d_useg = nullptr;          // nblk + 1 offsets into d_utasks
  DagTask* d_btasks = nullptr;    // bulk tasks (K = 256 updates): one queue per source outer block
  int* d_bseg = nullptr;          // nq + 1 offsets into d_btasks
  int nq = 0;
  int* d_sync = nullptr;          // flags [nrb][nblk] x 2 (in place / bulk buffer) | claim counters of the urgent segments [nblk] and of the bulk
                                  // queues [nq] | chain position | trace counter
  size_t sync_ints = 0;
  double* d_bulk = nullptr;       // accumulator of the K = 256 updates (same layout as the matrix; merged once per block)
  long long ld = 0;
  unsigned long long* d_trace = nullptr;   // optional (SPW_DAG_TRACE=1): 4 words per executed task
  int trace_cap = 0;
  int grid = 0;                   // CTAs of the persistent kernel (== SM count)
};

int dag_plan_build(DagPlan& plan, int n, int extra_rows);
void dag_plan_free(DagPlan& plan);
// L: one matrix (lower triangle, row-major; rows n .. n+extra_rows-1 appended); dinv: [nblk][64*64] scratch.
int potrf_dag(const DagPlan& plan, MatView L, double* dinv, int* status, cudaStream_t st);
// copies the trace of the last run to the host (4 words per task: packed id, start ns, end ns, SM id)
int dag_trace_read(const DagPlan& plan, std::vector<unsigned long long>& out);

}  // namespace spw
