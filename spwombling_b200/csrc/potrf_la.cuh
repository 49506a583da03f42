// Single-matrix blocked Cholesky with a two-step look-ahead (see potrf_la.cu): the chain of diagonal blocks runs on
// its own stream and never waits for a panel product; everything else hangs off it through graph edges.
// chol() of R/hlmBayes_sp.R:99,114,140,181 for one matrix at a time.
#pragma once
#include <vector>
#include "common.cuh"
#include "linalg.cuh"

namespace spw {

// status value when a wait on a flag of a concurrently running kernel gave up after two seconds: the kernels of the
// schedule were not running at the same time (a profiler replaying kernels one by one, a time-sliced GPU)
constexpr int LA_STALLED = -7;

struct LaPlan {
  int n = 0, rows = 0, nblk = 0, nouter = 0;
  int ow = OUTER_W;                             // columns per outer block of the bulk updates (SPW_LA_OUTER, in 64-column blocks)
  int horizon = 1 << 20;                        // outer blocks ahead of the chain that receive updates block by block (SPW_LA_HORIZON)
  std::vector<ChunkRange> lls;                  // per TARGET column c: the left-looking product of its recent sources
  struct BulkLaunch { int cb; ChunkRange tasks; };
  std::vector<std::vector<BulkLaunch>> bulk;    // per source outer block m: the updates launched when its panels are final,
                                                // one per target outer block, nearest target first
  GemmTask* d_tasks = nullptr;
  int n_tasks = 0;
  int* d_flags = nullptr;                       // [0]: diagonal steps published; [1 + c]: CTAs of LLS(c) that have finished
  int* d_expect = nullptr;                      // CTAs per LLS(c) launch
  unsigned long long* d_trace = nullptr;        // optional (SPW_LA_TRACE=1): 8 words per step (ns): chain wait begins, step begins, step published, PANEL2(j) first start / last end, LLS(j) first start / last end
};

struct LaStreams {
  static constexpr int NBULK = 4;               // bulk[0..2]: target outer block cb on bulk[cb % 3]; bulk[3]: the nearest target
  cudaStream_t side = nullptr, side2 = nullptr, side2b = nullptr, side3 = nullptr, bulk[NBULK] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> ev_p, ev_g;          // per outer block: its panels are final / the latest bulk update into it
  cudaEvent_t ev_bjoin[NBULK] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr, ev_join2b = nullptr, ev_join3 = nullptr;
  std::vector<cudaEvent_t> ev_s;                // per step: PANEL2(j) has been launched / finished on the side stream
};

int la_plan_build(LaPlan& plan, int n, int extra_rows);
void la_plan_free(LaPlan& plan);
int la_streams_reserve(LaStreams& s, int nblk);
int la_trace_read(const LaPlan& plan, std::vector<unsigned long long>& out);
// L: one matrix (lower triangle, row-major; rows n .. n+extra_rows-1 appended); dinv: [nblk][64*64] scratch.
// status[0] receives the order of the first non-positive leading minor (0 if none), like potrf_batched.
int potrf_lookahead(const LaPlan& plan, LaStreams& s, MatView L, double* dinv, int* status, cudaStream_t st);

}  // namespace spw
