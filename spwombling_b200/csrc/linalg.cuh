// Batched blocked fp64 dense linear algebra on row-major lower-triangular storage
// (bit-identical to R's column-major upper factor chol(): L[i][j] row-major == U[j, i] column-major).
#pragma once
#include <vector>
#include "common.cuh"
#include "../../include/spwombling.h"

namespace spw {

constexpr int NB = 64;        // diagonal block size of the blocked Cholesky / triangular inverse
constexpr int OUTER_W = 256;  // outer block width of the Cholesky (K of the big trailing updates)

// One output tile of the task-driven NT GEMM kernel.
struct GemmTask {
  int a_row, a_col;     // origin of the A block; k runs along columns a_col .. a_col + k
  int b_row, b_col;     // origin of the B block (rows = n index)
  int c_row, c_col;     // origin of the C tile
  int ct_row, ct_col;   // origin of the transposed store (row = n index, col = m index)
  int m, n, k;          // extents
  int a_tri, a_lim0, b_tri, b_lim0;   // 0: none, 1: zero for k > lim0 + row, 2: zero for k < lim0 + row (triangular operands)
  int flags;            // GT_*
};
enum : int { GT_ACCUM = 1, GT_STORE = 2, GT_STORE_T = 4 };

struct MatView {        // a batch of row-major matrices
  double* p;
  long long ld;         // doubles
  long long stride;     // doubles between consecutive batch entries
};

// A / B operands of the task-driven GEMM are described by (view, logical rows, logical cols): elements outside read
// as zero.
struct Operand {
  MatView v;
  int rows, cols;
};
// Optional coupling of a launch with kernels that run at the same time (the persistent diagonal chain of
// potrf_la.cu): every CTA starts once *wait_flag >= wait_val and adds 1 to *done_ctr after its stores.
struct GemmSync {
  const int* wait_flag = nullptr;
  int wait_val = 0;
  int* done_ctr = nullptr;
  const int* abort_flag = nullptr; // the waits end at once when *abort_flag == abort_val (another kernel has given up)
  int abort_val = 0;
  const int* late_flag = nullptr;  // the last late_tiles k-tiles of every task, in groups of late_group tiles, are loaded once
  int late_val[2] = {0, 0};        // late_flag[g] >= late_val[g] (g = 0, 1: group): a launch may start on the operands that
  int late_tiles = 0;              // exist and pick up the rest when kernels that are still running have produced them
  int late_group = 2;
  unsigned long long* stamp = nullptr;   // optional trace: stamp[0] = earliest CTA start, stamp[1] = latest CTA end (ns)
};
// Launches tasks[0 .. count) for every batch entry: C (+)= alpha * A_tile * B_tile^T.  shape: 0 = 128 x 128 tiles
// (1 CTA/SM), 1 = 128 x 64 (2 CTA/SM), 3 = 32 x 64 (3 CTA/SM); the task list must have been tiled accordingly.
// one_per_sm pads the shared memory so that a single CTA fits per SM (room for concurrent latency-critical kernels).
int launch_gemm_tasks(int shape, const GemmTask* tasks, int count, int batch, Operand A, Operand B, MatView C,
                      MatView CT, double alpha, cudaStream_t st, bool one_per_sm = false, const GemmSync* sync = nullptr);

// Loads the kernel of a tile shape (a first-use module load synchronises the device: callers that keep a kernel
// resident while they launch others call this first).
int gemm_prepare(int shape);

// Schedule for one matrix order n: per-step slices into one device-resident task array.
struct ChunkRange { int begin, count; };
struct LinalgPlan {
  int n = 0;            // matrix order the plan was built for
  int rows = 0;         // rows actually factored (n, or n + 1 for the augmented hlm system)
  int nblk = 0;
  int shape_panel = 1, shape_inner = 1, shape_outer = 1;   // GEMM tile shapes (0: 128x128, 1: 128x64, 3: 32x64)
  std::vector<ChunkRange> panel, trail;         // per 64-wide potrf step (trail: inside the outer block)
  std::vector<ChunkRange> outer;                // per outer block: update of everything to its right
  std::vector<ChunkRange> outer_next, outer_bulk;   // the same tasks split: next outer block's columns / the rest
  std::vector<ChunkRange> outer_first, outer_rest;  // outer_next split again: its first 64 columns / the others
  std::vector<ChunkRange> inv1, inv2;           // per trtri level
  ChunkRange prod{0, 0};                        // T^T T of an inverted factor (spsample)
  GemmTask* d_tasks = nullptr;
  int n_tasks = 0;
};

// single != 0: schedule for one matrix at a time (hlm chain): small tiles on the latency-critical steps
extern long long* g_diag_dbg;
extern int g_diag_dbg_device;
int plan_build(LinalgPlan& plan, int n, int extra_rows, int single);
void plan_free(LinalgPlan& plan);

// Batched Cholesky of the leading n x n lower triangle of L (in place).  If extra_rows > 0 the rows
// n .. n+extra_rows-1 (full-length rows appended below the matrix) are carried through the panel
// solves: on exit row n holds (L^-1 y)^T when it held y^T on entry.  dinv: [batch][nblk][NB*NB] scratch
// for the inverted diagonal blocks.  linv (optional): receives the inverted diagonal blocks (lower part
// and mirrored upper part) as the level-0 state of the triangular inverse.
// status[b] is set to the (1-based) order of the first non-positive leading minor, 0 if none.
// Optional lookahead schedule (single-matrix chains): a second stream and per-outer-block events.
struct Lookahead {
  cudaStream_t bulk = nullptr;
  int mode = 1;   // 1: bulk of each trailing update concurrent with the next panel; 2: bulk cut into pieces that run in
                  // the shadow of the single-CTA diagonal-block kernels of the next outer block
  std::vector<cudaEvent_t> ev_chain, ev_bulk;   // mode 1: one pair per outer block; modes 2, 3: one pair per 64-wide step
  std::vector<cudaEvent_t> ev_aux;              // modes 2, 3: one per outer block
};
int potrf_batched(const LinalgPlan& plan, MatView L, double* dinv, MatView* linv, int batch, int* status,
                  cudaStream_t st, Lookahead* la = nullptr);

// Batched inverse of the lower-triangular factor: on exit the lower triangle of linv holds L^-1 and the
// upper triangle its transpose.  Uses the strict upper triangle of L as scratch.  Requires the
// level-0 state written by potrf_batched.
int trtri_batched(const LinalgPlan& plan, MatView L, MatView linv, int batch, cudaStream_t st);

// out = Linv^T Linv (full symmetric matrix) from the state trtri_batched leaves in linv: chol2inv() without dpotri
int gram_of_inverse_batched(const LinalgPlan& plan, MatView linv, MatView out, int batch, cudaStream_t st);
// M_b = alpha[b] * M_b + diag[b] * I
int scale_add_diag_batched(int n, MatView M, const double* alpha, const double* diag, int batch, cudaStream_t st);
// v_b = M_b x_b (+ add_b); lower_only: use only k <= row
int gemv_batched(int n, MatView M, const double* x, long long ldx, const double* add, long long ldadd, double* v,
                 long long ldv, int lower_only, int batch, cudaStream_t st);

// v[b][i] = sum_{k<=i} T[b][i][k] * x[b][k]   (lower-triangular matrix-vector product)
int trmv_lower_batched(int n, MatView T, const double* x, long long ldx, double* v, long long ldv, int batch,
                       cudaStream_t st);

// sum_i log(L[i][i]) and sum_k row[k]^2 per batch entry (hlm log-determinant and quadratic form)
int logdet_quad_batched(int n, MatView L, double* logdet_half, double* quad, int batch, cudaStream_t st);

}  // namespace spw
