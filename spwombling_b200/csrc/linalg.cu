// Batched blocked Cholesky (right-looking, NB = 64 diagonal blocks inside 256-wide outer blocks, 128x128 DMMA tiles) and
// level-synchronous divide-and-conquer triangular inverse, both driven by precomputed tile task lists.
// Replaces the LAPACK calls the reference reaches through base R:
//   chol()      R/hlmBayes_sp.R:99,114,140,181; R/spatial_gradient.R:216,265,304   (dpotrf)
//   chol2inv()  same lines (dpotri) -- not formed: K^-1 = Linv^T Linv is applied through Linv only.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "linalg.cuh"
#include "tma.cuh"
#include "diag_factor.cuh"

namespace spw {

// ---------------------------------------------------------------------------------------------------
// Task-driven NT GEMM on the FP64 tensor pipe: one CTA per (task, batch entry), BM x BN output tile, 8 warps (4 x 2).
// The A and B k-tiles arrive by TMA (boxes of 36 x BM and 36 x BN doubles: 32 columns used + the 4-double pad that
// keeps the DMMA fragment loads conflict-free) into a ring guarded by mbarriers; the load of tile kt+STAGES-1 is
// issued by one lane of warp (kt mod 8) once tile kt has landed, so no warp is dedicated to loading (a 9th warp
// would be allocated as 12 and cap the registers at 168) and there is no CTA barrier in the loop.  Rows / columns of
// a box outside the task are either other matrix data (their accumulators are not stored) or outside the tensor
// (zero-filled by TMA), so the only masking left is the triangular one.
//   L: 128 x 128, 3 stages, 1 CTA/SM   (long-k products of the triangular inverse)
//   M: 128 x  64, 2 stages, 2 CTA/SM   (Cholesky panel / trailing updates: the second CTA hides the C read-modify-write)
//   T:  32 x  64, 2 stages, 3 CTA/SM   (every step of a single matrix: many short CTAs; for the panel product rows,
//                                       not columns, are split because it overwrites its own A operand)
// ---------------------------------------------------------------------------------------------------
namespace {
constexpr int G_BK = 32, G_LDS = G_BK + 4, G_CW = 8, G_THREADS = 32 * G_CW;

template <int BM_, int BN_, int STAGES_, int MINB_>
struct GCfg {
  static constexpr int BM = BM_, BN = BN_, STAGES = STAGES_, MINB = MINB_;
  static constexpr int WM = BM / 4, WN = BN / 2, MF = WM / 8, NF = WN / 8;
  static constexpr int A_STAGE = BM * G_LDS, B_STAGE = BN * G_LDS, STAGE = A_STAGE + B_STAGE;
  static constexpr uint32_t STAGE_BYTES = STAGE * sizeof(double);
  static constexpr size_t SMEM = size_t(STAGES) * STAGE_BYTES + 2 * STAGES * sizeof(unsigned long long) + 64;
};
using GCfgL = GCfg<128, 128, 3, 1>;
using GCfgM = GCfg<128, 64, 2, 2>;
using GCfgT = GCfg<32, 64, 2, 3>;
}  // namespace

template <class Cfg>
__global__ void __launch_bounds__(G_THREADS, Cfg::MINB)
gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const GemmTask* __restrict__ tasks, MatView C, MatView CT, double alpha, const GemmSync sy) {
  extern __shared__ __align__(128) unsigned char gsm[];
  double* pipe = reinterpret_cast<double*>(gsm);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(gsm + size_t(Cfg::STAGES) * Cfg::STAGE_BYTES);
  unsigned long long* empty = full + Cfg::STAGES;
  const GemmTask t = tasks[blockIdx.x];
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nkt = (t.k + G_BK - 1) / G_BK;
  if (tid == 0) {
    if (sy.stamp) atomicMin(sy.stamp, globaltimer());
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], G_CW);
    }
    mbar_fence_init();
    if (sy.wait_flag) {                       // operands written by a kernel that is still running (see GemmSync)
      const unsigned long long t0 = globaltimer();          // give up after 2 s: never hang the device
      while (ld_acquire(sy.wait_flag) < sy.wait_val && globaltimer() - t0 < 2000000000ull &&
             !(sy.abort_flag && ld_relaxed(sy.abort_flag) == sy.abort_val))
        __nanosleep(40);
      fence_proxy_async();
    }
  }
  __syncthreads();
  auto issue = [&](int kt) {   // one thread: wait for the slot, arm the barrier, launch the two tile loads
    const int s = kt % Cfg::STAGES;
    mbar_wait(&empty[s], ((kt / Cfg::STAGES) & 1) ^ 1);
    if (sy.late_flag && kt >= nkt - sy.late_tiles) {
      const int grp = min(1, (kt - (nkt - sy.late_tiles)) / sy.late_group);
      const unsigned long long t0 = globaltimer();
      while (ld_acquire(sy.late_flag + grp) < sy.late_val[grp] && globaltimer() - t0 < 2000000000ull &&
             !(sy.abort_flag && ld_relaxed(sy.abort_flag) == sy.abort_val))
        __nanosleep(40);
      fence_proxy_async();
    }
    double* st = pipe + (size_t)s * Cfg::STAGE;
    mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
    tma_load_3d(st, &tmA, t.a_col + kt * G_BK, t.a_row, b, &full[s]);
    tma_load_3d(st + Cfg::A_STAGE, &tmB, t.b_col + kt * G_BK, t.b_row, b, &full[s]);
  };
  if (tid == 0) {
    for (int kt = 0; kt < Cfg::STAGES - 1 && kt < nkt; ++kt) issue(kt);
  }
  const int warp_m = warp & 3, warp_n = warp >> 2;
  const int g = lane >> 2, q = lane & 3;
  double acc[Cfg::MF][Cfg::NF][2];
#pragma unroll
  for (int i = 0; i < Cfg::MF; ++i)
#pragma unroll
    for (int j = 0; j < Cfg::NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int kt = 0; kt < nkt; ++kt) {
    const int s = kt % Cfg::STAGES;
    mbar_wait(&full[s], (kt / Cfg::STAGES) & 1);
    if (lane == 0 && warp == (kt & (G_CW - 1)) && kt + Cfg::STAGES - 1 < nkt) issue(kt + Cfg::STAGES - 1);
    __syncwarp();
    const double* As = pipe + (size_t)s * Cfg::STAGE;
    const double* a_base = As + (warp_m * Cfg::WM + g) * G_LDS + q;
    const double* b_base = As + Cfg::A_STAGE + (warp_n * Cfg::WN + g) * G_LDS + q;
    const int k0 = kt * G_BK;
    bool masked = false;
    if (t.a_tri == 1) masked |= (k0 + G_BK - 1 > t.a_lim0);
    if (t.a_tri == 2) masked |= (k0 < t.a_lim0 + Cfg::BM - 1);
    if (t.b_tri == 1) masked |= (k0 + G_BK - 1 > t.b_lim0);
    if (t.b_tri == 2) masked |= (k0 < t.b_lim0 + Cfg::BN - 1);
#pragma unroll
    for (int kk = 0; kk < G_BK / 4; ++kk) {
      double a[Cfg::MF], bf[Cfg::NF];
#pragma unroll
      for (int i = 0; i < Cfg::MF; ++i) a[i] = a_base[i * 8 * G_LDS + kk * 4];
#pragma unroll
      for (int j = 0; j < Cfg::NF; ++j) bf[j] = b_base[j * 8 * G_LDS + kk * 4];
      if (masked) {
        const int k = k0 + kk * 4 + q;
        if (t.a_tri == 1) {
#pragma unroll
          for (int i = 0; i < Cfg::MF; ++i)
            if (k > t.a_lim0 + warp_m * Cfg::WM + i * 8 + g) a[i] = 0.0;
        } else if (t.a_tri == 2) {
#pragma unroll
          for (int i = 0; i < Cfg::MF; ++i)
            if (k < t.a_lim0 + warp_m * Cfg::WM + i * 8 + g) a[i] = 0.0;
        }
        if (t.b_tri == 1) {
#pragma unroll
          for (int j = 0; j < Cfg::NF; ++j)
            if (k > t.b_lim0 + warp_n * Cfg::WN + j * 8 + g) bf[j] = 0.0;
        } else if (t.b_tri == 2) {
#pragma unroll
          for (int j = 0; j < Cfg::NF; ++j)
            if (k < t.b_lim0 + warp_n * Cfg::WN + j * 8 + g) bf[j] = 0.0;
        }
      }
#pragma unroll
      for (int i = 0; i < Cfg::MF; ++i)
#pragma unroll
        for (int j = 0; j < Cfg::NF; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bf[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
  // ---- epilogue: C = / += alpha * acc, optional transposed store.  The C reads of one fragment row are issued
  //      together (NF loads in flight) before they are consumed. ----
  double* Cp = C.p + (long long)b * C.stride + (long long)t.c_row * C.ld + t.c_col;
  double* CTp = (t.flags & GT_STORE_T) ? CT.p + (long long)b * CT.stride + (long long)t.ct_row * CT.ld + t.ct_col : nullptr;
  const bool accum = (t.flags & GT_ACCUM) != 0, store = (t.flags & GT_STORE) != 0;
#pragma unroll
  for (int i = 0; i < Cfg::MF; ++i) {
    const int r = warp_m * Cfg::WM + i * 8 + g;
    if (r >= t.m) continue;
    double* crow = Cp + (long long)r * C.ld;
    double2 old[Cfg::NF];
#pragma unroll
    for (int j = 0; j < Cfg::NF; ++j) {
      const int c = warp_n * Cfg::WN + j * 8 + 2 * q;
      old[j] = make_double2(0.0, 0.0);
      if (accum && c < t.n) {
        if (c + 1 < t.n) old[j] = *reinterpret_cast<const double2*>(crow + c);
        else old[j].x = crow[c];
      }
    }
#pragma unroll
    for (int j = 0; j < Cfg::NF; ++j) {
      const int c = warp_n * Cfg::WN + j * 8 + 2 * q;
      if (c >= t.n) continue;
      const double v0 = fma(alpha, acc[i][j][0], old[j].x), v1 = fma(alpha, acc[i][j][1], old[j].y);
      const bool two = (c + 1 < t.n);
      if (store) {
        if (two) *reinterpret_cast<double2*>(crow + c) = make_double2(v0, v1);
        else crow[c] = v0;
      }
      if (CTp) {
        CTp[(long long)c * CT.ld + r] = v0;
        if (two) CTp[(long long)(c + 1) * CT.ld + r] = v1;
      }
    }
  }
  if (sy.done_ctr) {
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(sy.done_ctr, 1);
      if (sy.stamp) atomicMax(sy.stamp + 1, globaltimer());
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Diagonal block: factor the bs x bs lower block (bs <= 64) and invert the factor in the same sweep.
// The sweep is bound by ONE warp: the one that carries the pivot chain (cycle probes: a lone warp issues a DFMA every
// 2.1 cycles, a dependent one every 8.3, the reciprocal square root takes 49; that warp ends up limited by the issue
// of its ~260 instructions per step, not by the chain alone), so the design keeps that warp's instruction stream
// minimal and divergence-free, keeps its scheduler free of other busy warps and gives every other job to other
// warps.  384 threads in 12 warps:
//   * warp 0: the owners of tile column s (4 pivots), two lanes per tile (two rows each).  They read the column as
//     published one step earlier, apply the pending rank-4 update of column s-1 to the pivot tile and to their own
//     tile, factor the pivot tile (redundantly, reciprocal square roots) and solve X = C Ld^-T: final tiles of L.
//   * 136 threads (warps 1,2,3,5,6) each keep one 4 x 4 tile (ti, tk), ti >= tk, of the trailing block in
//     registers, apply the update of column s-1 and publish column s+1 for the next step.
//   * 120 threads (warps 7,9,10,11) each build one off-diagonal 4 x 4 tile (i, c) of X = L^-1,
//     X[i][c] = -Xd_i sum_{k=c}^{i-1} L[i][k] X[k][c], one tile product per step as its operands become final.
//   * one thread of warp 4 turns the pivot-tile numbers of the previous step into the diagonal tiles of L and X and
//     checks the pivots; warp 8 is idle (warps 4 and 8 sit on warp 0's scheduler).
// Measured and dropped: letting the other warps wait for warp 0's shared-memory loads (named barrier), moving the
// solve to a second warp fed through shared memory (+14 % sweep time: the hand-off costs more than the overlap of
// solve and pivot chain it gives up), issuing the four reciprocal square roots together.
// One barrier per step, 18 steps; final tiles live in shared memory (packed 4 x 4, 18-double stride: conflict-free
// 128-bit accesses) and all global writes (factor, 64 x 64 inverse block for the panel product, mirrored copy in
// the triangular-inverse buffer) happen once at the end, by all threads.
// ---------------------------------------------------------------------------------------------------
long long* g_diag_dbg = nullptr;   // optional device buffer of 16 cycle counters (spw_debug_diag_timing)
int g_diag_dbg_device = -1;        // the device it was allocated on

__global__ void __launch_bounds__(DIAG_THREADS, 1)
diag_factor_kernel(MatView L, int blk, int n, double* __restrict__ dinv, int nblk, MatView linv, int has_linv,
                   int* __restrict__ status, long long* __restrict__ dbg) {
  __shared__ __align__(16) DiagSmem sm;
  diag_factor_block<false>(sm, L, blockIdx.x, blk, n, dinv, nblk, linv, has_linv, status, dbg);
}

// ---------------------------------------------------------------------------------------------------
// v = T x for lower-triangular T: one warp per row.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
trmv_lower_kernel(int n, MatView T, const double* __restrict__ x, long long ldx, double* __restrict__ v, long long ldv) {
  const int b = blockIdx.y;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const double* Tp = T.p + (long long)b * T.stride + (long long)row * T.ld;
  const double* xp = x + (long long)b * ldx;
  double s = 0.0;
  for (int k = lane; k <= row; k += 32) s = fma(Tp[k], xp[k], s);
  s = warp_sum(s);
  if (lane == 0) v[(long long)b * ldv + row] = s;
}

// sum log diag(L) and ||row n||^2 (deterministic single-CTA reduction per batch entry)
__global__ void __launch_bounds__(1024)
logdet_quad_kernel(int n, MatView L, double* __restrict__ logdet_half, double* __restrict__ quad) {
  __shared__ double red[2][32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const double* Lp = L.p + (long long)b * L.stride;
  double sl = 0.0, sq = 0.0;
  for (int i = tid; i < n; i += blockDim.x) {
    sl += log(Lp[(long long)i * L.ld + i]);
    if (quad) {
      const double w = Lp[(long long)n * L.ld + i];
      sq = fma(w, w, sq);
    }
  }
  sl = warp_sum(sl);
  sq = warp_sum(sq);
  if ((tid & 31) == 0) {
    red[0][tid >> 5] = sl;
    red[1][tid >> 5] = sq;
  }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, c = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      a += red[0][w];
      c += red[1][w];
    }
    logdet_half[b] = a;
    if (quad) quad[b] = c;
  }
}

// ---------------------------------------------------------------------------------------------------
// Host side: task schedule
// ---------------------------------------------------------------------------------------------------
static GemmTask blank_task() {
  GemmTask t;
  memset(&t, 0, sizeof(t));
  return t;
}

static void shape_dims(int shape, int& tm, int& tn) {
  tm = (shape == 3) ? 32 : 128;
  tn = (shape == 0) ? 128 : 64;
}

int plan_build(LinalgPlan& plan, int n, int extra_rows, int single) {
  plan_free(plan);
  plan.n = n;
  plan.rows = n + extra_rows;
  plan.nblk = ceil_div(n, NB);
  // tile shapes: a single matrix needs many small CTAs on the latency-critical 64-wide steps; batches fill the
  // machine with 128 x 64 tiles everywhere
  plan.shape_panel = single ? 3 : 1;
  plan.shape_inner = single ? 3 : 1;
  plan.shape_outer = single ? 3 : 1;     // one matrix: small tiles keep the tail of every launch short and free their
                                         // SM slot within microseconds (the shadow schedule depends on that)
  std::vector<GemmTask> tasks;
  const int T = 128;   // tile edge of the triangular-inverse products (shape L)
  plan.panel.resize(plan.nblk);
  plan.trail.resize(plan.nblk);
  plan.outer.clear();
  plan.outer_next.clear();
  plan.outer_bulk.clear();
  plan.outer_first.clear();
  plan.outer_rest.clear();
  int pm, pn, im, in_, om, on;
  shape_dims(plan.shape_panel, pm, pn);
  shape_dims(plan.shape_inner, im, in_);
  shape_dims(plan.shape_outer, om, on);
  (void)pn;
  // Right-looking with two block sizes: inside an outer block of OUTER_W columns the 64-wide steps update only the
  // rest of that outer block (K = 64, few tiles); the matrix to the right is updated once per outer block with
  // K = OUTER_W, where the tensor pipe runs long k loops and every C tile is read and written once per 256 columns.
  for (int J0 = 0; J0 < n; J0 += OUTER_W) {
    const int J1 = std::min(J0 + OUTER_W, n);
    for (int c0 = J0; c0 < J1; c0 += NB) {
      const int j = c0 / NB, bs = std::min(NB, n - c0), r_begin = c0 + bs;
      // panel: rows below the diagonal block (including appended rows): L_ij = A_ij * Dinv_j^T
      plan.panel[j].begin = (int)tasks.size();
      for (int r = r_begin; r < plan.rows; r += pm) {
        GemmTask t = blank_task();
        t.a_row = r; t.a_col = c0;
        t.b_row = j * NB; t.b_col = 0;          // in the dinv view (ld = NB)
        t.c_row = r; t.c_col = c0;
        t.m = std::min(pm, plan.rows - r); t.n = bs; t.k = bs;
        t.flags = GT_STORE;
        tasks.push_back(t);
      }
      plan.panel[j].count = (int)tasks.size() - plan.panel[j].begin;
      // inner trailing update: columns of this outer block to the right of block j
      plan.trail[j].begin = (int)tasks.size();
      for (int r = r_begin; r < plan.rows; r += im) {
        const int m = std::min(im, plan.rows - r);
        for (int c = r_begin; c < J1 && c <= r + m - 1; c += in_) {
          GemmTask t = blank_task();
          t.a_row = r; t.a_col = c0;
          t.b_row = c; t.b_col = c0;
          t.c_row = r; t.c_col = c;
          t.m = m; t.n = std::min(in_, J1 - c); t.k = bs;
          t.flags = GT_ACCUM | GT_STORE;
          tasks.push_back(t);
        }
      }
      plan.trail[j].count = (int)tasks.size() - plan.trail[j].begin;
    }
    // outer trailing update: C[I][K] -= L[I][J0:J1] * L[K][J0:J1]^T for the lower triangle right of the outer block;
    // listed as three consecutive groups: the first 64 columns of the NEXT outer block (what its first step needs),
    // the other columns of that block (needed from its first in-block update on), and the rest (the bulk)
    ChunkRange o, oa, ob, of, orr;
    o.begin = oa.begin = of.begin = (int)tasks.size();
    const int strip = std::max(on, NB);
    for (int part = 0; part < 3; ++part) {
      if (part == 1) orr.begin = (int)tasks.size();
      if (part == 2) ob.begin = (int)tasks.size();
      for (int r = J1; r < plan.rows; r += om) {
        const int m = std::min(om, plan.rows - r);
        for (int c = J1; c < n && c <= r + m - 1; c += on) {
          const int p = (c < J1 + strip) ? 0 : (c < J1 + OUTER_W) ? 1 : 2;
          if (p != part) continue;
          GemmTask t = blank_task();
          t.a_row = r; t.a_col = J0;
          t.b_row = c; t.b_col = J0;
          t.c_row = r; t.c_col = c;
          t.m = m; t.n = std::min(on, n - c); t.k = J1 - J0;
          t.flags = GT_ACCUM | GT_STORE;
          tasks.push_back(t);
        }
      }
      if (part == 0) of.count = (int)tasks.size() - of.begin;
      if (part == 1) {
        orr.count = (int)tasks.size() - orr.begin;
        oa.count = (int)tasks.size() - oa.begin;
      }
    }
    ob.count = (int)tasks.size() - ob.begin;
    o.count = (int)tasks.size() - o.begin;
    plan.outer.push_back(o);
    plan.outer_next.push_back(oa);
    plan.outer_bulk.push_back(ob);
    plan.outer_first.push_back(of);
    plan.outer_rest.push_back(orr);
  }
  // triangular inverse levels: s = NB, 2NB, ... ; pairs (block1 = [o, o+s), block2 = [o+s, min(o+2s, n)))
  for (int s = NB; s < n; s *= 2) {
    ChunkRange r1, r2;
    r1.begin = (int)tasks.size();
    for (int o = 0; o + s < n; o += 2 * s) {
      const int b1 = o, b2 = o + s, r2n = std::min(s, n - b2);
      // GEMM1: T[m][nn] = sum_{k >= nn} L21[m][k] * Linv11[k][nn]; stored transposed in the upper triangle of L
      for (int m0 = 0; m0 < r2n; m0 += T)
        for (int n0 = 0; n0 < s; n0 += T) {
          GemmTask t = blank_task();
          t.a_row = b2 + m0; t.a_col = b1 + n0;
          t.b_row = b1 + n0; t.b_col = b1 + n0;     // upper triangle of linv = Linv^T
          t.ct_row = b1 + n0; t.ct_col = b2 + m0;   // T^T into L's upper triangle
          t.m = std::min(T, r2n - m0); t.n = std::min(T, s - n0); t.k = s - n0;
          t.b_tri = 2; t.b_lim0 = 0;
          t.flags = GT_STORE_T;
          tasks.push_back(t);
        }
    }
    r1.count = (int)tasks.size() - r1.begin;
    r2.begin = (int)tasks.size();
    for (int o = 0; o + s < n; o += 2 * s) {
      const int b1 = o, b2 = o + s, r2n = std::min(s, n - b2);
      // GEMM2: Linv21[m][nn] = - sum_{k <= m} Linv22[m][k] * T[k][nn]
      for (int m0 = 0; m0 < r2n; m0 += T)
        for (int n0 = 0; n0 < s; n0 += T) {
          GemmTask t = blank_task();
          t.m = std::min(T, r2n - m0); t.n = std::min(T, s - n0); t.k = std::min(r2n, m0 + T);
          t.a_row = b2 + m0; t.a_col = b2;          // Linv22 (lower part of linv)
          t.b_row = b1 + n0; t.b_col = b2;          // T^T in L's upper triangle
          t.c_row = b2 + m0; t.c_col = b1 + n0;     // Linv21
          t.ct_row = b1 + n0; t.ct_col = b2 + m0;   // mirrored
          t.a_tri = 1; t.a_lim0 = m0;
          t.flags = GT_STORE | GT_STORE_T;
          tasks.push_back(t);
        }
    }
    r2.count = (int)tasks.size() - r2.begin;
    plan.inv1.push_back(r1);
    plan.inv2.push_back(r2);
  }
  // P = T^T T for a lower-triangular T whose transpose sits in the upper triangle of the same buffer (the state the
  // triangular inverse leaves): P[m][n] = sum_{k >= max(m, n)} TT[m][k] TT[n][k]; lower tiles, mirrored on store
  plan.prod.begin = (int)tasks.size();
  for (int m0 = 0; m0 < n; m0 += T)
    for (int n0 = 0; n0 <= m0; n0 += T) {
      GemmTask t = blank_task();
      t.a_row = m0; t.a_col = m0;
      t.b_row = n0; t.b_col = m0;
      t.c_row = m0; t.c_col = n0;
      t.ct_row = n0; t.ct_col = m0;
      t.m = std::min(T, n - m0); t.n = std::min(T, n - n0); t.k = n - m0;
      t.a_tri = 2; t.a_lim0 = 0;
      t.b_tri = (n0 == m0) ? 2 : 0; t.b_lim0 = 0;
      t.flags = GT_STORE | GT_STORE_T;
      tasks.push_back(t);
    }
  plan.prod.count = (int)tasks.size() - plan.prod.begin;
  plan.n_tasks = (int)tasks.size();
  if (plan.n_tasks > 0) {
    SPW_CUDA(cudaMalloc(&plan.d_tasks, sizeof(GemmTask) * tasks.size()));
    SPW_CUDA(cudaMemcpy(plan.d_tasks, tasks.data(), sizeof(GemmTask) * tasks.size(), cudaMemcpyHostToDevice));
  }
  return SPW_OK;
}

void plan_free(LinalgPlan& plan) {
  if (plan.d_tasks) cudaFree(plan.d_tasks);
  plan = LinalgPlan();
}


// one_per_sm: pad the dynamic shared memory so that only one CTA fits per SM (leaves room for the CTAs of a
// concurrently running latency-critical kernel)
template <class Cfg>
static int launch_gemm_cfg(const GemmTask* tasks, int count, int batch, Operand A, Operand B, MatView C, MatView CT,
                           double alpha, cudaStream_t st, bool one_per_sm, const GemmSync& sy) {
  if (count <= 0 || batch <= 0) return SPW_OK;
  static bool attr_set[64] = {false};
  int dev = 0;
  SPW_CUDA(cudaGetDevice(&dev));
  const size_t smem_wide = std::max<size_t>(Cfg::SMEM, 117 * 1024);   // two such CTAs exceed the 228 KB of an SM
  if (!attr_set[dev & 63]) {
    SPW_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_wide));
    attr_set[dev & 63] = true;
  }
  const size_t smem = one_per_sm ? smem_wide : Cfg::SMEM;
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = std::min(65535, batch - b0);
    MatView a = A.v, b = B.v, c = C, ct = CT;
    a.p += (long long)b0 * A.v.stride;
    b.p += (long long)b0 * B.v.stride;
    if (c.p) c.p += (long long)b0 * C.stride;
    if (ct.p) ct.p += (long long)b0 * CT.stride;
    CUtensorMap tmA, tmB;
    SPW_TRY(make_tensor_map(&tmA, a, A.cols, A.rows, nb, G_LDS, Cfg::BM));
    SPW_TRY(make_tensor_map(&tmB, b, B.cols, B.rows, nb, G_LDS, Cfg::BN));
    gemm_tma_kernel<Cfg><<<dim3(count, nb), G_THREADS, smem, st>>>(tmA, tmB, tasks, c, ct, alpha, sy);
    SPW_LAUNCHED();
  }
  return SPW_OK;
}

template <class Cfg>
static int prepare_gemm_cfg() {
  SPW_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)std::max<size_t>(Cfg::SMEM, 117 * 1024)));
  cudaFuncAttributes fa;
  SPW_CUDA(cudaFuncGetAttributes(&fa, gemm_tma_kernel<Cfg>));
  return SPW_OK;
}

int gemm_prepare(int shape) {
  switch (shape) {
    case 0: return prepare_gemm_cfg<GCfgL>();
    case 1: return prepare_gemm_cfg<GCfgM>();
    case 3: return prepare_gemm_cfg<GCfgT>();
    default: set_error("gemm: unknown tile shape %d", shape); return SPW_ERR_ARG;
  }
}

// shape: 0 = L (128 x 128), 1 = M (128 x 64), 3 = T (32 x 64); the task list must have been tiled accordingly
int launch_gemm_tasks(int shape, const GemmTask* tasks, int count, int batch, Operand A, Operand B, MatView C,
                      MatView CT, double alpha, cudaStream_t st, bool one_per_sm, const GemmSync* sync) {
  const GemmSync sy = sync ? *sync : GemmSync();
  switch (shape) {
    case 0: return launch_gemm_cfg<GCfgL>(tasks, count, batch, A, B, C, CT, alpha, st, one_per_sm, sy);
    case 1: return launch_gemm_cfg<GCfgM>(tasks, count, batch, A, B, C, CT, alpha, st, one_per_sm, sy);
    case 3: return launch_gemm_cfg<GCfgT>(tasks, count, batch, A, B, C, CT, alpha, st, one_per_sm, sy);
    default: set_error("gemm: unknown tile shape %d", shape); return SPW_ERR_ARG;
  }
}

int potrf_batched(const LinalgPlan& plan, MatView L, double* dinv, MatView* linv, int batch, int* status,
                  cudaStream_t st, Lookahead* la) {
  MatView dv{dinv, NB, (long long)plan.nblk * NB * NB};
  MatView none{nullptr, 0, 0};
  MatView lv = linv ? *linv : none;
  Operand opL{L, plan.rows, plan.n}, opD{dv, plan.nblk * NB, NB};
  const int per_outer = OUTER_W / NB;
  const int nouter = (int)plan.outer.size();
  // SPW_POTRF_PROFILE=1: CUDA-event timing of the four kinds of launches (diagnostic; serialises nothing extra)
  static const bool prof = getenv("SPW_POTRF_PROFILE") && getenv("SPW_POTRF_PROFILE")[0] == '1';
  std::vector<cudaEvent_t> evs;
  auto mark = [&]() {
    if (!prof) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    evs.push_back(e);
  };
  const bool ahead = la && la->mode == 1 && la->bulk && (int)la->ev_chain.size() >= nouter &&
                     (int)la->ev_bulk.size() >= nouter && nouter > 2;
  // Shadow schedule (modes 2 and 3, one matrix): the trailing update of outer block m is cut in three.  Only the
  // first 64 columns of block m+1 are updated on the main stream before its first step.  The other columns of block
  // m+1 start with its first diagonal-block kernel on the second stream (the first in-block update waits for them),
  // followed by the bulk (columns beyond block m+1, no tile in common with the steps of block m+1), which is cut
  // into one piece per step: piece i starts together with the single-CTA diagonal-block kernel of step i, so the
  // other SMs have work while that kernel runs.  Mode 2: the panel product of the step waits for the piece (the
  // small kernels never compete with bulk CTAs); mode 3: only the next trailing update waits (work-conserving;
  // faster with 64 x 64 tiles, whose CTAs free their SM within a few microseconds).
  const bool shadow = la && la->mode >= 2 && la->bulk && batch == 1 && (int)la->ev_chain.size() >= plan.nblk &&
                      (int)la->ev_bulk.size() >= plan.nblk && (int)la->ev_aux.size() >= nouter && nouter > 2;
  int sm_count = 148, pending = -1;   // pending: last piece the main stream has not waited for (mode 3)
  // the cross-stream edges cost a few microseconds each: a trailing update with little bulk runs unsplit instead
  static const int shadow_min = getenv("SPW_SHADOW_MIN") ? atoi(getenv("SPW_SHADOW_MIN")) : 300;
  auto split = [&](int m) { return shadow && m >= 0 && m < nouter && plan.outer_bulk[m].count >= shadow_min; };
  if (shadow) {
    int dev = 0;
    SPW_CUDA(cudaGetDevice(&dev));
    SPW_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  long long* dbg = nullptr;
  if (g_diag_dbg) {
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev == g_diag_dbg_device) dbg = g_diag_dbg;
  }
  for (int j = 0; j < plan.nblk; ++j) {
    mark();
    const int mb = j / per_outer;                      // outer block of this step
    if (split(mb - 1)) SPW_CUDA(cudaEventRecord(la->ev_chain[j], st));
    diag_factor_kernel<<<batch, DIAG_THREADS, 0, st>>>(L, j, plan.n, dinv, plan.nblk, lv, linv ? 1 : 0, status, dbg);
    SPW_LAUNCHED();
    bool rest_wait = false;
    if (split(mb - 1)) {
      const ChunkRange& ob = plan.outer_bulk[mb - 1];
      const ChunkRange& orr = plan.outer_rest[mb - 1];
      const int steps = std::min(per_outer, plan.nblk - mb * per_outer), i = j - mb * per_outer;
      bool forked = false;
      if (i == 0 && orr.count > 0) {            // the columns the first in-block update of this block touches
        SPW_CUDA(cudaStreamWaitEvent(la->bulk, la->ev_chain[j], 0));
        SPW_TRY(launch_gemm_tasks(plan.shape_outer, plan.d_tasks + orr.begin, orr.count, batch, opL, opL, L, none, -1.0,
                                  la->bulk));
        SPW_CUDA(cudaEventRecord(la->ev_aux[mb], la->bulk));
        forked = rest_wait = true;
      }
      // piece i of bulk(mb - 1): whole waves of tiles where the bulk is large
      const int wave = (plan.shape_outer == 3 ? GCfgT::MINB : 2) * (sm_count - 1);   // tiles resident next to the diagonal-block CTA
      int per = (ob.count + steps - 1) / steps;
      if (ob.count >= steps * wave) per = (ob.count / wave / steps) * wave;
      const int b0 = std::min(ob.count, i * per), b1 = (i == steps - 1) ? ob.count : std::min(ob.count, (i + 1) * per);
      if (b1 > b0) {
        if (!forked) SPW_CUDA(cudaStreamWaitEvent(la->bulk, la->ev_chain[j], 0));
        SPW_TRY(launch_gemm_tasks(plan.shape_outer, plan.d_tasks + ob.begin + b0, b1 - b0, batch, opL, opL, L, none, -1.0,
                                  la->bulk));
        forked = true;
      }
      if (forked) {
        SPW_CUDA(cudaEventRecord(la->ev_bulk[j], la->bulk));
        if (la->mode == 2) {
          SPW_CUDA(cudaStreamWaitEvent(st, la->ev_bulk[j], 0));
          rest_wait = false;
        } else {
          pending = j;
        }
      }
    }
    mark();
    SPW_TRY(launch_gemm_tasks(plan.shape_panel, plan.d_tasks + plan.panel[j].begin, plan.panel[j].count, batch, opL, opD,
                              L, none, 1.0, st));
    mark();
    if (rest_wait) SPW_CUDA(cudaStreamWaitEvent(st, la->ev_aux[mb], 0));
    SPW_TRY(launch_gemm_tasks(plan.shape_inner, plan.d_tasks + plan.trail[j].begin, plan.trail[j].count, batch, opL, opL,
                              L, none, -1.0, st));
    mark();
    if ((j + 1) % per_outer == 0 || j + 1 == plan.nblk) {
      const int m = j / per_outer;
      if (shadow) {
        if (pending >= 0) SPW_CUDA(cudaStreamWaitEvent(st, la->ev_bulk[pending], 0));
        pending = -1;
        const ChunkRange& oa = split(m) ? plan.outer_first[m] : plan.outer[m];
        SPW_TRY(launch_gemm_tasks(plan.shape_outer, plan.d_tasks + oa.begin, oa.count, batch, opL, opL, L, none, -1.0, st));
      } else if (!ahead) {
        const ChunkRange& o = plan.outer[m];
        SPW_TRY(launch_gemm_tasks(plan.shape_outer, plan.d_tasks + o.begin, o.count, batch, opL, opL, L, none, -1.0, st));
      } else {
        // Lookahead: the panel of outer block m is final.  The bulk of its trailing update (columns beyond the next
        // outer block) goes to the second stream, one CTA per SM so that the CTAs of the next panel's small kernels
        // find room at once; the main stream only updates the next outer block's columns and moves on to its panel.
        //   bulk(m)  waits for: panel m (ev_chain[m]) and bulk(m-1) (stream order)
        //   next(m)  waits for: bulk(m-1) (ev_bulk[m-1]): it touches columns bulk(m-1) also updates
        const ChunkRange& oa = plan.outer_next[m];
        const ChunkRange& ob = plan.outer_bulk[m];
        SPW_CUDA(cudaEventRecord(la->ev_chain[m], st));
        SPW_CUDA(cudaStreamWaitEvent(la->bulk, la->ev_chain[m], 0));
        static const bool bulk_one_per_sm = !(getenv("SPW_LA_WIDE") && getenv("SPW_LA_WIDE")[0] == '0');
        SPW_TRY(launch_gemm_tasks(plan.shape_outer, plan.d_tasks + ob.begin, ob.count, batch, opL, opL, L, none, -1.0,
                                  la->bulk, bulk_one_per_sm));
        SPW_CUDA(cudaEventRecord(la->ev_bulk[m], la->bulk));
        if (m > 0) SPW_CUDA(cudaStreamWaitEvent(st, la->ev_bulk[m - 1], 0));
        SPW_TRY(launch_gemm_tasks(plan.shape_outer, plan.d_tasks + oa.begin, oa.count, batch, opL, opL, L, none, -1.0, st));
      }
    }
  }
  if (ahead) SPW_CUDA(cudaStreamWaitEvent(st, la->ev_bulk[nouter - 1], 0));   // join
  if (prof && !evs.empty()) {
    mark();
    cudaStreamSynchronize(st);
    double t[4] = {0, 0, 0, 0};
    for (size_t i = 0; i + 1 < evs.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, evs[i], evs[i + 1]);
      t[i % 4] += ms;   // diag, panel, inner, (outer + gap to the next diag)
    }
    fprintf(stderr, "[spw] potrf n=%d batch=%d: diag %.3f ms, panel %.3f ms, inner %.3f ms, outer %.3f ms\n", plan.n, batch,
            t[0], t[1], t[2], t[3]);
    for (auto e : evs) cudaEventDestroy(e);
  }
  return SPW_OK;
}

int trtri_batched(const LinalgPlan& plan, MatView L, MatView linv, int batch, cudaStream_t st) {
  MatView none{nullptr, 0, 0};
  Operand opL{L, plan.n, plan.n}, opI{linv, plan.n, plan.n};
  for (size_t lev = 0; lev < plan.inv1.size(); ++lev) {
    SPW_TRY(launch_gemm_tasks(0, plan.d_tasks + plan.inv1[lev].begin, plan.inv1[lev].count, batch, opL, opI, none, L, 1.0, st));
    SPW_TRY(launch_gemm_tasks(0, plan.d_tasks + plan.inv2[lev].begin, plan.inv2[lev].count, batch, opI, opL, linv, linv, -1.0, st));
  }
  return SPW_OK;
}

int gram_of_inverse_batched(const LinalgPlan& plan, MatView linv, MatView out, int batch, cudaStream_t st) {
  Operand opI{linv, plan.n, plan.n};
  return launch_gemm_tasks(0, plan.d_tasks + plan.prod.begin, plan.prod.count, batch, opI, opI, out, out, 1.0, st);
}

// out_b = alpha_b * out_b + diag_b * I   (full n x n matrices)
__global__ void __launch_bounds__(256)
scale_add_diag_kernel(int n, MatView M, const double* __restrict__ alpha, const double* __restrict__ diag) {
  const int b = blockIdx.z;
  const int r = blockIdx.y;
  const int c = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (c >= n) return;
  double* p = M.p + (long long)b * M.stride + (long long)r * M.ld + c;
  const double a = alpha[b], d = diag[b];
  if (c + 1 < n) {
    double2 v = *reinterpret_cast<double2*>(p);
    v.x = a * v.x + ((r == c) ? d : 0.0);
    v.y = a * v.y + ((r == c + 1) ? d : 0.0);
    *reinterpret_cast<double2*>(p) = v;
  } else {
    p[0] = a * p[0] + ((r == c) ? d : 0.0);
  }
}

int scale_add_diag_batched(int n, MatView M, const double* alpha, const double* diag, int batch, cudaStream_t st) {
  if (batch <= 0) return SPW_OK;
  if (n > 65535 || batch > 65535) { set_error("scale_add_diag: size limits"); return SPW_ERR_ARG; }
  scale_add_diag_kernel<<<dim3(ceil_div(ceil_div(n, 2), 256), n, batch), 256, 0, st>>>(n, M, alpha, diag);
  SPW_LAUNCHED();
  return SPW_OK;
}

// v = M x (+ add) for full row-major matrices: one warp per row
__global__ void __launch_bounds__(256)
gemv_kernel(int n, MatView M, const double* __restrict__ x, long long ldx, const double* __restrict__ add, long long ldadd,
            double* __restrict__ v, long long ldv, int lower_only) {
  const int b = blockIdx.y;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const double* Mp = M.p + (long long)b * M.stride + (long long)row * M.ld;
  const double* xp = x + (long long)b * ldx;
  const int kend = lower_only ? row + 1 : n;
  double s = 0.0;
  for (int k = lane; k < kend; k += 32) s = fma(Mp[k], xp[k], s);
  s = warp_sum(s);
  if (lane == 0) v[(long long)b * ldv + row] = s + (add ? add[(long long)b * ldadd + row] : 0.0);
}

int gemv_batched(int n, MatView M, const double* x, long long ldx, const double* add, long long ldadd, double* v,
                 long long ldv, int lower_only, int batch, cudaStream_t st) {
  if (batch <= 0) return SPW_OK;
  gemv_kernel<<<dim3(ceil_div(n, 8), batch), 256, 0, st>>>(n, M, x, ldx, add, ldadd, v, ldv, lower_only);
  SPW_LAUNCHED();
  return SPW_OK;
}

int trmv_lower_batched(int n, MatView T, const double* x, long long ldx, double* v, long long ldv, int batch,
                       cudaStream_t st) {
  if (batch <= 0) return SPW_OK;
  trmv_lower_kernel<<<dim3(ceil_div(n, 8), batch), 256, 0, st>>>(n, T, x, ldx, v, ldv);
  SPW_LAUNCHED();
  return SPW_OK;
}

int logdet_quad_batched(int n, MatView L, double* logdet_half, double* quad, int batch, cudaStream_t st) {
  if (batch <= 0) return SPW_OK;
  // the diagonal is a strided, latency-bound read: one matrix gets a full-size CTA, batches fill the SMs anyway
  logdet_quad_kernel<<<batch, batch == 1 ? 1024 : 256, 0, st>>>(n, L, logdet_half, quad);
  SPW_LAUNCHED();
  return SPW_OK;
}

}  // namespace spw
