// Single-matrix blocked Cholesky (order n, NB = 64) with a two-step look-ahead, for the strictly sequential chain of
// hlmBayes_sp (chol() of R/hlmBayes_sp.R:114,140,181: one matrix of order N + 1 per proposal).
// DEFAULT from order 1024 on (SPW_LA=0 disables it, SPW_LA_MIN moves the threshold): 7-16 % less time than the shadow
// schedule of potrf_batched for orders 1000 .. 10000 (2.10 against 2.50 ms at 5001; hlmBayes_sp at N = 5000: 150
// against 125 iterations/s).
//
// The right-looking schedule has the critical path  diag(j) -> panel(j) -> update(j) -> diag(j+1).  Here the chain
// of diagonal blocks is ONE single-CTA kernel that stays resident for the whole factorisation and finishes its own
// row block itself, so that step j depends on step j-1 and on work launched TWO steps earlier only:
//
//   D(j)      [tile (j, j-1) -= L(j, j-2) L(j-1, j-2)^T          (source j-2, the last one missing): done by U1(j-1)]
//             L(j, j-1) = tile (j, j-1) * Dinv(j-1)^T           (panel product of its own row); la_done = j
//             tile (j, j) -= L(j, j-1) L(j, j-1)^T              (source j-1)
//             factor and invert tile (j, j)                     (the sweep of diag_factor.cuh)
//             starts when LLS(j) and U1(j-1) have finished (counters their CTAs increment), publishes d_done = j + 1.
//   side stream (highest priority), per step j:
//     PANEL2(j)  L(r, j) = (A(r, j) - L(r, j-1) L(j, j-1)^T) Dinv(j)^T, rows >= j+2; its CTAs wait for d_done > j and
//                count themselves in p2_done[j]
//   second side stream:
//     LLS(c)     sources s0 .. c-2 -> column c, rows >= c, ONE product with K = 64 (c - 1 - s0)  (left-looking over
//                the sources the bulk updates have not covered: s0 = 4 (c / 4 - 1)).  Launched a step EARLY, after
//                PANEL2(c-3); its last two k-tiles (source c-2) are loaded once p2_done[c-2] is complete.  (While the
//                trailing matrix has more than SPW_LA_EARLY_ROWS = 5100 rows: launched after PANEL2(c-2), side stream.)
//   third side stream:
//     U1(j)      tile (j+1, j) -= L(j+1, j-1) L(j, j-1)^T (PANEL2's kernel, update only, two CTAs) while the sweep of
//                block j runs: waits for la_done >= j and p2_done[j-1], counts itself in u1_done[j+1]
//   bulk streams (lowest priority), when the panels of outer block m (256 columns) are final:
//     G(m, cb), cb >= m+2      K = 256, target cb on stream cb mod 3, the nearest one (awaited by the first LLS into
//                              block m+2, two to three steps later) on a stream of its own
//     (SPW_LA_HORIZON = H: LL(0..m -> m+H), one product with K = 256 (m+1) when block m+H comes near, instead)
//
// Every product but PANEL2 / U1 is a task list of the TMA-fed DMMA kernel of linalg.cu.  Stream order, one event per
// outer block and target, and the flags are the only synchronisation; the pattern forks from and joins the caller's
// stream, so it is captured into a CUDA graph like any other.  Deterministic: every tile receives its updates in a
// fixed order.  A flag wait gives up after two seconds (status LA_STALLED).
//
// What the per-step trace (SPW_LA_TRACE=1, tools/la_trace.py) shows at order 5001: a chain step takes 17.5 us
// (loads + 3.2 us of look-ahead products at one SM's 250 GFLOP/s + 12.8 us sweep + hand-over), the side chain of a
// step 12 us.  The first ~30 steps are bound by the machine (~30 TFLOP/s sustained), the last ~50 by the chain.
// In between (outer blocks 2 .. 12): while the ~3000 CTAs of a bulk update wait for SM slots, a kernel launched on
// a side stream gets its first CTA running 15-60 us late (with a free slot on every SM and with the highest stream
// priority alike) and then shares its SM with two bulk CTAs.  Tried against it, all slower: two bulk CTAs per SM
// (shared-memory padding), bulk launches of 296 / 441 CTAs that stay resident and walk the task list (the side
// kernels then start at once, but two bulk CTAs per SM sustain only ~17 TFLOP/s and with three the side CTAs get a
// quarter of their SM: PANEL2 50 us, LLS 60 us), bulk streams by distance, 128 x 64 bulk tiles, outer blocks of
// 128 .. 512 columns, left-looking horizons 3 .. 8, four-stage rings for LLS, the nearest target's update in two
// halves.  There is no priority inside an SM.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "potrf_la.cuh"
#include "diag_factor.cuh"
#include "tile64.cuh"

namespace spw {

namespace {
constexpr int XP_DOUBLES = DIAG_LOW * DIAG_TS;    // packed inverse of one diagonal block
struct LaSmem {
  DiagSmem diag;                 // the 64 x 68 tile (j, j-1) lives over diag.Lt | diag.Xt until the sweep starts
  double bufB[C_TILE];           // L(j-1, j-2), then the updated tile (j, j) = input of the sweep
  double xp[XP_DOUBLES];         // packed inverse of diagonal block j-1
};
static_assert(2 * DIAG_LOW * DIAG_TS >= C_TILE, "tile (j, j-1) must fit over Lt | Xt");
static_assert(offsetof(DiagSmem, Xt) == sizeof(double) * DIAG_LOW * DIAG_TS, "Lt and Xt must be contiguous");
}  // namespace

// One step of the chain: the look-ahead products of row block j and the sweep of diagonal block j.  sm.xp holds the
// packed inverse of diagonal block j-1 on entry and that of block j on exit.
// la_done != nullptr: tile (j, j-1) arrives with source j-2 already applied (U1 kernel, see below), and L(j, j-1) is
// announced through *la_done = j as soon as it is in global memory (the U1 launch of the next row block waits for it).
__device__ __forceinline__ void la_step(LaSmem& sm, MatView L, int j, int n, double* __restrict__ dinv, int nblk,
                                        int* __restrict__ status, int* __restrict__ la_done,
                                        unsigned long long* __restrict__ tr = nullptr) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const MatView none{nullptr, 0, 0};
  const int r0 = j * NB, bs = min(NB, n - r0);
  const long long ld = L.ld;
  double* Lg = L.p;
  if (j > 0) {
    double* T = sm.diag.Lt;
    const bool two = j >= 2 && !la_done;
    const double* a1 = Lg + (long long)r0 * ld + (j - 1) * NB;            // tile (j, j-1)
    const double* b1 = Lg + (long long)(r0 - NB) * ld + (j - 2) * NB;     // L(j-1, j-2)
    const double* x1p = Lg + (long long)r0 * ld + (j - 2) * NB;           // L(j, j-2)
    const double* sp = Lg + (long long)r0 * ld + r0;                      // tile (j, j)
    // ---- every global load of the step is issued before the first one is used ----
    constexpr int NV = NB * NB / 2, PT = (NV + DIAG_THREADS - 1) / DIAG_THREADS;
    double2 vt[PT], vb[PT];
#pragma unroll
    for (int i = 0; i < PT; ++i) {
      const int e = tid + i * DIAG_THREADS, r = e >> 5, c = (e & 31) * 2;
      vt[i] = vb[i] = make_double2(0.0, 0.0);
      if (e < NV) {
        if (r < bs) vt[i] = __ldcg(reinterpret_cast<const double2*>(a1 + (long long)r * ld + c));
        if (two) vb[i] = __ldcg(reinterpret_cast<const double2*>(b1 + (long long)r * ld + c));
      }
    }
    double x1[16];
#pragma unroll
    for (int k4 = 0; k4 < 16; ++k4) x1[k4] = 0.0;
    if (two && warp < 8 && 8 * warp + g < bs) {
#pragma unroll
      for (int k4 = 0; k4 < 16; ++k4) x1[k4] = __ldcg(x1p + (long long)(8 * warp + g) * ld + 4 * k4 + q);
    }
    double2 sc[3];
    int smt[3], snt[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      int t = warp + 12 * i, mt = 0;
      while (t > mt) { t -= mt + 1; ++mt; }
      smt[i] = mt; snt[i] = t;
      const int row = 8 * mt + g, col = 8 * t + 2 * q;
      sc[i] = make_double2(0.0, 0.0);
      if (row < bs) {
        if (col + 1 < bs) sc[i] = __ldcg(reinterpret_cast<const double2*>(sp + (long long)row * ld + col));
        else if (col < bs) sc[i].x = __ldcg(sp + (long long)row * ld + col);
      }
    }
#pragma unroll
    for (int i = 0; i < PT; ++i) {
      const int e = tid + i * DIAG_THREADS, r = e >> 5, c = (e & 31) * 2;
      if (e < NV) {
        *reinterpret_cast<double2*>(T + r * C_LD + c) = vt[i];
        *reinterpret_cast<double2*>(sm.bufB + r * C_LD + c) = vb[i];
      }
    }
    __syncthreads();
    if (warp < 8) {
      if (two) {
        // ---- rows 8 warp .. 8 warp + 7 of tile (j, j-1) -= L(j, j-2) L(j-1, j-2)^T ----
        double acc[8][2];
        double* crow = T + (8 * warp + g) * C_LD + 2 * q;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const double2 v = *reinterpret_cast<const double2*>(crow + 8 * nt);
          acc[nt][0] = v.x; acc[nt][1] = v.y;
        }
        const double* brow = sm.bufB + g * C_LD + q;
#pragma unroll
        for (int k4 = 0; k4 < 16; ++k4) {
          const double a = -x1[k4];
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) dmma884(acc[nt][0], acc[nt][1], a, brow[8 * nt * C_LD + 4 * k4]);
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<double2*>(crow + 8 * nt) = make_double2(acc[nt][0], acc[nt][1]);
        __syncwarp();
      }
      // ---- L(j, j-1) = tile * Dinv(j-1)^T: in place and to global memory ----
      chain_panel_rows(T, sm.xp, warp, Lg + (long long)r0 * ld + (j - 1) * NB, ld, bs, NB, lane);
    }
    __syncthreads();
    // st.release.gpu after the CTA barrier publishes the writes of every thread of the CTA (release is cumulative);
    // no separate fence.sc in front of it
    if (la_done && tid == DIAG_THREADS - 32) st_release(la_done, j);   // a lane of the last warp: off the pivot warp's scheduler
    // ---- tile (j, j) -= L(j, j-1) L(j, j-1)^T: 36 lower 8 x 8 tiles, three per warp; the result feeds the sweep ----
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double2 v = sc[i];
      const double* arow = T + (8 * smt[i] + g) * C_LD + q;
      const double* brow = T + (8 * snt[i] + g) * C_LD + q;
#pragma unroll
      for (int k4 = 0; k4 < 16; ++k4) dmma884(v.x, v.y, -arow[4 * k4], brow[4 * k4]);
      *reinterpret_cast<double2*>(sm.bufB + (8 * smt[i] + g) * C_LD + 8 * snt[i] + 2 * q) = v;
    }
    __syncthreads();
    if (tr && tid == 0) tr[7] = globaltimer();           // the sweep starts
    diag_factor_block<false, false, true>(sm.diag, L, 0, j, n, dinv, nblk, none, 0, status, nullptr, sm.bufB);
  } else {
    diag_factor_block<true>(sm.diag, L, 0, j, n, dinv, nblk, none, 0, status, nullptr);
  }
  for (int e = tid; e < XP_DOUBLES / 2; e += DIAG_THREADS)
    reinterpret_cast<double2*>(sm.xp)[e] = reinterpret_cast<const double2*>(sm.diag.Xt)[e];
}

// The whole chain of diagonal blocks as ONE single-CTA kernel that stays resident for the factorisation (its SM
// slot and shared memory are never contended).  Coupling with the products the side stream launches meanwhile:
//   * step j starts once the LLS(j) launch has finished (every CTA of it adds 1 to lls_done[j]);
//   * d_done = j + 1 is published once the outputs of step j are in global memory; the CTAs of UCOL(j-1) / PANEL(j)
//     wait for it (GemmSync of linalg.cu).
// A wait that lasts longer than two seconds gives up (status LA_STALLED): a scheduling accident must not hang the
// device or go unnoticed.
__global__ void __launch_bounds__(DIAG_THREADS, 1)
diag_chain_kernel(MatView L, int n, double* __restrict__ dinv, int nblk, int* __restrict__ status,
                  int* __restrict__ flags, const int* __restrict__ expect, unsigned long long* __restrict__ trace,
                  int* __restrict__ la_done, const int* __restrict__ u1_done) {
  extern __shared__ __align__(16) unsigned char lsm[];
  LaSmem& sm = *reinterpret_cast<LaSmem*>(lsm);
  const int tid = threadIdx.x;
  bool stalled = false;                    // thread 0: a wait has given up; the result is void, do not wait again
  for (int j = 0; j < nblk; ++j) {
    if (trace && tid == 0) trace[8 * j] = globaltimer();
    if (tid == 0 && j >= 2 && !stalled) {
      const int want = expect[j];
      const unsigned long long t0 = globaltimer();
      const int want_u = (min(NB, n - j * NB) + 31) / 32;      // CTAs of the U1 launch for row block j
      while (ld_acquire(flags + 1 + j) < want || (u1_done && ld_acquire(u1_done + j) < want_u)) {
        if (globaltimer() - t0 > 2000000000ull) {
          atomicExch(status, LA_STALLED);
          stalled = true;
          break;
        }
      }
    }
    __syncthreads();
    if (trace && tid == 0) trace[8 * j + 1] = globaltimer();
    la_step(sm, L, j, n, dinv, nblk, status, la_done, trace ? trace + 8 * j : nullptr);
    __syncthreads();
    if (tid == 0) {
      st_release(flags, j + 1);
      if (trace) trace[8 * j + 2] = globaltimer();
    }
  }
}


// PANEL(j) with the last missing update folded in, for the rows below the diagonal kernels' band (row r >= rbeg):
//   L(r, j) = (A(r, j) - L(r, j-1) L(j, j-1)^T) Dinv(j)^T
// One CTA per 64 rows, 8 warps x 8 rows, both products on the tensor pipe out of shared memory, one round of global
// loads.  Replaces two launches of the generic tile kernel (each ~9 us of latency chain for K = 64) on the side
// stream.  The CTAs start once the chain kernel has published step j (d_done >= j + 1).
struct Panel2Smem {
  double T[C_TILE / 2];  // A(r, j), up to 32 rows -> updated rows
  double B[C_TILE];      // L(j, j-1), then Dinv(j): 52 KB per CTA, so that it fits beside resident bulk CTAs
};

template <int NW>     // warps per CTA = 8-row groups (4: 32 rows per CTA)
__global__ void __launch_bounds__(32 * NW, 3)    // <= 168 registers: a CTA fits the slot a 32 x 64 GEMM CTA frees
panel2_kernel(MatView L, int j, int n, int rows, int rbeg, const double* __restrict__ dinv, const int* __restrict__ d_done,
              int d_want, int* __restrict__ status, unsigned long long* __restrict__ stamp, int* __restrict__ p2_done,
              const int* __restrict__ flag2, int want2, int update_only) {
  extern __shared__ __align__(16) unsigned char psm[];
  Panel2Smem& sm = *reinterpret_cast<Panel2Smem*>(psm);
  constexpr int NT = 32 * NW, RB = 8 * NW;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int r0 = rbeg + blockIdx.x * RB, nr = min(RB, rows - r0), bs = min(NB, n - j * NB);
  const long long ld = L.ld;
  if (d_done) {
    if (tid == 0) {
      const unsigned long long t0 = globaltimer();
      while (ld_acquire(d_done) < d_want || (flag2 && ld_acquire(flag2) < want2)) {
        if (globaltimer() - t0 > 2000000000ull || ld_relaxed(status) == LA_STALLED) {
          atomicExch(status, LA_STALLED);
          break;
        }
        __nanosleep(40);
      }
    }
    __syncthreads();
  }
  if (stamp && tid == 0) atomicMin(stamp, globaltimer());
  const bool upd = j >= 1;
  const double* a = L.p + (long long)r0 * ld + j * NB;                     // A(r, j)
  const double* x1p = L.p + (long long)r0 * ld + (j - 1) * NB;             // L(r, j-1)
  const double* b1 = L.p + (long long)(j * NB) * ld + (j - 1) * NB;        // L(j, j-1)
  const double* dv = dinv + (long long)j * NB * NB;                        // Dinv(j), ld = NB, zero outside bs x bs
  constexpr int PA = RB * 32 / NT, PB = NB * 32 / NT;                      // double2 per thread: own rows / 64 x 64 operands
  double2 vt[PA], vb[PB], vd[PB];
#pragma unroll
  for (int i = 0; i < PA; ++i) {
    const int e = tid + i * NT, r = e >> 5, c = (e & 31) * 2;
    vt[i] = make_double2(0.0, 0.0);
    if (r < nr) {
      if (c + 1 < bs) vt[i] = __ldcg(reinterpret_cast<const double2*>(a + (long long)r * ld + c));
      else if (c < bs) vt[i].x = __ldcg(a + (long long)r * ld + c);
    }
  }
#pragma unroll
  for (int i = 0; i < PB; ++i) {
    const int e = tid + i * NT, r = e >> 5, c = (e & 31) * 2;
    vb[i] = make_double2(0.0, 0.0);
    vd[i] = make_double2(0.0, 0.0);
    if (upd && r < bs) vb[i] = __ldcg(reinterpret_cast<const double2*>(b1 + (long long)r * ld + c));
    if (!update_only) vd[i] = __ldcg(reinterpret_cast<const double2*>(dv + r * NB + c));
  }
  double x1[16];
#pragma unroll
  for (int k4 = 0; k4 < 16; ++k4) x1[k4] = 0.0;
  if (upd && 8 * warp + g < nr) {
#pragma unroll
    for (int k4 = 0; k4 < 16; ++k4) x1[k4] = __ldcg(x1p + (long long)(8 * warp + g) * ld + 4 * k4 + q);
  }
#pragma unroll
  for (int i = 0; i < PA; ++i) {
    const int e = tid + i * NT, r = e >> 5, c = (e & 31) * 2;
    *reinterpret_cast<double2*>(sm.T + r * C_LD + c) = vt[i];
  }
#pragma unroll
  for (int i = 0; i < PB; ++i) {
    const int e = tid + i * NT, r = e >> 5, c = (e & 31) * 2;
    *reinterpret_cast<double2*>(sm.B + r * C_LD + c) = upd ? vb[i] : vd[i];
  }
  __syncthreads();
  double acc[8][2] = {};
  double* crow = sm.T + (8 * warp + g) * C_LD + 2 * q;
  if (upd) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const double2 v = *reinterpret_cast<const double2*>(crow + 8 * nt);
      acc[nt][0] = v.x; acc[nt][1] = v.y;
    }
    const double* brow = sm.B + g * C_LD + q;
#pragma unroll
    for (int k4 = 0; k4 < 16; ++k4) {
      const double av = -x1[k4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) dmma884(acc[nt][0], acc[nt][1], av, brow[8 * nt * C_LD + 4 * k4]);
    }
    if (!update_only) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<double2*>(crow + 8 * nt) = make_double2(acc[nt][0], acc[nt][1]);
      __syncthreads();                   // every warp is done with L(j, j-1): the buffer takes Dinv(j)
#pragma unroll
      for (int i = 0; i < PB; ++i) {
        const int e = tid + i * NT, r = e >> 5, c = (e & 31) * 2;
        *reinterpret_cast<double2*>(sm.B + r * C_LD + c) = vd[i];
      }
      __syncthreads();
    }
  }
  if (!update_only) {
    // rows 8 warp .. 8 warp + 7 times Dinv(j)^T (lower triangular: k <= column)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
    const double* arow = sm.T + (8 * warp + g) * C_LD + q;
    const double* brow = sm.B + g * C_LD + q;
#pragma unroll
    for (int k4 = 0; k4 < 16; ++k4) {
      const double av = arow[4 * k4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        if (nt >= (k4 >> 1)) dmma884(acc[nt][0], acc[nt][1], av, brow[8 * nt * C_LD + 4 * k4]);
    }
  }
  // update_only: acc holds A(r, j) - L(r, j-1) L(j, j-1)^T, written back in place (the tile the chain kernel picks up)
  const int r = 8 * warp + g;
  if (r < nr) {
    double* dst = L.p + (long long)(r0 + r) * ld + j * NB;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = 8 * nt + 2 * q;
      if (c + 1 < bs) __stcg(reinterpret_cast<double2*>(dst + c), make_double2(acc[nt][0], acc[nt][1]));
      else if (c < bs) __stcg(dst + c, acc[nt][0]);
    }
  }
  if (stamp || p2_done) {
    __syncthreads();
    if (tid == 0) {
      if (p2_done) {                       // the late k-tiles of LLS(j + 2), already running, wait for every CTA of this launch
        __threadfence();
        atomicAdd(p2_done, 1);
      }
      if (stamp) atomicMax(stamp + 1, globaltimer());
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------
static GemmTask la_blank() {
  GemmTask t;
  memset(&t, 0, sizeof(t));
  return t;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

static const int SMALL_M = 32, SMALL_N = 64;     // tile shape 3 of linalg.cu: the latency-critical products

int la_plan_build(LaPlan& plan, int n, int extra_rows) {
  la_plan_free(plan);
  plan.n = n;
  plan.rows = n + extra_rows;
  plan.nblk = ceil_div(n, NB);
  plan.ow = NB * std::max(1, env_int("SPW_LA_OUTER", OUTER_W / NB));      // columns per outer block of the bulk updates
  const int OW = plan.ow;
  plan.nouter = ceil_div(n, OW);
  plan.horizon = std::max(3, env_int("SPW_LA_HORIZON", 1 << 20));
  const int per = OW / NB, H = plan.horizon, rows = plan.rows, nblk = plan.nblk;
  std::vector<GemmTask> tasks;
  const int BULK_M = SMALL_M, BULK_N = SMALL_N;
  plan.lls.assign(nblk, ChunkRange{0, 0});
  auto bsz = [&](int c) { return std::min(NB, n - c * NB); };
  // LLS(c): sources s0(c) .. c-2 -> column c, rows >= c
  for (int c = 2; c < nblk; ++c) {
    const int s0 = per * std::max(0, c / per - 1), K = NB * (c - 1 - s0);
    plan.lls[c].begin = (int)tasks.size();
    for (int r = c * NB; r < rows; r += SMALL_M) {
      GemmTask t = la_blank();
      t.a_row = r; t.a_col = s0 * NB;
      t.b_row = c * NB; t.b_col = s0 * NB;
      t.c_row = r; t.c_col = c * NB;
      t.m = std::min(SMALL_M, rows - r); t.n = bsz(c); t.k = K;
      t.flags = GT_ACCUM | GT_STORE;
      tasks.push_back(t);
    }
    plan.lls[c].count = (int)tasks.size() - plan.lls[c].begin;
  }
  // bulk updates of outer block cb by the source outer blocks [m0, m1]: lower tiles, column by column
  auto bulk = [&](int m0, int m1, int cb) {
    const int c_begin = cb * OW, c_end = std::min(c_begin + OW, n);
    for (int c = c_begin; c < c_end; c += BULK_N)
      for (int r = c_begin; r < rows; r += BULK_M) {
        const int m = std::min(BULK_M, rows - r);
        if (c > r + m - 1) continue;
        GemmTask t = la_blank();
        t.a_row = r; t.a_col = m0 * OW;
        t.b_row = c; t.b_col = m0 * OW;
        t.c_row = r; t.c_col = c;
        t.m = m; t.n = std::min(BULK_N, n - c); t.k = (m1 - m0 + 1) * OW;
        t.flags = GT_ACCUM | GT_STORE;
        tasks.push_back(t);
      }
  };
  plan.bulk.assign(plan.nouter, {});
  for (int m = 0; m < plan.nouter; ++m)
    for (int cb = m + 2; cb <= m + H && cb < plan.nouter; ++cb) {
      LaPlan::BulkLaunch bl;
      bl.cb = cb;
      bl.tasks.begin = (int)tasks.size();
      if (cb == m + H) bulk(0, m, cb);        // outer block cb enters the horizon: every source block so far at once
      else bulk(m, m, cb);
      bl.tasks.count = (int)tasks.size() - bl.tasks.begin;
      if (bl.tasks.count > 0) plan.bulk[m].push_back(bl);
    }
  plan.n_tasks = (int)tasks.size();
  if (plan.n_tasks > 0) {
    SPW_CUDA(cudaMalloc(&plan.d_tasks, sizeof(GemmTask) * tasks.size()));
    SPW_CUDA(cudaMemcpy(plan.d_tasks, tasks.data(), sizeof(GemmTask) * tasks.size(), cudaMemcpyHostToDevice));
  }
  std::vector<int> expect(nblk, 0);
  for (int c = 0; c < nblk; ++c) expect[c] = plan.lls[c].count;
  SPW_CUDA(cudaMalloc(&plan.d_flags, sizeof(int) * (size_t)(3 * nblk + 2)));
  SPW_CUDA(cudaMalloc(&plan.d_expect, sizeof(int) * (size_t)nblk));
  if (env_int("SPW_LA_TRACE", 0)) {
    SPW_CUDA(cudaMalloc(&plan.d_trace, sizeof(unsigned long long) * 8 * (size_t)nblk));
  }
  SPW_CUDA(cudaMemcpy(plan.d_expect, expect.data(), sizeof(int) * (size_t)nblk, cudaMemcpyHostToDevice));
  return SPW_OK;
}

void la_plan_free(LaPlan& plan) {
  if (plan.d_tasks) cudaFree(plan.d_tasks);
  if (plan.d_flags) cudaFree(plan.d_flags);
  if (plan.d_expect) cudaFree(plan.d_expect);
  if (plan.d_trace) cudaFree(plan.d_trace);
  plan = LaPlan();
}

int la_trace_read(const LaPlan& plan, std::vector<unsigned long long>& out) {
  out.clear();
  if (!plan.d_trace) return SPW_OK;
  out.resize(8 * (size_t)plan.nblk);
  SPW_CUDA(cudaMemcpy(out.data(), plan.d_trace, sizeof(unsigned long long) * out.size(), cudaMemcpyDeviceToHost));
  return SPW_OK;
}

int la_streams_reserve(LaStreams& s, int nblk) {
  if (!s.side) {
    int lo = 0, hi = 0;
    SPW_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    SPW_CUDA(cudaStreamCreateWithPriority(&s.side, cudaStreamNonBlocking, hi));
    SPW_CUDA(cudaStreamCreateWithPriority(&s.side2, cudaStreamNonBlocking, hi));
    SPW_CUDA(cudaStreamCreateWithPriority(&s.side2b, cudaStreamNonBlocking, hi));
    SPW_CUDA(cudaEventCreateWithFlags(&s.ev_join2b, cudaEventDisableTiming));
    SPW_CUDA(cudaStreamCreateWithPriority(&s.side3, cudaStreamNonBlocking, hi));
    SPW_CUDA(cudaEventCreateWithFlags(&s.ev_join3, cudaEventDisableTiming));
    SPW_CUDA(cudaEventCreateWithFlags(&s.ev_join2, cudaEventDisableTiming));
    for (int i = 0; i < LaStreams::NBULK; ++i) {
      // bulk[3] carries the nearest target of every outer block: one priority level above the rest where there is one
      SPW_CUDA(cudaStreamCreateWithPriority(&s.bulk[i], cudaStreamNonBlocking, (i == 3 && lo - hi >= 2) ? lo - 1 : lo));   // (the priority makes no measurable difference)
      SPW_CUDA(cudaEventCreateWithFlags(&s.ev_bjoin[i], cudaEventDisableTiming));
    }
    SPW_CUDA(cudaEventCreateWithFlags(&s.ev_fork, cudaEventDisableTiming));
    SPW_CUDA(cudaEventCreateWithFlags(&s.ev_join, cudaEventDisableTiming));
  }
  auto grow = [&](std::vector<cudaEvent_t>& v, size_t need) -> int {
    while (v.size() < need) {
      cudaEvent_t e;
      SPW_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      v.push_back(e);
    }
    return SPW_OK;
  };
  SPW_TRY(grow(s.ev_s, nblk + 1));
  SPW_TRY(grow(s.ev_p, nblk + 2));
  SPW_TRY(grow(s.ev_g, nblk + 2));
  return SPW_OK;
}

int potrf_lookahead(const LaPlan& plan, LaStreams& s, MatView L, double* dinv, int* status, cudaStream_t st) {
  const int nblk = plan.nblk, per = plan.ow / NB;
  SPW_TRY(la_streams_reserve(s, nblk));
  static bool attr_set[64] = {false};
  int dev = 0;
  SPW_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    SPW_CUDA(cudaFuncSetAttribute(diag_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LaSmem)));
    SPW_CUDA(cudaFuncSetAttribute(panel2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Panel2Smem)));
    // every kernel the side streams launch must be loaded before the chain kernel starts: a first-use module load
    // waits for the device to drain, which the resident chain kernel (waiting for those very launches) never lets
    // happen
    cudaFuncAttributes fa;
    SPW_CUDA(cudaFuncGetAttributes(&fa, panel2_kernel<4>));
    SPW_TRY(gemm_prepare(3));
    attr_set[dev & 63] = true;
  }
  MatView none{nullptr, 0, 0};
  Operand opL{L, plan.rows, plan.n};
  auto run = [&](const ChunkRange& r, cudaStream_t q, const GemmSync* sy) -> int {
    if (r.count == 0) return SPW_OK;
    return launch_gemm_tasks(3, plan.d_tasks + r.begin, r.count, 1, opL, opL, L, none, -1.0, q, false, sy);
  };
  int* d_done = plan.d_flags;
  int* lls_done = plan.d_flags + 1;
  int* p2_done = plan.d_flags + 1 + nblk;
  // SPW_LA_EARLY (default 1): LLS(j + 2) is launched one step early, on a second side stream, right after PANEL2(j - 1):
  // it works through the sources that exist and loads its last two k-tiles (source j) once every CTA of PANEL2(j)
  // has counted itself in p2_done[j].  The side chain of a step shrinks from PANEL2 + LLS to PANEL2 + two k-tiles.
  static const bool early = env_int("SPW_LA_EARLY", 1) != 0;
  // SPW_LA_U1 (default 1, needs SPW_LA_EARLY): the first look-ahead product of the chain, tile (j+1, j) -= L(j+1, j-1)
  // L(j, j-1)^T, is taken off the chain: a two-CTA launch per step on a third side stream does it while the sweep of
  // block j runs (it waits for la_done >= j, which the chain kernel publishes right after L(j, j-1), and for PANEL2(j-1))
  static const bool u1 = early && env_int("SPW_LA_U1", 1) != 0;
  int* la_done = plan.d_flags + 2 * nblk + 1;
  int* u1_done = plan.d_flags + 2 * nblk + 2;
  static const bool by_distance = env_int("SPW_LA_BULK_BY_DISTANCE", 1) != 0;
  static const int bulk_delay = by_distance ? env_int("SPW_LA_BULK_DELAY", 1) : 0;
  std::vector<char> touched(plan.nouter + 1, 0);
  SPW_CUDA(cudaMemsetAsync(plan.d_flags, 0, sizeof(int) * (size_t)(3 * nblk + 2), st));
  if (plan.d_trace) {                                        // start stamps are minima, end stamps maxima
    for (int j = 0; j < nblk; ++j) {
      SPW_CUDA(cudaMemsetAsync(plan.d_trace + 8 * j, 0, 8 * sizeof(unsigned long long), st));
      SPW_CUDA(cudaMemsetAsync(plan.d_trace + 8 * j + 3, 0xff, sizeof(unsigned long long), st));
      SPW_CUDA(cudaMemsetAsync(plan.d_trace + 8 * j + 5, 0xff, sizeof(unsigned long long), st));
    }
  }
  // fork: the side and bulk streams join the caller's stream (and its graph capture)
  SPW_CUDA(cudaEventRecord(s.ev_fork, st));
  SPW_CUDA(cudaStreamWaitEvent(s.side, s.ev_fork, 0));
  SPW_CUDA(cudaStreamWaitEvent(s.side2, s.ev_fork, 0));
  SPW_CUDA(cudaStreamWaitEvent(s.side2b, s.ev_fork, 0));
  SPW_CUDA(cudaStreamWaitEvent(s.side3, s.ev_fork, 0));
  for (int i = 0; i < LaStreams::NBULK; ++i) SPW_CUDA(cudaStreamWaitEvent(s.bulk[i], s.ev_fork, 0));
  diag_chain_kernel<<<1, DIAG_THREADS, sizeof(LaSmem), st>>>(L, plan.n, dinv, nblk, status, plan.d_flags, plan.d_expect,
                                                             plan.d_trace, u1 ? la_done : nullptr, u1 ? u1_done : nullptr);
  SPW_LAUNCHED();
  auto p2_ctas = [&](int jj) { return std::max(0, ceil_div(plan.rows - std::min((jj + 2) * NB, plan.n), 32)); };
  if (u1) {
    for (int j = 1; j + 1 < nblk; ++j) {       // U1(j): source j-1 applied to tile (j+1, j)
      const int rb0 = (j + 1) * NB, rend = std::min((j + 2) * NB, plan.n);
      panel2_kernel<4><<<ceil_div(rend - rb0, 32), 128, sizeof(Panel2Smem), s.side3>>>(
          L, j, plan.n, rend, rb0, dinv, la_done, j, status, nullptr, u1_done + (j + 1), p2_done + (j - 1), p2_ctas(j - 1), 1);
      SPW_LAUNCHED();
    }
  }
  // LLS(c) on stream q; late_src >= 0: its last two k-tiles (source late_src = c - 2) wait for PANEL2(late_src)
  auto launch_lls = [&](int c, cudaStream_t q, int late_src) -> int {
    GemmSync done;
    done.done_ctr = lls_done + c;
    done.abort_flag = status;
    done.abort_val = LA_STALLED;
    if (plan.d_trace) done.stamp = plan.d_trace + 8 * c + 5;
    if (late_src >= 0) {                   // sources late_src .. c - 2 (one or two) come from panels that are still to run
      const int nlate = c - 1 - late_src;
      done.late_flag = p2_done + late_src;
      for (int g = 0; g < nlate && g < 2; ++g) done.late_val[g] = p2_ctas(late_src + g);
      done.late_group = NB / 32;
      done.late_tiles = nlate * (NB / 32);
    }
    // an LLS into outer block cb follows the last bulk update into it, G(cb - 2, cb) (every LLS waits, not only the
    // first one of the block: they are spread over several streams)
    if (c / per >= 2) SPW_CUDA(cudaStreamWaitEvent(q, s.ev_g[c / per], 0));
    return run(plan.lls[c], q, &done);
  };
  static const int early_rows = env_int("SPW_LA_EARLY_ROWS", 5100);
  static const int early2_rows = env_int("SPW_LA_EARLY2_ROWS", 3500);
  auto depth = [&](int c) {
    if (!early || c < 3) return 0;
    const int left = plan.rows - c * NB;
    if (c >= 4 && left <= early2_rows) return 2;
    return left <= early_rows ? 1 : 0;
  };
  for (int j = 0; j < nblk; ++j) {
    const int rbeg = std::min((j + 2) * NB, plan.n), nrow = plan.rows - rbeg;
    if (nrow > 0) {
      panel2_kernel<4><<<ceil_div(nrow, 32), 128, sizeof(Panel2Smem), s.side>>>(L, j, plan.n, plan.rows, rbeg, dinv, d_done,
                                                                                 j + 1, status, plan.d_trace ? plan.d_trace + 8 * j + 3 : nullptr,
                                                                                 early ? p2_done + j : nullptr, nullptr, 0, 0);
      SPW_LAUNCHED();
    }
    SPW_CUDA(cudaEventRecord(s.ev_s[j], s.side));            // PANEL2(j) has finished: sources <= j exist
    // bulk updates by outer block m once its panels are final: the nearest target (awaited two to three steps from
    // now) at once and on a stream of its own, so that it never queues behind the previous block's updates of far
    // targets; the other targets (stream cb mod 3) one step later (SPW_LA_BULK_DELAY), so that the nearest one has
    // the machine to itself for a step.  Successive updates of one target are ordered through its event.
    auto launch_bulk = [&](int m, bool nearest, cudaEvent_t after) -> int {
      int idx = 0;
      for (const auto& bl : plan.bulk[m]) {
        const bool urgent = by_distance && idx == 0;
        ++idx;
        if (urgent != nearest) continue;
        cudaStream_t q = s.bulk[urgent ? 3 : bl.cb % 3];
        SPW_CUDA(cudaStreamWaitEvent(q, after, 0));
        if (urgent && touched[bl.cb]) SPW_CUDA(cudaStreamWaitEvent(q, s.ev_g[bl.cb], 0));
        SPW_TRY(run(bl.tasks, q, nullptr));
        SPW_CUDA(cudaEventRecord(s.ev_g[bl.cb], q));
        touched[bl.cb] = 1;
      }
      return SPW_OK;
    };
    if ((j + 1) % per == 0 && !plan.bulk[j / per].empty()) {      // the panels of outer block j / per are final
      SPW_TRY(launch_bulk(j / per, true, s.ev_s[j]));
      if (bulk_delay == 0 || j + 1 >= nblk) SPW_TRY(launch_bulk(j / per, false, s.ev_s[j]));
    }
    if (bulk_delay > 0 && j >= 1 && j % per == 0 && !plan.bulk[j / per - 1].empty()) SPW_TRY(launch_bulk(j / per - 1, false, s.ev_s[j]));
    // LLS(c): depth 0 -- launched now for c = j + 2, on the side stream; depth 1 -- a step early (c = j + 3, second
    // side stream, last source through p2_done); depth 2 -- two steps early (c = j + 4, alternating between two
    // streams, last two sources through p2_done).  Early launches pay when the chain is the bottleneck: their launch
    // latency under load is hidden; while the machine is the bottleneck (a large trailing matrix) their waiting CTAs
    // only take SM slots from the bulk updates.
    for (int d = 0; d <= 2; ++d) {
      const int c = j + 2 + d;
      if (c >= nblk || depth(c) != d) continue;
      if (d == 0) {
        SPW_TRY(launch_lls(c, s.side, -1));
      } else {
        cudaStream_t q = (d == 2 && (c & 1)) ? s.side2b : s.side2;
        SPW_CUDA(cudaStreamWaitEvent(q, s.ev_s[j], 0));
        SPW_TRY(launch_lls(c, q, j + 1));
      }
    }
  }
  // join
  SPW_CUDA(cudaEventRecord(s.ev_join, s.side));
  SPW_CUDA(cudaStreamWaitEvent(st, s.ev_join, 0));
  SPW_CUDA(cudaEventRecord(s.ev_join2, s.side2));
  SPW_CUDA(cudaStreamWaitEvent(st, s.ev_join2, 0));
  SPW_CUDA(cudaEventRecord(s.ev_join2b, s.side2b));
  SPW_CUDA(cudaStreamWaitEvent(st, s.ev_join2b, 0));
  SPW_CUDA(cudaEventRecord(s.ev_join3, s.side3));
  SPW_CUDA(cudaStreamWaitEvent(st, s.ev_join3, 0));
  for (int i = 0; i < LaStreams::NBULK; ++i) {
    SPW_CUDA(cudaEventRecord(s.ev_bjoin[i], s.bulk[i]));
    SPW_CUDA(cudaStreamWaitEvent(st, s.ev_bjoin[i], 0));
  }
  return SPW_OK;
}

}  // namespace spw
