// Single-matrix blocked Cholesky as ONE persistent cooperative kernel (sm_100a, fp64).  EXPERIMENTAL: opt-in with
// SPW_DAG=1; the multi-launch shadow schedule of linalg.cu is still faster at N = 5000 (2.74 ms against 3.2 ms per
// factorisation on the same box, see DESIGN.md section 4 for the per-task traces behind every decision below).
//
// Why: the chain of hlmBayes_sp (R/hlmBayes_sp.R:114,140,181) factors one matrix of order N+1 per proposal, strictly
// one after the other.  As a sequence of launches the factorisation is bound by its critical path (79 diagonal
// blocks, each followed by a panel product and an update before the next diagonal block can start) and by the
// stream edges between ~400 small kernels.  Here the dependencies live in global-memory flags instead:
//
//   * CTA 0 ("chain") factors and inverts the 64 x 64 diagonal blocks one after the other (the warp-role sweep of
//     diag_factor.cuh) and owns the two row blocks below each of them (dag_chain_band);
//   * every other CTA holds two independent 8-warp "sub-workers" (the 128 x 64 / 64 x 64 TMA-fed DMMA tile engine;
//     two per SM so that one hides the epilogue of the other).  A sub-worker holds ONE item at a time:
//       - a row task: one row tile against one source column -- its panel product, then the K = 64 updates of the
//         near columns, left to right, on the same sub-worker (the links of a row's dependency chain, panel(k) ->
//         update(k+1) -> panel(k+1), are the latency every row shares with the chain).  Queued per source column,
//         released when the chain has published that column, taken by ticket in row order;
//       - or a bulk task (K = 256 update), from one queue per source outer block, earliest deadline first.
//   * every block accumulates its bulk updates in a separate buffer and its single-column updates in place, each in
//     source order (deterministic), merged once (see dag_s0 / dag_merges): a late bulk update never holds up the
//     single-column updates of the same block.
//   * flags: upd[row block][column block] / fb[...] count the source columns applied; the panel product of column c
//     needs upd == c and leaves c + 1 ("final").
//
// Deadlock freedom: row tasks and bulk tickets are taken in topological order; a holder polls only its own item; at
// most active_cap sub-workers hold row tasks, so bulk updates always find sub-workers; all CTAs are resident
// (cooperative launch).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "potrf_dag.cuh"
#include "tma.cuh"
#include "diag_factor.cuh"
#include "tile64.cuh"

namespace spw {

namespace {
constexpr int D_BK = 32, D_LDS = D_BK + 4, D_BM = 128, D_BN = 64, D_STAGES = 2;
constexpr int D_SUB = 256;                 // threads of a sub-worker (8 warps)
constexpr int D_THREADS = 2 * D_SUB;       // two sub-workers per CTA; the chain CTA uses 12 of its 16 warps
constexpr int D_A_STAGE = D_BM * D_LDS, D_B_STAGE = D_BN * D_LDS, D_STAGE = D_A_STAGE + D_B_STAGE;   // doubles
constexpr uint32_t D_STAGE_BYTES = D_STAGE * sizeof(double);
constexpr size_t D_SUB_SMEM = size_t(D_STAGES) * D_STAGE_BYTES;
constexpr size_t D_SMEM = 2 * D_SUB_SMEM + 256;
#ifndef SPW_DAG_OUTER
#define SPW_DAG_OUTER 3
#endif
constexpr int OUTER = SPW_DAG_OUTER;       // source columns per bulk update (K = 64 * OUTER); >= 3: see dag_merges
static_assert(sizeof(DiagSmem) <= 2 * D_SUB_SMEM, "chain CTA state must fit in the worker ring");

__device__ __forceinline__ void sub_barrier(int sw) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(1 + sw), "n"(D_SUB) : "memory");
}
__device__ __forceinline__ unsigned smid() {
  unsigned v;
  asm volatile("mov.u32 %0, %%smid;\n" : "=r"(v));
  return v;
}
}  // namespace

struct DagParams {
  const DagTask* utasks;
  const int* useg;
  const DagTask* btasks;   // bulk tasks, one queue per source outer block: queue m = [bseg[m], bseg[m+1]), by target column
  const int* bseg;
  int nq;
  int* upd;        // [nrb][nblk]: main flags (single-column updates and panel products of a block)
  int* fb;         // [nrb][nblk]: source columns accumulated into the bulk buffer of a block
  double* Bp;      // bulk accumulator, same layout as L
  int* uhead;      // [nblk]
  int* bhead;      // [nq]
  int* chain_pos;  // 1
  int* active;     // 1: row tasks being executed
  int active_cap;
  MatView L;
  double* dinv;
  int n, rows, nblk, nrb;
  int* status;
  unsigned long long* trace;
  int trace_cap;
  int* trace_ctr;
};

__device__ __forceinline__ int dag_rstart(int b, int n, int rows, int nblk) {
  return b <= nblk ? min(b * NB, n) : rows;
}

// Every block below the diagonal accumulates its updates in TWO places, each strictly in source-column order (the
// result is deterministic):
//   * the K = 256 bulk updates of the sources [0, s0(c)) go to the bulk buffer B (first one stores, the others add);
//     fb[block] = sources accumulated so far;
//   * the single-column updates of the sources [s0(c), c) are subtracted from the block in place;
//     upd[block] = 0 before the first one, k + 1 after source k, c + 1 ("final") after the panel product.
// The two are merged (block -= B) exactly once: by the last single-column update (source c - 1) -- or, for the three
// blocks of a row that the chain CTA takes over, when it loads them.  A late bulk update therefore never holds up the
// single-column updates of the same block; its only deadline is that merge.
__device__ __forceinline__ int dag_s0(int c) { return OUTER * max(0, c / OUTER - 1); }
// does this single-column update merge the bulk buffer?  The last source (c - 1) of a block below the band of the chain
// CTA; the three blocks (R, R-2 .. R) of a row that the chain finishes itself are merged by the chain when it loads
// them at step R - 1 (the bulk updates onto them then have three more steps than the last helper update would leave).
__device__ __forceinline__ bool dag_merges(const DagTask& t) {
  return t.type == 1 && t.k1 - t.k0 == 1 && t.k0 == t.cb - 1 && t.rb >= t.cb + 3 && dag_s0(t.cb) > 0;
}

// One polling round over the flags a task depends on (warp-collective; every lane reads at most one flag).
//   panel product of column c: the diagonal block is published; every source column < c has been applied to its blocks
//   update: the previous update of the same kind on its block(s) is done; its operand blocks are final; a merging
//   single-column update also needs the bulk buffer of its block complete
// Returns 0: ready; 1: waiting for panel products / single-column updates / the chain; 2: waiting for bulk updates only.
__device__ __forceinline__ int dag_state(const DagTask& t, const int* upd, const int* fb, int nblk, int lane) {
  const int* f = upd;
  int b = -1, c = 0, want = 0;
  bool bulk_dep = false;
  if (t.type == 0) {
    if (lane == 0) { b = t.cb; c = t.cb; want = t.cb + 1; }
    else if (lane <= t.nrb) { b = t.rb + lane - 1; c = t.cb; want = t.cb; }
  } else {
    const int nk = t.k1 - t.k0, na = t.nrb * nk;
    if (lane < t.nrb) {
      b = t.rb + lane; c = t.cb;
      if (nk > 1) { f = fb; want = t.k0; bulk_dep = true; }
      else want = (t.k0 == dag_s0(t.cb)) ? 0 : t.k0;
    } else if (lane < t.nrb + na) { const int i = lane - t.nrb; b = t.rb + i / nk; c = t.k0 + i % nk; want = c + 1; }
    else if (lane < t.nrb + na + nk) { b = t.cb; c = t.k0 + (lane - t.nrb - na); want = c + 1; }
    else if (lane < 2 * t.nrb + na + nk && dag_merges(t)) {
      b = t.rb + (lane - t.nrb - na - nk); c = t.cb; f = fb; want = dag_s0(t.cb); bulk_dep = true;
    }
  }
  bool ok = true;
  if (b >= 0) ok = (ld_acquire(f + b * nblk + c) == want);
  if (__all_sync(0xffffffffu, ok)) return 0;
  return __all_sync(0xffffffffu, ok || bulk_dep) ? 2 : 1;
}
__device__ __forceinline__ bool dag_ready(const DagTask& t, const int* upd, const int* fb, int nblk, int lane) {
  return dag_state(t, upd, fb, nblk, lane) == 0;
}

// NARROW (one row block, m <= 64): the 8 warps are laid out 2 x 4 (warp tile 32 x 16) so that all four tensor pipes
// work on the 64 x 64 tile -- these are the tasks next to the band of the chain CTA, whose latency the chain sees.
template <bool NARROW>
__device__ __forceinline__ void dag_execute(const DagTask& t, const DagParams& P, const CUtensorMap* tmA,
                                            const CUtensorMap* tmB, const CUtensorMap* tmD, double* pipe,
                                            unsigned long long* full, unsigned long long* empty, unsigned& kbase,
                                            int sw, int wtid) {
  const int lane = wtid & 31, warp = wtid >> 5;
  const int row0 = dag_rstart(t.rb, P.n, P.rows, P.nblk);
  const int m = dag_rstart(t.rb + t.nrb, P.n, P.rows, P.nblk) - row0;
  const int c0 = t.cb * NB, ncol = min(NB, P.n - c0);
  const bool panel = (t.type == 0);
  const int a_col = panel ? c0 : t.k0 * NB;
  const int K = panel ? ncol : (t.k1 - t.k0) * NB;
  const int nkt = (K + D_BK - 1) / D_BK;
  const int b_row = c0, b_col = panel ? 0 : t.k0 * NB;
  const CUtensorMap* tb = panel ? tmD : tmB;
  fence_proxy_async();
  auto issue = [&](int kt) {
    const unsigned f = kbase + kt;
    const int s = f % D_STAGES;
    mbar_wait(&empty[s], ((f / D_STAGES) & 1) ^ 1);
    double* st = pipe + (size_t)s * D_STAGE;
    mbar_expect_tx(&full[s], D_STAGE_BYTES);
    tma_load_3d(st, tmA, a_col + kt * D_BK, row0, 0, &full[s]);
    tma_load_3d(st + D_A_STAGE, tb, b_col + kt * D_BK, b_row, 0, &full[s]);
  };
  if (wtid == 0) {
    for (int kt = 0; kt < D_STAGES - 1 && kt < nkt; ++kt) issue(kt);
  }
  constexpr int MF = 4, NF = NARROW ? 2 : 4, WN = 8 * NF;
  const int warp_m = NARROW ? (warp & 1) : (warp & 3), warp_n = NARROW ? (warp >> 1) : (warp >> 2);
  const int g = lane >> 2, q = lane & 3;
  double acc[MF][NF][2];
#pragma unroll
  for (int i = 0; i < MF; ++i)
#pragma unroll
    for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const bool live = warp_m * 32 < m;       // warps whose rows lie outside the tile only keep the ring moving
  for (int kt = 0; kt < nkt; ++kt) {
    const unsigned f = kbase + kt;
    const int s = f % D_STAGES;
    mbar_wait(&full[s], (f / D_STAGES) & 1);
    if (lane == 0 && warp == (kt & 7) && kt + D_STAGES - 1 < nkt) issue(kt + D_STAGES - 1);
    __syncwarp();
    if (live) {
      const double* As = pipe + (size_t)s * D_STAGE;
      const double* a_base = As + (warp_m * 32 + g) * D_LDS + q;
      const double* b_base = As + D_A_STAGE + (warp_n * WN + g) * D_LDS + q;
#pragma unroll
      for (int kk = 0; kk < D_BK / 4; ++kk) {
        double a[MF], bf[NF];
#pragma unroll
        for (int i = 0; i < MF; ++i) a[i] = a_base[i * 8 * D_LDS + kk * 4];
#pragma unroll
        for (int j = 0; j < NF; ++j) bf[j] = b_base[j * 8 * D_LDS + kk * 4];
#pragma unroll
        for (int i = 0; i < MF; ++i)
#pragma unroll
          for (int j = 0; j < NF; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bf[j]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
  kbase += nkt;
  // ---- epilogue.  panel: C = acc.  single-column update: C -= acc (and -= B when it merges the bulk buffer).
  //      bulk update: B = acc (first source block) or B += acc.  The blocks may have been written by other SMs during
  //      this launch: read them past L1 (ld.global.cg) ----
  const bool bulk = !panel && (t.k1 - t.k0 > 1);
  const bool merge = dag_merges(t);
  const bool first_bulk = bulk && t.k0 == 0;
  double* Cp = (bulk ? P.Bp : P.L.p) + (long long)row0 * P.L.ld + c0;
  const double* Bq = P.Bp + (long long)row0 * P.L.ld + c0;
  if (live) {
#pragma unroll
    for (int i = 0; i < MF; ++i) {
      const int r = warp_m * 32 + i * 8 + g;
      if (r >= m) continue;
      double* crow = Cp + (long long)r * P.L.ld;
      const double* brow = Bq + (long long)r * P.L.ld;
      double2 old[NF];
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        const int c = warp_n * WN + j * 8 + 2 * q;
        old[j] = make_double2(0.0, 0.0);
        if (!panel && !first_bulk && c < ncol) {
          if (c + 1 < ncol) old[j] = __ldcg(reinterpret_cast<const double2*>(crow + c));
          else old[j].x = __ldcg(crow + c);
          if (merge) {
            if (c + 1 < ncol) {
              const double2 bv = __ldcg(reinterpret_cast<const double2*>(brow + c));
              old[j].x -= bv.x; old[j].y -= bv.y;
            } else {
              old[j].x -= __ldcg(brow + c);
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        const int c = warp_n * WN + j * 8 + 2 * q;
        if (c >= ncol) continue;
        const double v0 = panel ? acc[i][j][0] : (bulk ? old[j].x + acc[i][j][0] : old[j].x - acc[i][j][0]);
        const double v1 = panel ? acc[i][j][1] : (bulk ? old[j].y + acc[i][j][1] : old[j].y - acc[i][j][1]);
        if (c + 1 < ncol) __stcg(reinterpret_cast<double2*>(crow + c), make_double2(v0, v1));
        else __stcg(crow + c, v0);
      }
    }
  }
  fence_proxy_async();
  sub_barrier(sw);
  if (wtid == 0) {
    __threadfence();
    int* f = bulk ? P.fb : P.upd;
    const int val = panel ? t.cb + 1 : t.k1;
    for (int i = 0; i < t.nrb; ++i) st_release(f + (t.rb + i) * P.nblk + t.cb, val);
  }
}

// ---------------------------------------------------------------------------------------------------
// The chain CTA in band mode: it owns the diagonal blocks AND the two row blocks below each of them.
// Why: with diagonal blocks only, every step of the chain waits for two helper hops (panel product of tile
// (j+1, j), then the update of (j+1, j+1)), each a flag round trip + claim + a small tile product at half an SM's
// rate: 32 us per step measured against 15 us for the sweep itself.  Here the tiles within two blocks of the
// diagonal never leave this SM between their last helper update and their final value:
//   * while warps 0-11 run the sweep of block j, warps 12-15 ("background") fetch the three tiles of row block
//     j+1 -- (j+1, j-1), (j+1, j), (j+1, j+1) -- as soon as the helpers have applied all sources k < j-1, finish
//     L(j+1, j-1) = tile * Dinv_{j-1}^T, and apply source j-1 to the other two;
//   * after the sweep all warps compute L(j+1, j) = tile * Dinv_j^T and S = A(j+1, j+1) - L(j+1, j) L(j+1, j)^T on
//     the tensor pipe straight from shared memory (~2 us), and S is the input of the next sweep.
// The helpers therefore have a whole step of slack for everything the chain consumes.
// ---------------------------------------------------------------------------------------------------
struct ChainSmem {
  DiagSmem diag;
  double xprev[DIAG_LOW * DIAG_TS];   // packed inverse of the previous diagonal block
  double bufS[C_TILE];                // tile (j+1, j+1): input of the next sweep
  double bufX[2][C_TILE];             // tile (j+1, j) -> L(j+1, j); the other one holds L(j, j-1)
  double bufT[C_TILE];                // tile (j+1, j-1) -> L(j+1, j-1)
};
static_assert(sizeof(ChainSmem) <= D_SMEM, "chain state must fit");

__device__ __forceinline__ void dag_chain_band(const DagParams& P, unsigned char* dsm) {
  ChainSmem& sm = *reinterpret_cast<ChainSmem*>(dsm);
  __shared__ unsigned s_ct[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 4) s_ct[tid] = 0;
  const MatView none{nullptr, 0, 0};
  const long long ld = P.L.ld;
  double* Lg = P.L.p;
  auto rstart = [&](int b) { return dag_rstart(b, P.n, P.rows, P.nblk); };
  auto cs = [&](int k) { return min(NB, P.n - k * NB); };
  chain_load_tile<D_THREADS>(sm.bufS, Lg, nullptr, ld, rstart(1) - rstart(0), cs(0), tid);     // A(0, 0)
  __syncthreads();
  for (int j = 0; j < P.nblk; ++j) {
    const int i1 = j + 1;
    const bool has_r1 = i1 < P.nrb, r1_col = i1 < P.nblk;
    const int r1 = rstart(i1), nr = has_r1 ? rstart(i1 + 1) - r1 : 0;
    double* Xcur = sm.bufX[j & 1];
    const double* Xold = sm.bufX[(j & 1) ^ 1];
    unsigned long long t0 = 0;
    if (P.trace && (tid == 0 || tid == 384)) t0 = globaltimer();
    if (warp < 12) {
      // ---- the sweep of diagonal block j (input: bufS) ----
      diag_factor_block<false, true>(sm.diag, P.L, 0, j, P.n, P.dinv, P.nblk, none, 0, P.status, nullptr, sm.bufS);
      fence_proxy_async();
      diag_sync<true>();
      if (tid == 0) {
        __threadfence();
        st_release(P.upd + j * P.nblk + j, j + 1);
        st_release(P.chain_pos, j + 1);
        if (P.trace) s_ct[0] = (unsigned)(globaltimer() - t0);
      }
    } else {
      // ---- background: row block j + 1 ----
      const int bt = tid - 12 * 32, bw = warp - 12;
      auto bg_sync = [&]() { asm volatile("bar.sync 5, 128;\n" ::: "memory"); };
      if (has_r1) {
        if (bw == 0) {
          // the helpers have applied every source k < j - 1 to the three blocks: the single-column updates in place
          // (main flag) and the bulk updates in the bulk buffer (fb), which is merged while loading
          const int want = max(0, j - 1);
          bool ok;
          do {
            ok = true;
            if (lane == 0 && j >= 1) ok = (ld_acquire(P.upd + i1 * P.nblk + (j - 1)) == want);
            if (lane == 1) ok = (ld_acquire(P.upd + i1 * P.nblk + j) == want);
            if (lane == 2 && r1_col) ok = (ld_acquire(P.upd + i1 * P.nblk + i1) == want);
            if (lane == 3 && j >= 1) ok = (ld_acquire(P.fb + i1 * P.nblk + (j - 1)) == dag_s0(j - 1));
            if (lane == 4) ok = (ld_acquire(P.fb + i1 * P.nblk + j) == dag_s0(j));
            if (lane == 5 && r1_col) ok = (ld_acquire(P.fb + i1 * P.nblk + i1) == dag_s0(i1));
          } while (!__all_sync(0xffffffffu, ok));
        }
        bg_sync();
        if (P.trace && bt == 0) s_ct[1] = (unsigned)(globaltimer() - t0);
        if (j >= 1) {
          const long long o = (long long)r1 * ld + (j - 1) * NB;
          chain_load_tile<128>(sm.bufT, Lg + o, dag_s0(j - 1) > 0 ? P.Bp + o : nullptr, ld, nr, NB, bt);
        }
        {
          const long long o = (long long)r1 * ld + j * NB;
          chain_load_tile<128>(Xcur, Lg + o, dag_s0(j) > 0 ? P.Bp + o : nullptr, ld, nr, cs(j), bt);
        }
      }
      asm volatile("bar.sync 4, 512;\n" ::: "memory");      // the sweep threads hold block j in registers: bufS is free
      if (has_r1 && r1_col) {
        const long long o = (long long)r1 * ld + i1 * NB;
        chain_load_tile<128>(sm.bufS, Lg + o, dag_s0(i1) > 0 ? P.Bp + o : nullptr, ld, nr, cs(i1), bt);
      }
      bg_sync();
      if (has_r1 && j >= 1) {
        // L(j+1, j-1) = tile * Dinv_{j-1}^T, 16 rows per warp
        chain_panel_rows(sm.bufT, sm.xprev, 2 * bw, Lg + (long long)r1 * ld + (j - 1) * NB, ld, nr, NB, lane);
        chain_panel_rows(sm.bufT, sm.xprev, 2 * bw + 1, Lg + (long long)r1 * ld + (j - 1) * NB, ld, nr, NB, lane);
        fence_proxy_async();
        bg_sync();
        if (bt == 0) {
          __threadfence();
          st_release(P.upd + i1 * P.nblk + (j - 1), j);
        }
        // source j-1 applied to tiles (j+1, j) and (j+1, j+1)
        chain_update_rows(Xcur, sm.bufT, Xold, 2 * bw, lane);
        chain_update_rows(Xcur, sm.bufT, Xold, 2 * bw + 1, lane);
        if (r1_col)
          for (int t = bw; t < 36; t += 4) chain_syrk_tile(sm.bufS, sm.bufT, t, lane);
      }
      if (P.trace && bt == 0) s_ct[2] = (unsigned)(globaltimer() - t0);
    }
    __syncthreads();
    if (P.trace && tid == 0) {
      const int slot = atomicAdd(P.trace_ctr, 1);
      if (slot < P.trace_cap) {
        unsigned long long* w = P.trace + 4 * (size_t)slot;
        w[0] = (2ull << 48) | ((unsigned long long)j << 32) | (unsigned long long)j << 16;
        w[1] = t0;
        w[2] = globaltimer();
        // sweep end | background flags seen | background end, in units of 16 ns from the start of the step
        w[3] = (unsigned long long)(s_ct[0] >> 4) | ((unsigned long long)(s_ct[1] >> 4) << 20) |
               ((unsigned long long)(s_ct[2] >> 4) << 40);
      }
    }
    if (has_r1) {
      // ---- critical: L(j+1, j) = tile * Dinv_j^T (warps 0-7, 8 rows each); the other warps save Dinv_j ----
      if (warp < 8) {
        chain_panel_rows(Xcur, sm.diag.Xt, warp, Lg + (long long)r1 * ld + j * NB, ld, nr, cs(j), lane);
        fence_proxy_async();
      } else {
        for (int e = tid - 256; e < DIAG_LOW * DIAG_TS / 2; e += 256)
          reinterpret_cast<double2*>(sm.xprev)[e] = reinterpret_cast<const double2*>(sm.diag.Xt)[e];
      }
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        st_release(P.upd + i1 * P.nblk + j, j + 1);
      }
      // ---- critical: S = A(j+1, j+1) - L(j+1, j) L(j+1, j)^T, the input of the next sweep ----
      if (r1_col) {
        for (int t = warp; t < 36; t += 16) chain_syrk_tile(sm.bufS, Xcur, t, lane);
        __syncthreads();
      }
    }
  }
}

__global__ void __launch_bounds__(D_THREADS, 1)
potrf_dag_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmD, const DagParams P) {
  extern __shared__ __align__(128) unsigned char dsm[];
  const int tid = threadIdx.x;
  if (blockIdx.x == 0) {
    dag_chain_band(P, dsm);
    return;
  }
  // ================================ workers: two sub-workers per CTA =====================================
  const int sw = tid >> 8, wtid = tid & 255, warp = wtid >> 5, lane = tid & 31;
  double* pipe = reinterpret_cast<double*>(dsm + (size_t)sw * D_SUB_SMEM);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(dsm + 2 * D_SUB_SMEM) + sw * 2 * D_STAGES;
  unsigned long long* full = bars;
  unsigned long long* empty = bars + D_STAGES;
  __shared__ DagTask s_task[2];
  __shared__ int s_have[2];
  if (wtid == 0) {
    for (int s = 0; s < D_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 8);
    }
    mbar_fence_init();
  }
  sub_barrier(sw);
  unsigned kbase = 0;
  // scheduler state (meaningful in warp 0 of the sub-worker only, uniform across its lanes)
  DagTask hb;
  bool hb_valid = false, b_done = (P.nq == 0);
  int seg_lo = 0;
  hb.type = 1;
  auto trace_task = [&](const DagTask& t, unsigned long long t0) {
    const int slot = atomicAdd(P.trace_ctr, 1);
    if (slot < P.trace_cap) {
      unsigned long long* w = P.trace + 4 * (size_t)slot;
      w[0] = ((unsigned long long)t.type << 48) | ((unsigned long long)t.cb << 32) | ((unsigned long long)t.rb << 16) |
             ((unsigned long long)(t.nrb == 1) << 15) | ((unsigned long long)t.k0 << 8) | (unsigned long long)(t.k1 - t.k0);
      w[1] = t0;
      w[2] = globaltimer();
      w[3] = smid() * 2 + sw;
    }
  };
  // the row task in progress: rows [rt.rb, rt.rb + rt.nrb) against source column rt.cb; rt_c = the column of its next
  // sub-step (rt.cb: the panel product; then the near columns rt.k0 .. rt.k1-1, left to right)
  DagTask rt;
  bool rt_valid = false;
  int rt_c = 0;
  rt.type = 2; rt.rb = rt.nrb = rt.cb = rt.k0 = rt.k1 = 0;
  auto row_sub = [&](const DagTask& r, int c) {
    DagTask sub;
    sub.pad0 = sub.pad1 = 0;
    const int rlast = r.rb + r.nrb - 1;
    if (c == r.cb) {
      sub.type = 0; sub.rb = r.rb; sub.nrb = r.nrb; sub.cb = r.cb; sub.k0 = sub.k1 = 0;
    } else {
      sub.type = 1; sub.cb = (short)c; sub.k0 = r.cb; sub.k1 = (short)(r.cb + 1);
      sub.rb = (short)max((int)r.rb, c);                   // a block above the diagonal is not touched
      sub.nrb = (short)(rlast - sub.rb + 1);
    }
    return sub;
  };
  for (;;) {
    if (warp == 0) {
      int have = 0;
      DagTask run = hb;
      while (!have) {
        // ---- 1. the row task in progress: its next sub-step ----
        // The links of a row's dependency chain follow each other within microseconds and the chain CTA cannot run
        // faster than the rows below it, so the holder of a row task polls and takes no new work -- unless the only
        // thing its next sub-step waits for is bulk progress (the merge of the bulk buffer): then it helps with that.
        int rstate = 0;
        if (rt_valid) {
          const DagTask sub = row_sub(rt, rt_c);
          rstate = dag_state(sub, P.upd, P.fb, P.nblk, lane);
          if (rstate == 0) { run = sub; have = 2; break; }
        } else if (!hb_valid && seg_lo < P.nblk) {
          // ---- 2. claim a row task of the lowest released column that still has unclaimed ones ----
          const int pos = ld_acquire(P.chain_pos);
          if (seg_lo < pos) {
            const bool room = ld_relaxed(P.active) < P.active_cap;
            for (int s = seg_lo; s < pos && !rt_valid; ++s) {
              const int base = P.useg[s], size = P.useg[s + 1] - base;
              int tkt = size;
              const int head = ld_relaxed(P.uhead + s);
              if (!room && head >= 3 && head < size) break;     // only the tasks of the row next to the band pass the cap
              if (head < size) {
                if (lane == 0) tkt = atomicAdd(P.uhead + s, 1);
                tkt = __shfl_sync(0xffffffffu, tkt, 0);
              }
              if (tkt < size) {
                rt = P.utasks[base + tkt];
                rt_valid = true;
                rt_c = (rt.type == 3) ? rt.k0 : rt.cb;
                if (lane == 0) atomicAdd(P.active, 1);
              } else if (s == seg_lo) {
                ++seg_lo;
              }
            }
            if (rt_valid) continue;
          }
        }
        // ---- 3. the bulk slot: a ticket already held runs as soon as it is ready (whatever else this sub-worker holds:
        //      nobody ever sits on a ready task); a new ticket is taken only by a sub-worker without a row task, or one
        //      whose row task waits for bulk progress ----
        if (hb_valid && dag_ready(hb, P.upd, P.fb, P.nblk, lane)) { run = hb; hb_valid = false; have = 1; break; }
        if (!hb_valid && !b_done && (!rt_valid || rstate == 2)) {
          // Bulk queues: one per source outer block m, each ordered by target column c; a queue is released when the
          // chain has published the last column of its block.  The bulk updates of one block are applied in source
          // order, one after the other, and must all be in when the block is merged, i.e. when the chain reaches
          // column c - 1: the head with the smallest (c - 1) - (number of chunks still to follow on the same block)
          // goes first (earliest deadline, one chain step per later chunk).  Smallest c alone leaves all chunks of a
          // block to the last moment, in sequence; smallest m alone lets a backlog of far-right updates of old blocks
          // delay the columns the chain is about to reach.  Lane m inspects queue m.
          const int pos = ld_acquire(P.chain_pos);
          unsigned key = 0xffffffffu;
          bool exhausted = true;
          if (lane < P.nq) {
            const int base = P.bseg[lane], size = P.bseg[lane + 1] - base;
            const int h = ld_relaxed(P.bhead + lane);
            if (h < size) {
              exhausted = false;
              if (pos >= min(OUTER * (lane + 1), P.nblk)) {
                const int c = P.btasks[base + h].cb;
                key = ((unsigned)(2 * c - 2 * (c / OUTER - 2 - lane)) << 8) | (unsigned)lane;
              }
            }
          }
          const unsigned best = __reduce_min_sync(0xffffffffu, key);
          if (__all_sync(0xffffffffu, exhausted)) {
            b_done = true;
            continue;
          }
          if (best != 0xffffffffu) {
            const int m = (int)(best & 0xffu);
            const int base = P.bseg[m], size = P.bseg[m + 1] - base;
            int tkt = 0;
            if (lane == 0) tkt = atomicAdd(P.bhead + m, 1);
            tkt = __shfl_sync(0xffffffffu, tkt, 0);
            if (tkt < size) {
              hb = P.btasks[base + tkt];
              hb_valid = true;
            }
            continue;
          }
        }
        if (!rt_valid && !hb_valid && b_done && seg_lo >= P.nblk) { have = -1; break; }
        __nanosleep(64);
      }
      if (have == 2) {
        // bookkeeping of the row task: next sub-step, or done
        const int rlast = rt.rb + rt.nrb - 1;
        rt_c = (rt_c == rt.cb) ? rt.k0 : rt_c + 1;
        if (rt_c >= rt.k1 || rt_c > rlast) rt_valid = false;
      }
      if (lane == 0) {
        s_task[sw] = run;
        s_have[sw] = have | ((have == 2 && !rt_valid) ? 4 : 0);
      }
    }
    sub_barrier(sw);
    const int have = s_have[sw];
    if (have < 0) break;
    const DagTask t = s_task[sw];
    unsigned long long t0 = 0;
    if (P.trace && wtid == 0) t0 = globaltimer();
    if (t.nrb == 1) dag_execute<true>(t, P, &tmA, &tmB, &tmD, pipe, full, empty, kbase, sw, wtid);
    else dag_execute<false>(t, P, &tmA, &tmB, &tmD, pipe, full, empty, kbase, sw, wtid);
    if (P.trace && wtid == 0) trace_task(t, t0);
    if ((have & 4) && wtid == 0) atomicSub(P.active, 1);      // last sub-step of a row task
    sub_barrier(sw);      // s_task / s_have may be rewritten only after everybody has read them
  }
}

// ---------------------------------------------------------------------------------------------------
// Host side: the task lists
// ---------------------------------------------------------------------------------------------------
// Source-column chunks applied to target column c, in order.  Columns of the same 256-wide outer block and the
// first NEAR_EXTRA columns of the next one are updated column by column (K = 64, as soon as each source column is
// final: nothing waits at an outer-block boundary); everything further right gets one K = 256 update per outer block.
static void chunks_of(int c, int near_extra, std::vector<std::pair<int, int>>& out) {
  out.clear();
  const int mc = c / OUTER;
  for (int m = 0; m < mc; ++m) {
    if (c - (m + 1) * OUTER < near_extra) {
      for (int k = m * OUTER; k < (m + 1) * OUTER; ++k) out.emplace_back(k, k + 1);
    } else {
      out.emplace_back(m * OUTER, (m + 1) * OUTER);
    }
  }
  for (int k = mc * OUTER; k < c; ++k) out.emplace_back(k, k + 1);
}

int dag_plan_build(DagPlan& plan, int n, int extra_rows) {
  dag_plan_free(plan);
  plan.n = n;
  plan.rows = n + extra_rows;
  plan.nblk = ceil_div(n, NB);
  plan.nrb = plan.nblk + (extra_rows > 0 ? 1 : 0);
  // the columns of an outer block receive single-column updates from the sources of their own and of the previous
  // outer block (device side: dag_s0), K = 256 bulk updates from everything older
  const int near_extra = OUTER;
  // Row tasks within `near_rows` row blocks of the diagonal are one-block (64-row) tiles on all four tensor pipes: the
  // lowest latency per link of a row's dependency chain, for the rows the chain reaches soon.  Further down two-block
  // tiles do the same work in fewer, more efficient tasks (SPW_DAG_NEAR_ROWS, default: every row task is one block).
  const char* env_nr = getenv("SPW_DAG_NEAR_ROWS");
  const int near_rows = env_nr ? atoi(env_nr) : 1 << 20;
  int dev = 0;
  SPW_CUDA(cudaGetDevice(&dev));
  SPW_CUDA(cudaDeviceGetAttribute(&plan.grid, cudaDevAttrMultiProcessorCount, dev));
  const int nblk = plan.nblk, nrb = plan.nrb;
  std::vector<std::vector<DagTask>> useg(nblk);
  std::vector<DagTask> bulk;
  std::vector<std::pair<int, int>> ch;
  // bulk queues: one per source outer block m, ordered by target column, then rows (two-block tiles)
  std::vector<int> boff;
  for (int m = 0; (m + 1) * OUTER < nblk; ++m) {
    boff.push_back((int)bulk.size());
    for (int c = (m + 1) * OUTER; c < nblk; ++c) {
      chunks_of(c, near_extra, ch);
      for (auto& kk : ch)
        if (kk.second - kk.first > 1 && kk.first == m * OUTER) {
          DagTask t;
          memset(&t, 0, sizeof(t));
          t.type = 1; t.cb = (short)c; t.k0 = (short)kk.first; t.k1 = (short)kk.second;
          // the three blocks next to the diagonal are what the chain CTA waits for: one-block tiles (half the time of a
          // two-block tile), first in the column
          int b = c;
          for (; b < std::min(nrb, c + 3); ++b) {
            t.rb = (short)b; t.nrb = 1;
            bulk.push_back(t);
          }
          for (; b < nrb; b += 2) {
            t.rb = (short)b; t.nrb = (short)std::min(2, nrb - b);
            bulk.push_back(t);
          }
        }
    }
  }
  boff.push_back((int)bulk.size());
  plan.nq = (int)boff.size() - 1;
  if (plan.nq > 32) { set_error("potrf_dag: order %d too large (more than 32 outer blocks)", n); return SPW_ERR_ARG; }
  // row tasks: segment j = one task per row tile below the band of the chain CTA (row blocks >= j + 3), from the
  // diagonal downwards.  A row task does the panel product of its rows against column j and then applies source j
  // to its rows of the near columns j+1 .. near_end(j) (those that receive single-column updates from j, see
  // chunks_of), left to right.
  for (int j = 0; j < nblk; ++j) {
    const int near_end = std::min(nblk - 1, (j / OUTER) * OUTER + OUTER - 1 + near_extra);
    int b = j + 3;
    if (b < nrb) {
      // the row block next to the band (the chain CTA consumes it two steps later): its three updates run in parallel
      // on three sub-workers (types 3: updates only; they start when the panel product of the first task is out)
      DagTask t;
      memset(&t, 0, sizeof(t));
      t.cb = (short)j; t.rb = (short)b; t.nrb = 1;
      for (int c = j + 1; c <= std::min(near_end, b); ++c) {
        t.type = (short)(c == j + 1 ? 2 : 3);
        t.k0 = (short)c; t.k1 = (short)(c + 1);
        useg[j].push_back(t);
      }
      if (near_end < j + 1) {      // no near column at all (last column): the panel product alone
        t.type = 2; t.k0 = (short)(j + 1); t.k1 = (short)(j + 1);
        useg[j].push_back(t);
      }
      ++b;
    }
    while (b < nrb) {
      // two-block tiles start at an even distance from column j so that the same pairs recur in every segment of
      // their life (a block must not change partner while updates of the pair are in flight in the same column)
      const int nb = (b - j - 3 >= near_rows && ((b & 1) == 0) && b + 1 < nrb) ? 2 : 1;
      DagTask t;
      memset(&t, 0, sizeof(t));
      t.type = 2; t.cb = (short)j; t.rb = (short)b; t.nrb = (short)nb;
      t.k0 = (short)(j + 1); t.k1 = (short)(near_end + 1);
      useg[j].push_back(t);
      b += nb;
    }
  }
  std::vector<DagTask> ut;
  std::vector<int> off(nblk + 1, 0);
  for (int j = 0; j < nblk; ++j) {
    off[j] = (int)ut.size();
    ut.insert(ut.end(), useg[j].begin(), useg[j].end());
  }
  off[nblk] = (int)ut.size();
  plan.n_urgent = (int)ut.size();
  plan.n_bulk = (int)bulk.size();
  SPW_CUDA(cudaMalloc(&plan.d_utasks, sizeof(DagTask) * std::max<size_t>(1, ut.size())));
  SPW_CUDA(cudaMalloc(&plan.d_btasks, sizeof(DagTask) * std::max<size_t>(1, bulk.size())));
  SPW_CUDA(cudaMalloc(&plan.d_useg, sizeof(int) * (nblk + 1)));
  if (!ut.empty()) SPW_CUDA(cudaMemcpy(plan.d_utasks, ut.data(), sizeof(DagTask) * ut.size(), cudaMemcpyHostToDevice));
  if (!bulk.empty()) SPW_CUDA(cudaMemcpy(plan.d_btasks, bulk.data(), sizeof(DagTask) * bulk.size(), cudaMemcpyHostToDevice));
  SPW_CUDA(cudaMemcpy(plan.d_useg, off.data(), sizeof(int) * (nblk + 1), cudaMemcpyHostToDevice));
  SPW_CUDA(cudaMalloc(&plan.d_bseg, sizeof(int) * boff.size()));
  SPW_CUDA(cudaMemcpy(plan.d_bseg, boff.data(), sizeof(int) * boff.size(), cudaMemcpyHostToDevice));
  plan.sync_ints = 2 * (size_t)nrb * nblk + nblk + plan.nq + 4;
  plan.ld = round_up_ll(n, 16);
  SPW_CUDA(cudaMalloc(&plan.d_bulk, sizeof(double) * (size_t)plan.rows * plan.ld));
  // at most this many row tasks run at once: the other sub-workers stay available for the bulk updates the row
  // tasks may be waiting for (deadlock freedom, see the header comment)
  plan.active_cap = std::max(1, 2 * (plan.grid - 1) * 2 / 3);
  SPW_CUDA(cudaMalloc(&plan.d_sync, sizeof(int) * plan.sync_ints));
  const char* env_tr = getenv("SPW_DAG_TRACE");
  if (env_tr && env_tr[0] == '1') {
    plan.trace_cap = plan.n_urgent * (2 + OUTER + near_extra) + plan.n_bulk + nblk + 16;
    SPW_CUDA(cudaMalloc(&plan.d_trace, sizeof(unsigned long long) * 4 * plan.trace_cap));
  }
  static bool attr_set[64] = {false};
  if (!attr_set[dev & 63]) {
    SPW_CUDA(cudaFuncSetAttribute(potrf_dag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D_SMEM));
    attr_set[dev & 63] = true;
  }
  return SPW_OK;
}

void dag_plan_free(DagPlan& plan) {
  if (plan.d_utasks) cudaFree(plan.d_utasks);
  if (plan.d_btasks) cudaFree(plan.d_btasks);
  if (plan.d_useg) cudaFree(plan.d_useg);
  if (plan.d_bseg) cudaFree(plan.d_bseg);
  if (plan.d_sync) cudaFree(plan.d_sync);
  if (plan.d_bulk) cudaFree(plan.d_bulk);
  if (plan.d_trace) cudaFree(plan.d_trace);
  plan = DagPlan();
}

int potrf_dag(const DagPlan& plan, MatView L, double* dinv, int* status, cudaStream_t st) {
  MatView dv{dinv, NB, (long long)plan.nblk * NB * NB};
  CUtensorMap tmA, tmB, tmD;
  SPW_TRY(make_tensor_map(&tmA, L, plan.n, plan.rows, 1, D_LDS, D_BM));
  SPW_TRY(make_tensor_map(&tmB, L, plan.n, plan.rows, 1, D_LDS, D_BN));
  SPW_TRY(make_tensor_map(&tmD, dv, NB, plan.nblk * NB, 1, D_LDS, D_BN));
  SPW_CUDA(cudaMemsetAsync(plan.d_sync, 0, sizeof(int) * plan.sync_ints, st));
  DagParams P;
  P.utasks = plan.d_utasks;
  P.useg = plan.d_useg;
  P.btasks = plan.d_btasks;
  P.bseg = plan.d_bseg;
  P.nq = plan.nq;
  if (L.ld != plan.ld) { set_error("potrf_dag: leading dimension %lld, plan expects %lld", L.ld, plan.ld); return SPW_ERR_ARG; }
  P.upd = plan.d_sync;
  P.fb = plan.d_sync + (size_t)plan.nrb * plan.nblk;
  P.Bp = plan.d_bulk;
  P.uhead = P.fb + (size_t)plan.nrb * plan.nblk;
  P.bhead = P.uhead + plan.nblk;
  P.chain_pos = P.bhead + plan.nq;
  P.trace_ctr = P.chain_pos + 1;
  P.active = P.chain_pos + 2;
  P.active_cap = plan.active_cap;
  P.L = L;
  P.dinv = dinv;
  P.n = plan.n;
  P.rows = plan.rows;
  P.nblk = plan.nblk;
  P.nrb = plan.nrb;
  P.status = status;
  P.trace = plan.d_trace;
  P.trace_cap = plan.trace_cap;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(plan.grid);
  cfg.blockDim = dim3(D_THREADS);
  cfg.dynamicSmemBytes = D_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;      // all CTAs resident at once: the workers spin on flags
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SPW_CUDA(cudaLaunchKernelEx(&cfg, potrf_dag_kernel, tmA, tmB, tmD, P));
  SPW_LAUNCHED();
  return SPW_OK;
}

int dag_trace_read(const DagPlan& plan, std::vector<unsigned long long>& out) {
  out.clear();
  if (!plan.d_trace) return SPW_OK;
  int cnt = 0;
  SPW_CUDA(cudaMemcpy(&cnt, plan.d_sync + 2 * (size_t)plan.nrb * plan.nblk + plan.nblk + plan.nq + 1, sizeof(int),
                      cudaMemcpyDeviceToHost));
  cnt = std::min(cnt, plan.trace_cap);
  out.resize(4 * (size_t)cnt);
  if (cnt > 0) SPW_CUDA(cudaMemcpy(out.data(), plan.d_trace, sizeof(unsigned long long) * 4 * cnt, cudaMemcpyDeviceToHost));
  return SPW_OK;
}

}  // namespace spw
