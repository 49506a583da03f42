// Batched blocked Cholesky (right-looking, NB = 64 diagonal blocks, 128x128 DMMA trailing tiles) and
// level-synchronous divide-and-conquer triangular inverse, both driven by precomputed tile task lists.
// Replaces the LAPACK calls the reference reaches through base R:
//   chol()      R/hlmBayes_sp.R:99,114,140,181; R/spatial_gradient.R:216,265,304   (dpotrf)
//   chol2inv()  same lines (dpotri) -- not formed: K^-1 = Linv^T Linv is applied through Linv only.
#include <algorithm>
#include <cstring>
#include "linalg.cuh"
#include "gemm_nt.cuh"

namespace spw {

using GemmCfg = NtCfg<128, 128, 16, 4, 4, 2>;   // 8 warps, warp tile 32 x 64

// ---------------------------------------------------------------------------------------------------
// Task-driven NT GEMM: one CTA per (task, batch entry).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GemmCfg::NTHREADS, 1)
gemm_tasks_kernel(const GemmTask* __restrict__ tasks, MatView A, MatView B, MatView C, MatView CT, double alpha) {
  extern __shared__ __align__(16) double smem[];
  const GemmTask t = tasks[blockIdx.x];
  const long long b = blockIdx.y;
  const double* Ap = A.p + b * A.stride + (long long)t.a_row * A.ld + t.a_col;
  const double* Bp = B.p + b * B.stride + (long long)t.b_row * B.ld + t.b_col;
  double acc[GemmCfg::MF][GemmCfg::NF][2];
#pragma unroll
  for (int i = 0; i < GemmCfg::MF; ++i)
#pragma unroll
    for (int j = 0; j < GemmCfg::NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  TriMask tm{t.a_tri, t.a_lim0, t.b_tri, t.b_lim0};
  nt_mainloop<GemmCfg>(Ap, A.ld, t.m, Bp, B.ld, t.n, t.k, tm, smem, acc);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_m = warp % GemmCfg::WARPS_M, warp_n = warp / GemmCfg::WARPS_M;
  const int g = lane >> 2, q = lane & 3;
  double* Cp = C.p + b * C.stride + (long long)t.c_row * C.ld + t.c_col;
  double* CTp = (t.flags & GT_STORE_T) ? CT.p + b * CT.stride + (long long)t.ct_row * CT.ld + t.ct_col : nullptr;
#pragma unroll
  for (int i = 0; i < GemmCfg::MF; ++i) {
    const int r = warp_m * GemmCfg::WM + i * 8 + g;
    if (r >= t.m) continue;
#pragma unroll
    for (int j = 0; j < GemmCfg::NF; ++j) {
      const int c = warp_n * GemmCfg::WN + j * 8 + 2 * q;
      if (c >= t.n) continue;
      double v0 = alpha * acc[i][j][0], v1 = alpha * acc[i][j][1];
      const bool two = (c + 1 < t.n);
      double* dst = Cp + (long long)r * C.ld + c;
      if (t.flags & GT_ACCUM) {
        if (two) {
          const double2 old = *reinterpret_cast<const double2*>(dst);
          v0 += old.x;
          v1 += old.y;
        } else {
          v0 += dst[0];
        }
      }
      if (t.flags & GT_STORE) {
        if (two) *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
        else dst[0] = v0;
      }
      if (CTp) {
        CTp[(long long)c * CT.ld + r] = v0;
        if (two) CTp[(long long)(c + 1) * CT.ld + r] = v1;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Diagonal block: factor the bs x bs lower block in shared memory (one thread per row, left-looking by
// columns), write it back, invert the factor, and publish the inverse.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NB, 1)
diag_factor_kernel(MatView L, int blk, int n, double* __restrict__ dinv, int nblk, MatView linv, int has_linv,
                   int* __restrict__ status) {
  extern __shared__ __align__(16) double diag_smem[];
  double (*Ls)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(diag_smem);
  double (*Xs)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(diag_smem + NB * (NB + 1));
  __shared__ int bad;
  const int b = blockIdx.x, i = threadIdx.x;
  const int r0 = blk * NB, bs = min(NB, n - r0);
  double* Lp = L.p + (long long)b * L.stride + (long long)r0 * L.ld + r0;
  if (i == 0) bad = 0;
  // load the lower block (thread i loads row i)
  for (int c = 0; c < NB; ++c) {
    // coalesce over columns: thread i reads column i of row c
    double v = 0.0;
    if (c < bs && i < bs && i <= c) v = Lp[(long long)c * L.ld + i];
    Ls[c][i] = v;
  }
  __syncthreads();
  // left-looking column Cholesky: thread i owns row i
  for (int j = 0; j < bs; ++j) {
    double s0 = 0.0, s1 = 0.0;
    if (i >= j && i < bs) {
      int k = 0;
      for (; k + 1 < j; k += 2) {
        s0 = fma(Ls[i][k], Ls[j][k], s0);
        s1 = fma(Ls[i][k + 1], Ls[j][k + 1], s1);
      }
      if (k < j) s0 = fma(Ls[i][k], Ls[j][k], s0);
    }
    const double s = (i >= j && i < bs) ? Ls[i][j] - (s0 + s1) : 0.0;
    __syncthreads();
    if (i == j) {
      if (!(s > 0.0)) {
        if (bad == 0) bad = r0 + j + 1;
        Ls[j][j] = nan("");
      } else {
        Ls[j][j] = sqrt(s);
      }
    }
    __syncthreads();
    if (i > j && i < bs) Ls[i][j] = s / Ls[j][j];
    __syncthreads();
  }
  if (i == 0 && bad) atomicCAS(&status[b], 0, bad);
  // write the factor back (coalesced over columns)
  for (int c = 0; c < bs; ++c)
    if (i <= c && i < bs) Lp[(long long)c * L.ld + i] = Ls[c][i];
  // inverse: thread c solves L x = e_c (column c of the inverse), rows in sequence
  for (int r = 0; r < NB; ++r) Xs[r][i] = 0.0;
  if (i < bs) {
    for (int r = i; r < bs; ++r) {
      double s0 = (r == i) ? 1.0 : 0.0, s1 = 0.0;
      int k = i;
      for (; k + 1 < r; k += 2) {
        s0 = fma(-Ls[r][k], Xs[k][i], s0);
        s1 = fma(-Ls[r][k + 1], Xs[k + 1][i], s1);
      }
      if (k < r) s0 = fma(-Ls[r][k], Xs[k][i], s0);
      Xs[r][i] = (s0 + s1) / Ls[r][r];
    }
  }
  __syncthreads();
  // publish: dinv block [NB x NB] row-major with explicit zeros above the diagonal / beyond bs
  double* dp = dinv + ((long long)b * nblk + blk) * (NB * NB);
  for (int r = 0; r < NB; ++r) dp[r * NB + i] = Xs[r][i];
  if (has_linv) {
    double* ip = linv.p + (long long)b * linv.stride + (long long)r0 * linv.ld + r0;
    for (int r = 0; r < bs; ++r)
      if (i < bs) ip[(long long)r * linv.ld + i] = (i <= r) ? Xs[r][i] : Xs[i][r];
  }
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
__global__ void __launch_bounds__(256)
logdet_quad_kernel(int n, MatView L, double* __restrict__ logdet_half, double* __restrict__ quad) {
  __shared__ double red[2][8];
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
    for (int w = 0; w < 8; ++w) {
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

int plan_build(LinalgPlan& plan, int n, int extra_rows) {
  plan_free(plan);
  plan.n = n;
  plan.rows = n + extra_rows;
  plan.nblk = ceil_div(n, NB);
  std::vector<GemmTask> tasks;
  const int T = GemmCfg::BM;   // 128
  plan.panel.resize(plan.nblk);
  plan.trail.resize(plan.nblk);
  for (int j = 0; j < plan.nblk; ++j) {
    const int c0 = j * NB, bs = std::min(NB, n - c0), r_begin = c0 + bs;
    // panel: rows below the diagonal block (including appended rows): L_ij = A_ij * Dinv_j^T
    plan.panel[j].begin = (int)tasks.size();
    for (int r = r_begin; r < plan.rows; r += T) {
      GemmTask t = blank_task();
      t.a_row = r; t.a_col = c0;
      t.b_row = j * NB; t.b_col = 0;          // in the dinv view (ld = NB)
      t.c_row = r; t.c_col = c0;
      t.m = std::min(T, plan.rows - r); t.n = bs; t.k = bs;
      t.flags = GT_STORE;
      tasks.push_back(t);
    }
    plan.panel[j].count = (int)tasks.size() - plan.panel[j].begin;
    // trailing update: C[I][K] -= L[I][c0:c0+bs] * L[K][c0:c0+bs]^T for the lower triangle of the rest
    plan.trail[j].begin = (int)tasks.size();
    for (int r = r_begin; r < plan.rows; r += T) {
      const int m = std::min(T, plan.rows - r);
      for (int c = r_begin; c < n && c <= r + m - 1; c += T) {
        GemmTask t = blank_task();
        t.a_row = r; t.a_col = c0;
        t.b_row = c; t.b_col = c0;
        t.c_row = r; t.c_col = c;
        t.m = m; t.n = std::min(T, n - c); t.k = bs;
        t.flags = GT_ACCUM | GT_STORE;
        tasks.push_back(t);
      }
    }
    plan.trail[j].count = (int)tasks.size() - plan.trail[j].begin;
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

static int launch_gemm_tasks(const GemmTask* tasks, int count, int batch, MatView A, MatView B, MatView C, MatView CT,
                             double alpha, cudaStream_t st) {
  if (count <= 0 || batch <= 0) return SPW_OK;
  static bool attr_set[64] = {false};
  int dev = 0;
  SPW_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    SPW_CUDA(cudaFuncSetAttribute(gemm_tasks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)GemmCfg::PIPE_BYTES));
    attr_set[dev & 63] = true;
  }
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = std::min(65535, batch - b0);
    MatView a = A, b = B, c = C, ct = CT;
    a.p += (long long)b0 * A.stride;
    b.p += (long long)b0 * B.stride;
    c.p += (long long)b0 * C.stride;
    if (ct.p) ct.p += (long long)b0 * CT.stride;
    gemm_tasks_kernel<<<dim3(count, nb), GemmCfg::NTHREADS, GemmCfg::PIPE_BYTES, st>>>(tasks, a, b, c, ct, alpha);
    SPW_LAUNCHED();
  }
  return SPW_OK;
}

int potrf_batched(const LinalgPlan& plan, MatView L, double* dinv, MatView* linv, int batch, int* status,
                  cudaStream_t st) {
  MatView dv{dinv, NB, (long long)plan.nblk * NB * NB};
  MatView none{nullptr, 0, 0};
  MatView lv = linv ? *linv : none;
  constexpr size_t DIAG_SMEM = 2 * NB * (NB + 1) * sizeof(double);
  static bool diag_attr[64] = {false};
  int dev = 0;
  SPW_CUDA(cudaGetDevice(&dev));
  if (!diag_attr[dev & 63]) {
    SPW_CUDA(cudaFuncSetAttribute(diag_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM));
    diag_attr[dev & 63] = true;
  }
  for (int j = 0; j < plan.nblk; ++j) {
    diag_factor_kernel<<<batch, NB, DIAG_SMEM, st>>>(L, j, plan.n, dinv, plan.nblk, lv, linv ? 1 : 0, status);
    SPW_LAUNCHED();
    SPW_TRY(launch_gemm_tasks(plan.d_tasks + plan.panel[j].begin, plan.panel[j].count, batch, L, dv, L, none, 1.0, st));
    SPW_TRY(launch_gemm_tasks(plan.d_tasks + plan.trail[j].begin, plan.trail[j].count, batch, L, L, L, none, -1.0, st));
  }
  return SPW_OK;
}

int trtri_batched(const LinalgPlan& plan, MatView L, MatView linv, int batch, cudaStream_t st) {
  MatView none{nullptr, 0, 0};
  for (size_t lev = 0; lev < plan.inv1.size(); ++lev) {
    SPW_TRY(launch_gemm_tasks(plan.d_tasks + plan.inv1[lev].begin, plan.inv1[lev].count, batch, L, linv, none, L, 1.0, st));
    SPW_TRY(launch_gemm_tasks(plan.d_tasks + plan.inv2[lev].begin, plan.inv2[lev].count, batch, linv, L, linv, linv, -1.0, st));
  }
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
  logdet_quad_kernel<<<batch, 256, 0, st>>>(n, L, logdet_half, quad);
  SPW_LAUNCHED();
  return SPW_OK;
}

}  // namespace spw
