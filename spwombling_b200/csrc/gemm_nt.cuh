// NT DMMA tile engine: C[m][n] = sum_k A[m][k] * B[n][k] with both operands k-contiguous in global
// memory, staged through shared memory with a multi-stage cp.async (LDGSTS) ring and consumed by
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4, the FP64 tensor pipe of sm_100a).
//
// Shared-memory layout: row-major tiles with leading dimension BK + 4 doubles (== 4 mod 16), which
// makes the 64-bit fragment loads (lane -> row lane>>2, k lane&3) conflict-free per half warp.
#pragma once
#include "common.cuh"

namespace spw {

template <int BM_, int BN_, int BK_, int STAGES_, int WARPS_M_, int WARPS_N_>
struct NtCfg {
  static constexpr int BM = BM_, BN = BN_, BK = BK_, STAGES = STAGES_;
  static constexpr int WARPS_M = WARPS_M_, WARPS_N = WARPS_N_;
  static constexpr int NWARPS = WARPS_M * WARPS_N, NTHREADS = 32 * NWARPS;
  static constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
  static constexpr int MF = WM / 8, NF = WN / 8;
  static constexpr int LDS = BK + 4;
  static constexpr int A_STAGE = BM * LDS, B_STAGE = BN * LDS;   // doubles per stage
  static constexpr int STAGE = A_STAGE + B_STAGE;
  static constexpr size_t PIPE_BYTES = size_t(STAGES) * STAGE * sizeof(double);
  static_assert(BM % (8 * WARPS_M) == 0 && BN % (8 * WARPS_N) == 0, "warp tiling");
  static_assert(BK % 4 == 0 && (LDS % 16) == 4, "smem leading dimension must be 4 mod 16");
};

// Issue the cp.async copies of one operand tile: `rows` x BK doubles starting at g (row stride ld,
// doubles), rows >= rows_valid and k >= k_valid are zero-filled.  g must be 16-byte aligned and ld even.
template <int ROWS, int BK, int LDS, int NTHREADS>
__device__ __forceinline__ void load_tile_async(double* s, const double* g, long long ld, int rows_valid,
                                                int k_valid, int tid) {
  constexpr int CH = BK / 2;              // 16-byte chunks per row
  constexpr int TOTAL = ROWS * CH;
#pragma unroll
  for (int it = 0; it < (TOTAL + NTHREADS - 1) / NTHREADS; ++it) {
    const int idx = tid + it * NTHREADS;
    if ((TOTAL % NTHREADS) != 0 && idx >= TOTAL) break;
    const int r = idx / CH, c = idx % CH;
    const int kleft = k_valid - 2 * c;
    int bytes = (r < rows_valid) ? (kleft >= 2 ? 16 : (kleft == 1 ? 8 : 0)) : 0;
    const double* src = bytes ? (g + r * ld + 2 * c) : g;
    cp_async16(s + r * LDS + 2 * c, src, bytes);
  }
}

// Triangular masks applied at fragment-load time on tiles that straddle the diagonal.
//   a_tri = 1: keep A[m][k] iff k <= a_lim0 + m          (lower-triangular A)
//   b_tri = 1: keep B[n][k] iff k <= b_lim0 + n          (lower-triangular B)
//   b_tri = 2: keep B[n][k] iff k >= b_lim0 + n          (upper-triangular B)
// m, n are tile-local rows; k is the k index local to the task (0 .. k_total).
struct TriMask {
  int a_tri, a_lim0, b_tri, b_lim0;
};

template <class Cfg, bool MASKED>
__device__ __forceinline__ void compute_ktile(const double* __restrict__ As, const double* __restrict__ Bs,
                                              double (&acc)[Cfg::MF][Cfg::NF][2], int warp_m, int warp_n,
                                              int lane, int k0, const TriMask& tm) {
  const int g = lane >> 2, q = lane & 3;
  const double* a_base = As + (warp_m * Cfg::WM + g) * Cfg::LDS + q;
  const double* b_base = Bs + (warp_n * Cfg::WN + g) * Cfg::LDS + q;
#pragma unroll
  for (int kk = 0; kk < Cfg::BK / 4; ++kk) {
    double a[Cfg::MF], b[Cfg::NF];
#pragma unroll
    for (int i = 0; i < Cfg::MF; ++i) a[i] = a_base[i * 8 * Cfg::LDS + kk * 4];
#pragma unroll
    for (int j = 0; j < Cfg::NF; ++j) b[j] = b_base[j * 8 * Cfg::LDS + kk * 4];
    if (MASKED) {
      const int k = k0 + kk * 4 + q;
      if (tm.a_tri == 1) {
#pragma unroll
        for (int i = 0; i < Cfg::MF; ++i)
          if (k > tm.a_lim0 + warp_m * Cfg::WM + i * 8 + g) a[i] = 0.0;
      }
      if (tm.b_tri == 1) {
#pragma unroll
        for (int j = 0; j < Cfg::NF; ++j)
          if (k > tm.b_lim0 + warp_n * Cfg::WN + j * 8 + g) b[j] = 0.0;
      } else if (tm.b_tri == 2) {
#pragma unroll
        for (int j = 0; j < Cfg::NF; ++j)
          if (k < tm.b_lim0 + warp_n * Cfg::WN + j * 8 + g) b[j] = 0.0;
      }
    }
#pragma unroll
    for (int i = 0; i < Cfg::MF; ++i)
#pragma unroll
      for (int j = 0; j < Cfg::NF; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}

// Does the k-tile [k0, k0 + BK) need per-element masking?
template <class Cfg>
__device__ __forceinline__ bool ktile_needs_mask(int k0, const TriMask& tm) {
  bool need = false;
  if (tm.a_tri == 1) need |= (k0 + Cfg::BK - 1 > tm.a_lim0);
  if (tm.b_tri == 1) need |= (k0 + Cfg::BK - 1 > tm.b_lim0);
  if (tm.b_tri == 2) need |= (k0 < tm.b_lim0 + Cfg::BN - 1);
  return need;
}

// Full mainloop for one output tile: acc += A[0:m_valid][0:k_total] * B[0:n_valid][0:k_total]^T.
// All threads of the CTA must call it; smem must hold Cfg::PIPE_BYTES.
template <class Cfg>
__device__ __forceinline__ void nt_mainloop(const double* __restrict__ A, long long lda, int m_valid,
                                            const double* __restrict__ B, long long ldb, int n_valid,
                                            int k_total, const TriMask& tm, double* smem,
                                            double (&acc)[Cfg::MF][Cfg::NF][2]) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_m = warp % Cfg::WARPS_M, warp_n = warp / Cfg::WARPS_M;
  const int nkt = (k_total + Cfg::BK - 1) / Cfg::BK;
#pragma unroll
  for (int s = 0; s < Cfg::STAGES - 1; ++s) {
    if (s < nkt) {
      double* st = smem + s * Cfg::STAGE;
      load_tile_async<Cfg::BM, Cfg::BK, Cfg::LDS, Cfg::NTHREADS>(st, A + s * Cfg::BK, lda, m_valid,
                                                                 k_total - s * Cfg::BK, tid);
      load_tile_async<Cfg::BN, Cfg::BK, Cfg::LDS, Cfg::NTHREADS>(st + Cfg::A_STAGE, B + s * Cfg::BK, ldb, n_valid,
                                                                 k_total - s * Cfg::BK, tid);
    }
    cp_async_commit();
  }
  for (int kt = 0; kt < nkt; ++kt) {
    cp_async_wait<Cfg::STAGES - 2>();
    __syncthreads();
    const int nx = kt + Cfg::STAGES - 1;
    if (nx < nkt) {
      double* st = smem + (nx % Cfg::STAGES) * Cfg::STAGE;
      load_tile_async<Cfg::BM, Cfg::BK, Cfg::LDS, Cfg::NTHREADS>(st, A + nx * Cfg::BK, lda, m_valid,
                                                                 k_total - nx * Cfg::BK, tid);
      load_tile_async<Cfg::BN, Cfg::BK, Cfg::LDS, Cfg::NTHREADS>(st + Cfg::A_STAGE, B + nx * Cfg::BK, ldb, n_valid,
                                                                 k_total - nx * Cfg::BK, tid);
    }
    cp_async_commit();
    const double* st = smem + (kt % Cfg::STAGES) * Cfg::STAGE;
    const int k0 = kt * Cfg::BK;
    if (ktile_needs_mask<Cfg>(k0, tm))
      compute_ktile<Cfg, true>(st, st + Cfg::A_STAGE, acc, warp_m, warp_n, lane, k0, tm);
    else
      compute_ktile<Cfg, false>(st, st + Cfg::A_STAGE, acc, warp_m, warp_n, lane, k0, tm);
  }
  cp_async_wait<0>();
  __syncthreads();
}

}  // namespace spw
