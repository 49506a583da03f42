// 64 x 64 tile products on the FP64 tensor pipe straight from shared memory (row-major tiles with leading dimension 68:
// conflict-free DMMA fragment loads).  Shared by the chain CTA of the persistent tile-DAG Cholesky (potrf_dag.cu) and
// the look-ahead diagonal kernel of the multi-launch schedule (potrf_la.cu).
#pragma once
#include "common.cuh"
#include "diag_factor.cuh"

namespace spw {

constexpr int C_LD = DIAG_SRC_LD;                 // 68
constexpr int C_TILE = NB * C_LD;

// dst = src (- sub, when the block has a bulk accumulator to merge)
template <int NT>
__device__ __forceinline__ void chain_load_tile(double* __restrict__ dst, const double* __restrict__ src,
                                                const double* __restrict__ sub, long long ld, int nr, int nc, int t) {
  // all loads of a batch are issued before the first use: a few L2 round trips per tile, not one per element
  constexpr int PER = NB * (NB / 2) / NT, BATCH = PER < 8 ? PER : 8;
#pragma unroll
  for (int i0 = 0; i0 < PER; i0 += BATCH) {
    double2 v[BATCH], w[BATCH];
#pragma unroll
    for (int i = 0; i < BATCH; ++i) {
      const int e = t + (i0 + i) * NT, r = e >> 5, c = (e & 31) * 2;
      v[i] = w[i] = make_double2(0.0, 0.0);
      if (r < nr) {
        const long long o = (long long)r * ld + c;
        if (c + 1 < nc) {
          v[i] = __ldcg(reinterpret_cast<const double2*>(src + o));
          if (sub) w[i] = __ldcg(reinterpret_cast<const double2*>(sub + o));
        } else if (c < nc) {
          v[i].x = __ldcg(src + o);
          if (sub) w[i].x = __ldcg(sub + o);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < BATCH; ++i) {
      const int e = t + (i0 + i) * NT, r = e >> 5, c = (e & 31) * 2;
      *reinterpret_cast<double2*>(dst + r * C_LD + c) = make_double2(v[i].x - w[i].x, v[i].y - w[i].y);
    }
  }
}

// Dinv[8 nt + g][4 k4 + q] from the packed 4 x 4 tiles of the sweep (lower triangular: zero above the diagonal)
__device__ __forceinline__ double chain_dinv_frag(const double* __restrict__ X, int nt, int k4, int g, int q) {
  const int tr = 2 * nt + (g >> 2);
  return (tr >= k4) ? X[diag_tile(tr, k4) * DIAG_TS + 4 * (g & 3) + q] : 0.0;
}

// rows 8 rg .. 8 rg + 7 of T * Dinv^T, written back in place and to global memory (bounded by nr x nc)
__device__ __forceinline__ void chain_panel_rows(double* __restrict__ T, const double* __restrict__ X, int rg,
                                                 double* __restrict__ gdst, long long ld, int nr, int nc, int lane) {
  const int g = lane >> 2, q = lane & 3;
  double acc[8][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
  const double* arow = T + (8 * rg + g) * C_LD + q;
#pragma unroll
  for (int k4 = 0; k4 < 16; ++k4) {
    const double a = arow[4 * k4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      if (nt >= (k4 >> 1)) dmma884(acc[nt][0], acc[nt][1], a, chain_dinv_frag(X, nt, k4, g, q));
  }
  __syncwarp();                      // every lane has read its part of the rows before they are overwritten
  const int r = 8 * rg + g;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int c = 8 * nt + 2 * q;
    *reinterpret_cast<double2*>(T + r * C_LD + c) = make_double2(acc[nt][0], acc[nt][1]);
    if (r < nr) {
      if (c + 1 < nc) __stcg(reinterpret_cast<double2*>(gdst + (long long)r * ld + c), make_double2(acc[nt][0], acc[nt][1]));
      else if (c < nc) __stcg(gdst + (long long)r * ld + c, acc[nt][0]);
    }
  }
}

// rows 8 rg .. 8 rg + 7 of C -= A B^T (all 64 columns, K = 64)
__device__ __forceinline__ void chain_update_rows(double* __restrict__ C, const double* __restrict__ A,
                                                  const double* __restrict__ B, int rg, int lane) {
  const int g = lane >> 2, q = lane & 3;
  double acc[8][2];
  double* crow = C + (8 * rg + g) * C_LD + 2 * q;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const double2 v = *reinterpret_cast<const double2*>(crow + 8 * nt);
    acc[nt][0] = v.x; acc[nt][1] = v.y;
  }
  const double* arow = A + (8 * rg + g) * C_LD + q;
  const double* brow = B + g * C_LD + q;
#pragma unroll
  for (int k4 = 0; k4 < 16; ++k4) {
    const double a = -arow[4 * k4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) dmma884(acc[nt][0], acc[nt][1], a, brow[8 * nt * C_LD + 4 * k4]);
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<double2*>(crow + 8 * nt) = make_double2(acc[nt][0], acc[nt][1]);
}

// 8 x 8 tile t (0 .. 35, lower triangle of the 8 x 8 tile grid, row-wise) of S -= X X^T
__device__ __forceinline__ void chain_syrk_tile(double* __restrict__ S, const double* __restrict__ X, int t, int lane) {
  const int g = lane >> 2, q = lane & 3;
  int mt = 0;
  while (t > mt) { t -= mt + 1; ++mt; }
  const int nt = t;
  double* cp = S + (8 * mt + g) * C_LD + 8 * nt + 2 * q;
  double2 v = *reinterpret_cast<const double2*>(cp);
  const double* arow = X + (8 * mt + g) * C_LD + q;
  const double* brow = X + (8 * nt + g) * C_LD + q;
#pragma unroll
  for (int k4 = 0; k4 < 16; ++k4) dmma884(v.x, v.y, -arow[4 * k4], brow[4 * k4]);
  *reinterpret_cast<double2*>(cp) = v;
}

}  // namespace spw
