// Diagonal-block factor + inverse sweep (device function shared by diag_factor_kernel and the persistent tile-DAG
// Cholesky).  See the comment block in linalg.cu for the warp roles.
#pragma once
#include "common.cuh"
#include "linalg.cuh"

namespace spw {

constexpr int DIAG_NT = NB / 4;                              // 16 tile rows
constexpr int DIAG_LOW = DIAG_NT * (DIAG_NT + 1) / 2;        // 136 lower tiles
constexpr int DIAG_OFF = DIAG_NT * (DIAG_NT - 1) / 2;        // 120 off-diagonal tiles
constexpr int DIAG_THREADS = 384;                            // 12 warps; warps 4 and 8 only take part in the barriers
constexpr int DIAG_HELPER = 4 * 32;                          // a lane of an otherwise idle warp
constexpr int DIAG_TS = 18;                                  // tile stride in doubles

__device__ __forceinline__ int diag_tile(int i, int k) { return DIAG_NT * k - (k * (k - 1)) / 2 + (i - k); }

__device__ __forceinline__ void tile_load(const double* __restrict__ p, double (&t)[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const double2 u0 = *reinterpret_cast<const double2*>(p + 4 * r);
    const double2 u1 = *reinterpret_cast<const double2*>(p + 4 * r + 2);
    t[r][0] = u0.x; t[r][1] = u0.y; t[r][2] = u1.x; t[r][3] = u1.y;
  }
}

__device__ __forceinline__ void tile_store(double* __restrict__ p, const double (&t)[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    *reinterpret_cast<double2*>(p + 4 * r) = make_double2(t[r][0], t[r][1]);
    *reinterpret_cast<double2*>(p + 4 * r + 2) = make_double2(t[r][2], t[r][3]);
  }
}

// c -= a b^T
__device__ __forceinline__ void tile_sub_abt(double (&c)[4][4], const double (&a)[4][4], const double (&b)[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double acc = c[r][q];
#pragma unroll
      for (int m = 0; m < 4; ++m) acc = fma(-a[r][m], b[q][m], acc);
      c[r][q] = acc;
    }
}

// c += a b
__device__ __forceinline__ void tile_add_ab(double (&c)[4][4], const double (&a)[4][4], const double (&b)[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double acc = c[r][q];
#pragma unroll
      for (int m = 0; m < 4; ++m) acc = fma(a[r][m], b[m][q], acc);
      c[r][q] = acc;
    }
}

// 4 x 4 tile at (row0, col0) of a row-major matrix; entries outside [0, bs) x [0, bs) are skipped (zero_outside = false)
// or written as zero (true).  tr: write the transposed tile.
__device__ __forceinline__ void tile_to_global(double* __restrict__ base, long long ld, int row0, int col0, int bs,
                                               const double (&t)[4][4], bool tr, bool zero_outside) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = row0 + r;
    if (!zero_outside && row >= bs) continue;
    double* dst = base + (long long)row * ld + col0;
#pragma unroll
    for (int q = 0; q < 4; q += 2) {
      double v0 = tr ? t[q][r] : t[r][q], v1 = tr ? t[q + 1][r] : t[r][q + 1];
      const bool in0 = row < bs && col0 + q < bs, in1 = row < bs && col0 + q + 1 < bs;
      if (zero_outside) {
        *reinterpret_cast<double2*>(dst + q) = make_double2(in0 ? v0 : 0.0, in1 ? v1 : 0.0);
      } else if (in1) {
        *reinterpret_cast<double2*>(dst + q) = make_double2(v0, v1);
      } else if (in0) {
        dst[q] = v0;
      }
    }
  }
}

__device__ __forceinline__ void half_tile_load(const double* __restrict__ p, double (&t)[2][4]) {
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const double2 u0 = *reinterpret_cast<const double2*>(p + 4 * r);
    const double2 u1 = *reinterpret_cast<const double2*>(p + 4 * r + 2);
    t[r][0] = u0.x; t[r][1] = u0.y; t[r][2] = u1.x; t[r][3] = u1.y;
  }
}

// shared memory of one sweep (46 848 bytes)
struct DiagSmem {
  double Lt[DIAG_LOW * DIAG_TS];          // final tiles of the factor
  double Xt[DIAG_LOW * DIAG_TS];          // final tiles of its inverse
  double Cn[2][DIAG_NT * DIAG_TS];        // tile column published for the next step
  double Dg[DIAG_NT][16];                 // pivot-tile numbers of each step
};

constexpr int DIAG_SRC_LD = NB + 4;      // leading dimension of a shared-memory source block (band mode)

// one barrier of the sweep: the whole CTA, or (BAND) named barrier 3 over the DIAG_THREADS sweep threads
template <bool BAND>
__device__ __forceinline__ void diag_sync() {
  if constexpr (BAND) asm volatile("bar.sync 3, %0;\n" ::"n"(DIAG_THREADS) : "memory");
  else __syncthreads();
}

// CG: read the block with ld.global.cg (persistent kernels: other SMs wrote it during this launch).
// Threads with tid >= DIAG_THREADS (if the CTA is larger) only take part in the barriers.
// BAND (persistent tile-DAG kernel, chain CTA): only the DIAG_THREADS sweep threads call this; the block is read from
// the shared-memory buffer `src` (row-major, leading dimension DIAG_SRC_LD) and, once every thread has its part in
// registers, the sweep threads arrive on named barrier 4 (CTA-wide count) so that the other warps may reuse `src`.
// SMEM_SRC without BAND (look-ahead diagonal kernel): the whole CTA calls this, the block is read from `src`, plain
// CTA barriers.
template <bool CG, bool BAND = false, bool SMEM_SRC = BAND>
__device__ __forceinline__ void diag_factor_block(DiagSmem& sm, MatView L, int b, int blk, int n,
                                                  double* __restrict__ dinv, int nblk, MatView linv, int has_linv,
                                                  int* __restrict__ status, long long* __restrict__ dbg,
                                                  const double* __restrict__ src = nullptr) {
  double (&Lt)[DIAG_LOW * DIAG_TS] = sm.Lt;
  double (&Xt)[DIAG_LOW * DIAG_TS] = sm.Xt;
  double (&Cn)[2][DIAG_NT * DIAG_TS] = sm.Cn;
  double (&Dg)[DIAG_NT][16] = sm.Dg;
  long long t0 = 0;
  if (dbg && threadIdx.x == 0 && blockIdx.x == 0) t0 = clock64();
#define DIAG_MARK(i) if (dbg && threadIdx.x == 0 && blockIdx.x == 0) dbg[i] = clock64() - t0;
  const int tid = threadIdx.x;
  const int r0 = blk * NB, bs = min(NB, n - r0);
  double* Lp = L.p + (long long)b * L.stride + (long long)r0 * L.ld + r0;
  // roles by warp.  Warps are dealt to the four schedulers round-robin, so warp 0 (the critical path) keeps its
  // scheduler's FP64 pipe to itself when warps 4 and 8 stay idle: update = warps 1,2,3,5,6; inverse = 7,9,10,11.
  const int warp = tid >> 5, lane = tid & 31;
  const bool owner = warp == 0;          // lane = tile row + 16 * (row pair of the tile)
  const int uidx = (warp < 4 ? warp - 1 : warp - 2) * 32 + lane;
  const int iidx = (warp == 7 ? 0 : warp - 8) * 32 + lane;
  const bool update = (warp == 1 || warp == 2 || warp == 3 || warp == 5 || warp == 6) && uidx < DIAG_LOW;
  const bool inverse = (warp == 7 || (warp >= 9 && warp < 12)) && iidx < DIAG_OFF;
  int ti = 0, tk = 0;
  if (update) {
    int rem = uidx;
    while (rem >= DIAG_NT - tk) { rem -= DIAG_NT - tk; ++tk; }
    ti = tk + rem;
  } else if (inverse) {
    int rem = iidx;
    while (rem >= DIAG_NT - 1 - tk) { rem -= DIAG_NT - 1 - tk; ++tk; }
    ti = tk + 1 + rem;
  } else if (owner) {
    ti = tid & 15;
  }
  const int half = (tid >> 4) & 1;
  double a[4][4];                        // update threads: the matrix tile; inverse threads: the running sum
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int q = 0; q < 4; ++q) a[r][q] = 0.0;
  if (update) {                          // identity-padded beyond bs; the strict upper part of diagonal tiles is unused
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int row = 4 * ti + r, col = 4 * tk + q;
        double v = (row == col) ? 1.0 : 0.0;
        if (row < bs && col <= row) {
          if constexpr (SMEM_SRC) v = src[row * DIAG_SRC_LD + col];
          else v = CG ? __ldcg(Lp + (long long)row * L.ld + col) : Lp[(long long)row * L.ld + col];
        }
        a[r][q] = v;
      }
    if (tk == 0) tile_store(Cn[0] + ti * DIAG_TS, a);
  }
  if constexpr (BAND) asm volatile("bar.arrive 4, 512;\n" ::: "memory");
  diag_sync<BAND>();
  DIAG_MARK(0)
  int base_prev = 0, base_cur = 0;       // diag_tile(i, s-1) = base_prev + i, diag_tile(i, s) = base_cur + i
#pragma unroll 1
  for (int s = 0; s <= DIAG_NT + 1; base_prev = base_cur, base_cur += DIAG_NT - 1 - s, ++s) {
    if (owner) {
      // ---- owners of tile column s: two lanes per tile (two rows each), uniform instruction stream: this is the
      //      critical path ----
      const bool live = ti >= s && s < DIAG_NT;
      double d[4][4], xk[4][4], c[2][4], xi[2][4];
      if (live) {
        tile_load(Cn[s & 1] + s * DIAG_TS, d);
        half_tile_load(Cn[s & 1] + ti * DIAG_TS + 8 * half, c);
        if (s > 0) {
          tile_load(Lt + (base_prev + s) * DIAG_TS, xk);
          half_tile_load(Lt + (base_prev + ti) * DIAG_TS + 8 * half, xi);
        }
      }
      if (live) {
        if (s > 0) {
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = 0; q <= r; ++q) {
              double acc = d[r][q];
#pragma unroll
              for (int m = 0; m < 4; ++m) acc = fma(-xk[r][m], xk[q][m], acc);
              d[r][q] = acc;
            }
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              double acc = c[r][q];
#pragma unroll
              for (int m = 0; m < 4; ++m) acc = fma(-xi[r][m], xk[q][m], acc);
              c[r][q] = acc;
            }
        }
        const double q0 = d[0][0], rs0 = fast_rsqrt(q0);
        const double l10 = d[1][0] * rs0, l20 = d[2][0] * rs0, l30 = d[3][0] * rs0;
        const double q1 = fma(-l10, l10, d[1][1]), rs1 = fast_rsqrt(q1);
        const double l21 = fma(-l20, l10, d[2][1]) * rs1, l31 = fma(-l30, l10, d[3][1]) * rs1;
        const double q2 = fma(-l21, l21, fma(-l20, l20, d[2][2])), rs2 = fast_rsqrt(q2);
        const double l32 = fma(-l31, l21, fma(-l30, l20, d[3][2])) * rs2;
        const double q3 = fma(-l32, l32, fma(-l31, l31, fma(-l30, l30, d[3][3]))), rs3 = fast_rsqrt(q3);
#pragma unroll
        for (int r = 0; r < 2; ++r) {          // X = C Ld^-T (on the diagonal lanes this is idle work)
          const double x0 = c[r][0] * rs0;
          const double x1 = fma(-x0, l10, c[r][1]) * rs1;
          const double x2 = fma(-x1, l21, fma(-x0, l20, c[r][2])) * rs2;
          const double x3 = fma(-x2, l32, fma(-x1, l31, fma(-x0, l30, c[r][3]))) * rs3;
          c[r][0] = x0; c[r][1] = x1; c[r][2] = x2; c[r][3] = x3;
        }
        if (ti == s) {                         // the pivot-tile numbers, turned into tiles by the helper next step
          if (half == 0) {
            c[0][0] = l10; c[0][1] = l20; c[0][2] = l30; c[0][3] = l21;
            c[1][0] = l31; c[1][1] = l32; c[1][2] = q0 * rs0; c[1][3] = q1 * rs1;
          } else {
            c[0][0] = q2 * rs2; c[0][1] = q3 * rs3; c[0][2] = rs0; c[0][3] = rs1;
            c[1][0] = rs2; c[1][1] = rs3;
          }
        }
        double* dst = (ti == s ? &Dg[s][0] : Lt + (base_cur + ti) * DIAG_TS) + 8 * half;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          *reinterpret_cast<double2*>(dst + 4 * r) = make_double2(c[r][0], c[r][1]);
          *reinterpret_cast<double2*>(dst + 4 * r + 2) = make_double2(c[r][2], c[r][3]);
        }
      }
    } else {
      if (update) {
        if (tk > s) {
          // ---- the trailing block: update of column s-1; publish column s+1 for the next step ----
          if (s > 0) {
            double xi[4][4], xk[4][4];
            tile_load(Lt + (base_prev + ti) * DIAG_TS, xi);
            tile_load(Lt + (base_prev + tk) * DIAG_TS, xk);
            tile_sub_abt(a, xi, xk);
          }
          if (tk == s + 1) tile_store(Cn[(s + 1) & 1] + ti * DIAG_TS, a);
        }
      } else if (inverse) {
        // ---- X tile (ti, tk): add the term whose operands became final; finish two steps after row ti of L ----
        const int k = (s == tk + 2) ? tk : s - 3;
        if ((s == tk + 2) || (k > tk && k < ti)) {
          double l[4][4], x[4][4];
          tile_load(Lt + diag_tile(ti, k) * DIAG_TS, l);
          tile_load(Xt + diag_tile(k, tk) * DIAG_TS, x);
          tile_add_ab(a, l, x);
        }
        if (s == ti + 2) {
          double xd[4][4], x[4][4];
          tile_load(Xt + diag_tile(ti, ti) * DIAG_TS, xd);
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              double acc = 0.0;
#pragma unroll
              for (int m = 0; m <= r; ++m) acc = fma(-xd[r][m], a[m][q], acc);
              x[r][q] = acc;
            }
          tile_store(Xt + diag_tile(ti, tk) * DIAG_TS, x);
        }
      } else if (tid == DIAG_HELPER && s > 0 && s <= DIAG_NT) {
        // ---- diagonal tiles of L and of X for column s-1, pivot check ----
        const int j = s - 1;
        const double* g = &Dg[j][0];
        const double l10 = g[0], l20 = g[1], l30 = g[2], l21 = g[3], l31 = g[4], l32 = g[5];
        const double d0 = g[6], d1 = g[7], d2 = g[8], d3 = g[9];
        const double i00 = g[10], i11 = g[11], i22 = g[12], i33 = g[13];
        int first_bad = 0;                     // d = q rsqrt(q) is positive exactly when the pivot q is
        if (!(d3 > 0.0)) first_bad = 4;
        if (!(d2 > 0.0)) first_bad = 3;
        if (!(d1 > 0.0)) first_bad = 2;
        if (!(d0 > 0.0)) first_bad = 1;
        if (first_bad) atomicCAS(&status[b], 0, r0 + 4 * j + first_bad);
        double t[4][4];
        t[0][0] = d0;  t[0][1] = 0.0; t[0][2] = 0.0; t[0][3] = 0.0;
        t[1][0] = l10; t[1][1] = d1;  t[1][2] = 0.0; t[1][3] = 0.0;
        t[2][0] = l20; t[2][1] = l21; t[2][2] = d2;  t[2][3] = 0.0;
        t[3][0] = l30; t[3][1] = l31; t[3][2] = l32; t[3][3] = d3;
        tile_store(Lt + diag_tile(j, j) * DIAG_TS, t);
        const double i10 = -l10 * i00 * i11;
        const double i21 = -l21 * i11 * i22;
        const double i32 = -l32 * i22 * i33;
        const double i20 = -(l20 * i00 + l21 * i10) * i22;
        const double i31 = -(l31 * i11 + l32 * i21) * i33;
        const double i30 = -(l30 * i00 + l31 * i10 + l32 * i20) * i33;
        t[0][0] = i00;
        t[1][0] = i10; t[1][1] = i11;
        t[2][0] = i20; t[2][1] = i21; t[2][2] = i22;
        t[3][0] = i30; t[3][1] = i31; t[3][2] = i32; t[3][3] = i33;
        tile_store(Xt + diag_tile(j, j) * DIAG_TS, t);
      }
    }
    diag_sync<BAND>();
  }
  DIAG_MARK(1)
  // ---- results to global memory, row by row: a warp writes one 512-byte row segment per pass (the per-tile version of
  //      round 1 took 1.3 us of the 12.8 us kernel: 20 scattered 16-byte stores per thread) ----
  double* dp = dinv + ((long long)b * nblk + blk) * (NB * NB);
  double* ip = has_linv ? linv.p + (long long)b * linv.stride + (long long)r0 * linv.ld + r0 : nullptr;
  auto low = [](const double* T, int row, int col) {       // entry (row, col), row >= col, of the packed tiles
    return T + diag_tile(row >> 2, col >> 2) * DIAG_TS + 4 * (row & 3) + (col & 3);
  };
  if (tid < DIAG_THREADS) {
    for (int e = tid; e < NB * NB / 2; e += DIAG_THREADS) {
      const int row = e >> 5, col = (e & 31) * 2;          // entries (row, col) and (row, col + 1): one tile row
      // the inverted block for the panel products: all of 64 x 64, zero above the diagonal and outside bs x bs
      double2 x = make_double2(0.0, 0.0);
      if (row < bs && col <= row) {
        const double2 v = *reinterpret_cast<const double2*>(low(Xt, row, col));
        x.x = v.x;
        x.y = (col + 1 <= row) ? v.y : 0.0;
      }
      *reinterpret_cast<double2*>(dp + row * NB + col) = x;
      if (row >= bs) continue;
      if (col <= row) {                                    // the factor: the strict upper triangle of the matrix buffer is
        const double2 v = *reinterpret_cast<const double2*>(low(Lt, row, col));   // scratch of the triangular inverse
        double* dst = Lp + (long long)row * L.ld + col;
        if (col + 1 <= row) *reinterpret_cast<double2*>(dst) = v;
        else dst[0] = v.x;
      }
      if (ip && col < bs) {                                // level-0 state of the triangular inverse: X and, above the
        double2 w;                                         // diagonal, its transpose
        w.x = (col <= row) ? x.x : *low(Xt, col, row);
        w.y = (col + 1 <= row) ? x.y : (col + 1 < bs ? *low(Xt, col + 1, row) : 0.0);
        double* dst = ip + (long long)row * linv.ld + col;
        if (col + 1 < bs) *reinterpret_cast<double2*>(dst) = w;
        else dst[0] = w.x;
      }
    }
  }
  DIAG_MARK(4)
#undef DIAG_MARK
}

}  // namespace spw
