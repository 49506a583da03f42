// spatial_gradient hot loop: for one posterior sample and a tile of grid points, stream
//   W = Linv * A^T   (Linv = chol(K)^-1 lower triangular, A = nabla.K.t rows of the grid tile)
// through the FP64 tensor pipe row block by row block, and fold each finished row block into the
// per-grid-point Gram matrix M = A K^-1 A^T = W^T W and mean = W^T (Linv z) -- W never leaves the SM.
// The last row block is followed by the reference's epilogue: var = V.0 - M * sign, upper-triangle
// Cholesky, draw = sqrt(sig2) * U eps + mean.
//   tmp / mean.grad / var.grad / grad.est   R/spatial_gradient.R:238-242 (gaussian), :274-278 (matern1),
//                                           :329-333 (matern2)
//   V.0                                     R/spatial_gradient.R:226-227, :271, :317-318
#include "gradient.cuh"
#include "gemm_nt.cuh"

namespace spw {

template <int NCOMP>
struct GradCfg;
template <>
struct GradCfg<5> {
  static constexpr int GPT = 16;                 // grid points per CTA tile
  using Cfg = NtCfg<128, 80, 16, 3, 4, 2>;       // warp tile 32 x 40
};
template <>
struct GradCfg<2> {
  static constexpr int GPT = 32;
  using Cfg = NtCfg<128, 64, 16, 3, 4, 2>;       // warp tile 32 x 32
};

template <int NCOMP>
struct GradSmem {
  using Cfg = typename GradCfg<NCOMP>::Cfg;
  static constexpr int GPT = GradCfg<NCOMP>::GPT;
  static constexpr int NPAIR = NCOMP * (NCOMP + 1) / 2;
  static constexpr int IPG = NPAIR + NCOMP;      // accumulators per grid point: Gram upper triangle + mean
  static constexpr int NITEMS = GPT * IPG;
  static constexpr int LDW = Cfg::BN + 8;        // == 8 mod 16: conflict-free 16-byte fragment stores
  static constexpr size_t PIPE = Cfg::PIPE_BYTES;
  static constexpr size_t WS = size_t(Cfg::BM) * LDW * sizeof(double);
  static constexpr size_t VS = size_t(Cfg::BM) * sizeof(double);
  static constexpr size_t MS = size_t(NITEMS) * sizeof(double);
  static constexpr size_t TOTAL = PIPE + WS + VS + MS;
  static_assert(NITEMS <= 2 * Cfg::NTHREADS, "two accumulators per thread");
  static_assert(LDW % 16 == 8, "staging leading dimension");
};

__device__ __forceinline__ void pair_of(int ncomp, int p, int& a, int& b) {
  // upper triangle, row-wise: (0,0),(0,1),...,(0,c-1),(1,1),...
  a = 0;
  int rem = p, len = ncomp;
  while (rem >= len) {
    rem -= len;
    --len;
    ++a;
  }
  b = a + rem;
}

template <int TYPE>
__global__ void __launch_bounds__(256, 1)
gradient_main_kernel(GradParams P) {
  constexpr int NCOMP = (TYPE == 1) ? 2 : 5;
  using SM = GradSmem<NCOMP>;
  using Cfg = typename SM::Cfg;
  constexpr int GPT = SM::GPT;
  extern __shared__ __align__(16) double smem[];
  double* pipe = smem;
  double* Ws = smem + SM::PIPE / sizeof(double);
  double* vs = Ws + Cfg::BM * SM::LDW;
  double* Ms = vs + Cfg::BM;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_m = warp % Cfg::WARPS_M, warp_n = warp / Cfg::WARPS_M;
  const int g = lane >> 2, q = lane & 3;
  const int b = blockIdx.y;
  const int g0 = blockIdx.x * GPT;                         // first grid point (chunk-local) of this tile
  const int n = P.n;
  const int n_valid = min(Cfg::BN, (P.g_count - g0) * NCOMP);
  const double* Linv = P.linv.p + (long long)b * P.linv.stride;
  const double* At = P.at.p + (long long)b * P.at.stride + (long long)g0 * NCOMP * P.at.ld;
  const double* vvec = P.v + (long long)b * P.ldv;

  const int nrb = (n + Cfg::BM - 1) / Cfg::BM;
  // flattened (row block, k tile) cursors
  int l_rb = 0, l_kt = 0, l_nkt = (min(Cfg::BM, n) + Cfg::BK - 1) / Cfg::BK;
  int c_rb = 0, c_kt = 0, c_nkt = l_nkt;
  int flat = 0;

  auto issue_load = [&](int stage) {
    if (l_rb < nrb) {
      const int r0 = l_rb * Cfg::BM, k0 = l_kt * Cfg::BK;
      const int k_end = min(r0 + Cfg::BM, n);
      double* st = pipe + stage * Cfg::STAGE;
      load_tile_async<Cfg::BM, Cfg::BK, Cfg::LDS, Cfg::NTHREADS>(st, Linv + (long long)r0 * P.linv.ld + k0, P.linv.ld,
                                                                 min(Cfg::BM, n - r0), k_end - k0, tid);
      load_tile_async<Cfg::BN, Cfg::BK, Cfg::LDS, Cfg::NTHREADS>(st + Cfg::A_STAGE, At + k0, P.at.ld, n_valid,
                                                                 k_end - k0, tid);
      if (++l_kt == l_nkt) {
        ++l_rb;
        l_kt = 0;
        l_nkt = (min((l_rb + 1) * Cfg::BM, n) + Cfg::BK - 1) / Cfg::BK;
      }
    }
    cp_async_commit();
  };

  double acc[Cfg::MF][Cfg::NF][2];
#pragma unroll
  for (int i = 0; i < Cfg::MF; ++i)
#pragma unroll
    for (int j = 0; j < Cfg::NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double gram[2] = {0.0, 0.0};

#pragma unroll
  for (int s = 0; s < Cfg::STAGES - 1; ++s) issue_load(s);

  while (c_rb < nrb) {
    cp_async_wait<Cfg::STAGES - 2>();
    __syncthreads();
    issue_load((flat + Cfg::STAGES - 1) % Cfg::STAGES);
    const double* st = pipe + (flat % Cfg::STAGES) * Cfg::STAGE;
    const int k0 = c_kt * Cfg::BK;
    TriMask tm{1, c_rb * Cfg::BM, 0, 0};
    if (ktile_needs_mask<Cfg>(k0, tm))
      compute_ktile<Cfg, true>(st, st + Cfg::A_STAGE, acc, warp_m, warp_n, lane, k0, tm);
    else
      compute_ktile<Cfg, false>(st, st + Cfg::A_STAGE, acc, warp_m, warp_n, lane, k0, tm);
    ++flat;
    if (++c_kt == c_nkt) {
      // ---- row block finished: stage W, fold into Gram / mean accumulators ----
      const int r0 = c_rb * Cfg::BM;
#pragma unroll
      for (int i = 0; i < Cfg::MF; ++i) {
        const int r = warp_m * Cfg::WM + i * 8 + g;
#pragma unroll
        for (int j = 0; j < Cfg::NF; ++j) {
          const int c = warp_n * Cfg::WN + j * 8 + 2 * q;
          *reinterpret_cast<double2*>(Ws + r * SM::LDW + c) = make_double2(acc[i][j][0], acc[i][j][1]);
          acc[i][j][0] = acc[i][j][1] = 0.0;
        }
      }
      if (tid < Cfg::BM) vs[tid] = (r0 + tid < n) ? vvec[r0 + tid] : 0.0;
      __syncthreads();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int it = tid + h * Cfg::NTHREADS;
        if (it < SM::NITEMS) {
          const int gl = it / SM::IPG, p = it % SM::IPG;
          double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
          if (p < SM::NPAIR) {
            int a, bb;
            pair_of(NCOMP, p, a, bb);
            const double* wa = Ws + gl * NCOMP + a;
            const double* wb = Ws + gl * NCOMP + bb;
#pragma unroll 4
            for (int r = 0; r < Cfg::BM; r += 4) {
              s0 = fma(wa[(r + 0) * SM::LDW], wb[(r + 0) * SM::LDW], s0);
              s1 = fma(wa[(r + 1) * SM::LDW], wb[(r + 1) * SM::LDW], s1);
              s2 = fma(wa[(r + 2) * SM::LDW], wb[(r + 2) * SM::LDW], s2);
              s3 = fma(wa[(r + 3) * SM::LDW], wb[(r + 3) * SM::LDW], s3);
            }
          } else {
            const double* wa = Ws + gl * NCOMP + (p - SM::NPAIR);
#pragma unroll 4
            for (int r = 0; r < Cfg::BM; r += 4) {
              s0 = fma(wa[(r + 0) * SM::LDW], vs[r + 0], s0);
              s1 = fma(wa[(r + 1) * SM::LDW], vs[r + 1], s1);
              s2 = fma(wa[(r + 2) * SM::LDW], vs[r + 2], s2);
              s3 = fma(wa[(r + 3) * SM::LDW], vs[r + 3], s3);
            }
          }
          gram[h] += (s0 + s1) + (s2 + s3);
        }
      }
      ++c_rb;
      c_kt = 0;
      c_nkt = (min((c_rb + 1) * Cfg::BM, n) + Cfg::BK - 1) / Cfg::BK;
    }
  }
  cp_async_wait<0>();

  // ---- publish accumulators, then one thread per grid point finishes the reference's epilogue ----
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int it = tid + h * Cfg::NTHREADS;
    if (it < SM::NITEMS) Ms[it] = gram[h];
  }
  __syncthreads();
  if (tid < GPT && g0 + tid < P.g_count) {
    const int gl = tid;
    const long long gidx = P.g_begin + g0 + gl;            // grid point index in the full grid
    const long long sidx = P.s_begin + b;                  // sample index in the shard
    const double* m = Ms + gl * SM::IPG;
    double M[NCOMP][NCOMP], mean[NCOMP];
    {
      int p = 0;
#pragma unroll
      for (int a = 0; a < NCOMP; ++a)
#pragma unroll
        for (int c = a; c < NCOMP; ++c) {
          M[a][c] = m[p];
          M[c][a] = m[p];
          ++p;
        }
#pragma unroll
      for (int a = 0; a < NCOMP; ++a) mean[a] = m[SM::NPAIR + a];
    }
    const double ph = P.phi[b];
    double V0[NCOMP][NCOMP];
#pragma unroll
    for (int a = 0; a < NCOMP; ++a)
#pragma unroll
      for (int c = 0; c < NCOMP; ++c) V0[a][c] = 0.0;
    if constexpr (TYPE == 1) {
      V0[0][0] = V0[1][1] = 3.0 * ph * ph;                                   // R/spatial_gradient.R:271
    } else {
      const double p2 = ph * ph, p4 = p2 * p2;
      const double vg = (TYPE == 0) ? 2.0 * p2 : 5.0 / 3.0 * p2;             // :226 / :317
      const double vh = (TYPE == 0) ? 4.0 * p4 : 25.0 / 3.0 * p4;            // :227 / :318  (times I.tilde)
      V0[0][0] = V0[1][1] = vg;
      V0[2][2] = V0[4][4] = 3.0 * vh;
      V0[3][3] = vh;
      V0[2][4] = V0[4][2] = vh;
    }
    // var.grad = V.0 - crossprod(tmp, nabla.K): column b of nabla.K carries sign s_b = -1 for the gradient
    double var[NCOMP][NCOMP];
#pragma unroll
    for (int a = 0; a < NCOMP; ++a)
#pragma unroll
      for (int c = 0; c < NCOMP; ++c) var[a][c] = V0[a][c] - M[a][c] * ((c < 2) ? -1.0 : 1.0);
    if (P.out_mean) {
      double* o = P.out_mean + (sidx * P.g_total + gidx) * NCOMP;
#pragma unroll
      for (int a = 0; a < NCOMP; ++a) o[a] = mean[a];
    }
    if (P.out_var) {
      double* o = P.out_var + (sidx * P.g_total + gidx) * (NCOMP * NCOMP);
#pragma unroll
      for (int a = 0; a < NCOMP; ++a)
#pragma unroll
        for (int c = 0; c < NCOMP; ++c) o[a + NCOMP * c] = var[a][c];   // column-major like R
    }
    // chol(var.grad): upper factor from the upper triangle only
    double U[NCOMP][NCOMP];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < NCOMP; ++j) {
#pragma unroll
      for (int i = 0; i <= j; ++i) {
        double s = var[i][j];
#pragma unroll
        for (int k = 0; k < i; ++k) s -= U[k][i] * U[k][j];
        if (i == j) {
          if (!(s > 0.0)) ok = false;
          U[j][j] = sqrt(s);
        } else {
          U[i][j] = s / U[i][i];
        }
      }
    }
    if (!ok) atomicCAS(&P.status[b], 0, (int)(gidx + 1));
    if (P.out_draws) {
      const double sd = sqrt(P.sig2[b]);
      const double* e = P.eps + (sidx * P.g_total + gidx) * NCOMP;
#pragma unroll
      for (int a = 0; a < NCOMP; ++a) {
        double s = 0.0;
#pragma unroll
        for (int c = a; c < NCOMP; ++c) s += U[a][c] * e[c];
        P.out_draws[(sidx * NCOMP + a) * P.g_total + gidx] = sd * s + mean[a];
      }
    }
  }
}

int gradient_main_launch(int type, const GradParams& P, int batch, cudaStream_t st) {
  if (batch <= 0 || P.g_count <= 0) return SPW_OK;
  static bool attr_set[64][3] = {{false}};
  int dev = 0;
  SPW_CUDA(cudaGetDevice(&dev));
  const int ncomp = (type == 1) ? 2 : 5;
  const int gpt = (ncomp == 5) ? GradCfg<5>::GPT : GradCfg<2>::GPT;
  const size_t smem = (ncomp == 5) ? GradSmem<5>::TOTAL : GradSmem<2>::TOTAL;
  if (!attr_set[dev & 63][type]) {
    switch (type) {
      case 0: SPW_CUDA(cudaFuncSetAttribute(gradient_main_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); break;
      case 1: SPW_CUDA(cudaFuncSetAttribute(gradient_main_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); break;
      case 2: SPW_CUDA(cudaFuncSetAttribute(gradient_main_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); break;
    }
    attr_set[dev & 63][type] = true;
  }
  dim3 grid(ceil_div(P.g_count, gpt), batch);
  switch (type) {
    case 0: gradient_main_kernel<0><<<grid, 256, smem, st>>>(P); break;
    case 1: gradient_main_kernel<1><<<grid, 256, smem, st>>>(P); break;
    case 2: gradient_main_kernel<2><<<grid, 256, smem, st>>>(P); break;
    default: set_error("bad cov type %d", type); return SPW_ERR_ARG;
  }
  SPW_LAUNCHED();
  return SPW_OK;
}

}  // namespace spw
