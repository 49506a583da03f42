// spatial_gradient hot loop, TMA / mbarrier / warp-specialised version (the default path).
//
// Same mathematics as gradient.cu (W = Linv * A^T streamed through the FP64 tensor pipe, folded on the fly into
// M = W^T W and mean = W^T (Linv z), then var = V.0 - M * sign, upper-triangle 5x5 Cholesky and the draw --
// R/spatial_gradient.R:238-242, :274-278, :329-333), different data movement:
//   * one producer thread issues two cp.async.bulk.tensor (TMA, SASS UTMALDG) loads per k tile -- the 128 x BK
//     Linv tile and the BN x BK cross-covariance tile -- into a 3-stage shared-memory ring; completion is tracked
//     by mbarrier transaction counts, slot reuse by a second mbarrier the consumer warps arrive on; there is no
//     CTA-wide barrier in the main loop;
//   * the TMA box is BK + 4 doubles wide, so rows land in shared memory with the 4-double pad that makes the
//     DMMA fragment loads bank-conflict free (the 4 extra columns are never read); out-of-range rows / columns
//     are zero-filled by the TMA unit, so ragged N and ragged grid tiles need no special code;
//   * 8 consumer warps (2 per SM sub-partition) own 32 x (BN/2) accumulator tiles; each warp folds its finished
//     row block into its own Gram / mean accumulators through a private staging buffer (warp-local, no CTA sync);
//     the four row-groups are summed in a fixed order at the very end (deterministic).
#include "gradient.cuh"
#include "tma.cuh"

namespace spw {

namespace {

constexpr int T_BM = 128, T_BK = 32, T_STAGES = 3, T_LDS = T_BK + 4;
constexpr int T_CONSUMER_WARPS = 8, T_THREADS = 32 * (T_CONSUMER_WARPS + 1);

template <int NCOMP>
struct TCfg {
  static constexpr int GPT = (NCOMP == 5) ? 16 : 32;      // grid points per CTA tile
  static constexpr int BN = GPT * NCOMP;                  // 80 / 64
  static constexpr int WN = BN / 2;                       // 40 / 32 : two warp columns
  static constexpr int MF = 4, NF = WN / 8;               // warp tile 32 x WN
  static constexpr int GPW = GPT / 2;                     // grid points per warp column
  static constexpr int NPAIR = NCOMP * (NCOMP + 1) / 2;
  static constexpr int IPG = NPAIR + NCOMP;
  static constexpr int NIT_W = GPW * IPG;                 // accumulators per warp: 160 / 80
  static constexpr int NACC = (NIT_W + 31) / 32;          // per lane: 5 / 3
  static constexpr int LDG = 40;                          // staging leading dimension (== 8 mod 16)
  static constexpr int A_STAGE = T_BM * T_LDS, B_STAGE = BN * T_LDS;          // doubles
  static constexpr int STAGE = A_STAGE + B_STAGE;
  static constexpr uint32_t STAGE_BYTES = STAGE * sizeof(double);
  static constexpr size_t PIPE_BYTES = size_t(T_STAGES) * STAGE_BYTES;
  static constexpr size_t STG_BYTES = size_t(T_CONSUMER_WARPS) * 16 * LDG * sizeof(double);   // 16-row halves
  static constexpr size_t BAR_BYTES = 2 * T_STAGES * sizeof(unsigned long long);
  static constexpr size_t TOTAL = PIPE_BYTES + STG_BYTES + BAR_BYTES + 128;
  static_assert(size_t(T_CONSUMER_WARPS) * NIT_W * sizeof(double) <= PIPE_BYTES, "final reduction aliases the ring");
};

struct TmaParams {
  const double* v;         // [batch][ldv]
  long long ldv;
  const double* phi;
  const double* sig2;
  const double* eps;
  double* out_draws;
  double* out_mean;
  double* out_var;
  double* out_gram;
  int* status;
  int n, g_begin, g_count, g_total, s_begin;
};

template <int TYPE>
__global__ void __launch_bounds__(T_THREADS, 1)
gradient_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TmaParams P) {
  constexpr int NCOMP = (TYPE == 1) ? 2 : 5;
  using C = TCfg<NCOMP>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* pipe = reinterpret_cast<double*>(smem_raw);
  double* stg_all = reinterpret_cast<double*>(smem_raw + C::PIPE_BYTES);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw + C::PIPE_BYTES + C::STG_BYTES);
  unsigned long long* empty = full + T_STAGES;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int g0 = blockIdx.x * C::GPT;
  const int n = P.n;
  // 128-row blocks aligned to the END of the matrix: block rb covers rows rb*128 - pad ... (+128); the ragged block
  // is the first one (rows < 0 are outside the tensor: TMA zero-fills them), where the k range is shortest
  const int pad = (T_BM - n % T_BM) % T_BM;
  const int nrb = (n + pad) / T_BM;

  if (tid == 0) {
    for (int s = 0; s < T_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], T_CONSUMER_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == T_CONSUMER_WARPS) {
    // ===================== producer: one thread feeds the ring with TMA tile loads =====================
    if (lane == 0) {
      int f = 0;
      for (int rb = 0; rb < nrb; ++rb) {
        const int r0 = rb * T_BM - pad;
        const int nkt = (r0 + T_BM + T_BK - 1) / T_BK;
        for (int kt = 0; kt < nkt; ++kt, ++f) {
          const int s = f % T_STAGES;
          const uint32_t ph = (f / T_STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          double* st = pipe + (size_t)s * C::STAGE;
          mbar_expect_tx(&full[s], C::STAGE_BYTES);
          tma_load_3d(st, &tmA, kt * T_BK, r0, b, &full[s]);
          tma_load_3d(st + C::A_STAGE, &tmB, kt * T_BK, g0 * NCOMP, b, &full[s]);
        }
      }
    }
  } else {
    // ===================== consumers: DMMA main loop + warp-local Gram epilogue =========================
    // the two warps of an SM sub-partition (warp & 3) own row strips wm and 3 - wm, so the strips' unequal shares of
    // the diagonal tiles (a strip skips the k tiles to the right of its last row) add up to the same work everywhere
    const int warp_n = warp >> 2;
    const int warp_m = warp_n ? (7 - warp) : warp;
    const int g = lane >> 2, q = lane & 3;
    double* stg = stg_all + (size_t)warp * 16 * C::LDG;
    const double* vvec = P.v + (long long)b * P.ldv;

    // per-lane Gram items: it = lane + 32 t  ->  (grid point in warp column, pair / mean component)
    int col_a[C::NACC], col_b[C::NACC];
    bool is_mean[C::NACC], live[C::NACC];
#pragma unroll
    for (int t = 0; t < C::NACC; ++t) {
      const int it = lane + 32 * t;
      live[t] = it < C::NIT_W;
      const int gl = live[t] ? it / C::IPG : 0, p = live[t] ? it % C::IPG : 0;
      is_mean[t] = p >= C::NPAIR;
      int a = 0, bb = 0;
      if (!is_mean[t]) {
        int rem = p, len = NCOMP;
        while (rem >= len) { rem -= len; --len; ++a; }
        bb = a + rem;
      } else {
        a = bb = p - C::NPAIR;
      }
      col_a[t] = gl * NCOMP + a;
      col_b[t] = gl * NCOMP + bb;
    }
    double gram[C::NACC];
#pragma unroll
    for (int t = 0; t < C::NACC; ++t) gram[t] = 0.0;

    double acc[C::MF][C::NF][2];
#pragma unroll
    for (int i = 0; i < C::MF; ++i)
#pragma unroll
      for (int j = 0; j < C::NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    int f = 0;
    for (int rb = 0; rb < nrb; ++rb) {
      const int r0 = rb * T_BM - pad;
      const int nkt = (r0 + T_BM + T_BK - 1) / T_BK;
      const int strip0 = r0 + warp_m * 32;                 // first row of this warp's strip (may be < 0 in block 0)
      for (int kt = 0; kt < nkt; ++kt, ++f) {
        const int s = f % T_STAGES;
        const uint32_t ph = (f / T_STAGES) & 1;
        mbar_wait(&full[s], ph);
        const double* As = pipe + (size_t)s * C::STAGE;
        const double* Bs = As + C::A_STAGE;
        const double* a_base = As + (warp_m * 32 + g) * T_LDS + q;
        const double* b_base = Bs + (warp_n * C::WN + g) * T_LDS + q;
        const int k0 = kt * T_BK;
        const bool masked = (k0 + T_BK - 1 > strip0);      // tile touches the diagonal: Linv's upper triangle holds
                                                           // the transposed copy, which must read as zero here
        if (k0 <= strip0 + 31) {                           // else: the whole tile lies right of the strip's diagonal
#pragma unroll
        for (int kk = 0; kk < T_BK / 4; ++kk) {
          double a[C::MF], bf[C::NF];
#pragma unroll
          for (int i = 0; i < C::MF; ++i) a[i] = a_base[i * 8 * T_LDS + kk * 4];
#pragma unroll
          for (int j = 0; j < C::NF; ++j) bf[j] = b_base[j * 8 * T_LDS + kk * 4];
          if (masked) {
            const int k = k0 + kk * 4 + q;
#pragma unroll
            for (int i = 0; i < C::MF; ++i)
              if (k > r0 + warp_m * 32 + i * 8 + g) a[i] = 0.0;
          }
#pragma unroll
          for (int i = 0; i < C::MF; ++i)
#pragma unroll
            for (int j = 0; j < C::NF; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bf[j]);
        }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
      }
      // ---- row block done: fold this warp's 32 x WN tile of W into its Gram / mean accumulators ----
      const int rw = strip0 + lane;
      const double vreg = (rw >= 0 && rw < n) ? vvec[rw] : 0.0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        __syncwarp();
#pragma unroll
        for (int ii = 0; ii < 2; ++ii) {
          const int i = 2 * h + ii;
#pragma unroll
          for (int j = 0; j < C::NF; ++j) {
            *reinterpret_cast<double2*>(stg + (ii * 8 + g) * C::LDG + j * 8 + 2 * q) =
                make_double2(acc[i][j][0], acc[i][j][1]);
            acc[i][j][0] = acc[i][j][1] = 0.0;
          }
        }
        __syncwarp();
#pragma unroll 4
        for (int rr = 0; rr < 16; ++rr) {
          const double vr = __shfl_sync(0xffffffffu, vreg, 16 * h + rr);
          const double* row = stg + rr * C::LDG;
#pragma unroll
          for (int t = 0; t < C::NACC; ++t) {
            const double x = row[col_a[t]];
            const double y = is_mean[t] ? vr : row[col_b[t]];
            gram[t] = fma(x, y, gram[t]);
          }
        }
      }
    }
    // ---- cross-warp reduction over the four row groups (fixed order), then the reference's epilogue ----
    asm volatile("bar.sync 1, %0;\n" ::"n"(32 * T_CONSUMER_WARPS));   // every consumer is past its last tile
    double* Ms = pipe;                                               // [warp][NIT_W], aliases the (drained) ring
#pragma unroll
    for (int t = 0; t < C::NACC; ++t)
      if (live[t]) Ms[(warp_n * 4 + warp_m) * C::NIT_W + lane + 32 * t] = gram[t];
    asm volatile("bar.sync 1, %0;\n" ::"n"(32 * T_CONSUMER_WARPS));
    if (tid < C::GPT && g0 + tid < P.g_count) {
      const int gl = tid;
      const int wn = gl / C::GPW, glw = gl % C::GPW;
      const long long gidx = P.g_begin + g0 + gl;
      const long long sidx = P.s_begin + b;
      double m[C::IPG];
#pragma unroll
      for (int p = 0; p < C::IPG; ++p) {
        double s = 0.0;
#pragma unroll
        for (int wm = 0; wm < 4; ++wm) s += Ms[(wn * 4 + wm) * C::NIT_W + glw * C::IPG + p];
        m[p] = s;
      }
      if (P.out_gram) {
        double* o = P.out_gram + (sidx * P.g_total + gidx) * C::IPG;
#pragma unroll
        for (int p = 0; p < C::IPG; ++p) o[p] = m[p];
      }
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
        for (int a = 0; a < NCOMP; ++a) mean[a] = m[C::NPAIR + a];
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
}

}  // namespace

int gradient_tma_launch(int type, const GradParams& G, int batch, cudaStream_t st) {
  if (batch <= 0 || G.g_count <= 0) return SPW_OK;
  const int ncomp = (type == 1) ? 2 : 5;
  const int gpt = (ncomp == 5) ? TCfg<5>::GPT : TCfg<2>::GPT;
  const int bn = (ncomp == 5) ? TCfg<5>::BN : TCfg<2>::BN;
  const size_t smem = (ncomp == 5) ? TCfg<5>::TOTAL : TCfg<2>::TOTAL;
  static bool attr_set[64][3] = {{false}};
  int dev = 0;
  SPW_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev & 63][type]) {
    switch (type) {
      case 0: SPW_CUDA(cudaFuncSetAttribute(gradient_tma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); break;
      case 1: SPW_CUDA(cudaFuncSetAttribute(gradient_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); break;
      case 2: SPW_CUDA(cudaFuncSetAttribute(gradient_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); break;
      default: set_error("bad cov type %d", type); return SPW_ERR_ARG;
    }
    attr_set[dev & 63][type] = true;
  }
  CUtensorMap tmA, tmB;
  SPW_TRY(make_tensor_map(&tmA, G.linv, G.n, G.n, batch, T_LDS, T_BM));
  SPW_TRY(make_tensor_map(&tmB, G.at, G.n, G.g_count * ncomp, batch, T_LDS, bn));
  TmaParams P;
  P.v = G.v; P.ldv = G.ldv; P.phi = G.phi; P.sig2 = G.sig2; P.eps = G.eps;
  P.out_draws = G.out_draws; P.out_mean = G.out_mean; P.out_var = G.out_var; P.out_gram = G.out_gram; P.status = G.status;
  P.n = G.n; P.g_begin = G.g_begin; P.g_count = G.g_count; P.g_total = G.g_total; P.s_begin = G.s_begin;
  dim3 grid(ceil_div(G.g_count, gpt), batch);
  switch (type) {
    case 0: gradient_tma_kernel<0><<<grid, T_THREADS, smem, st>>>(tmA, tmB, P); break;
    case 1: gradient_tma_kernel<1><<<grid, T_THREADS, smem, st>>>(tmA, tmB, P); break;
    case 2: gradient_tma_kernel<2><<<grid, T_THREADS, smem, st>>>(tmA, tmB, P); break;
  }
  SPW_LAUNCHED();
  return SPW_OK;
}

}  // namespace spw
