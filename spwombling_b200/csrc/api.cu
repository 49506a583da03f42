// C ABI of libspwombling: argument checking, device workspaces, host<->device marshalling of R-layout
// arrays, and the per-shard driver of the spatial_gradient pipeline.  See include/spwombling.h.
#include <atomic>
#include <cstdarg>
#include <ctime>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "../../include/spwombling.h"
#include "common.cuh"
#include "cov.cuh"
#include "gradient.cuh"
#include "hlm.cuh"
#include "linalg.cuh"
#include "potrf_dag.cuh"
#include "potrf_la.cuh"
#include "summary.cuh"
#include "tma.cuh"
#include "wombling.cuh"

namespace spw {

static thread_local char tl_error[1024] = "";
static char g_error[1024] = "";
static std::mutex g_err_mu;
std::atomic<long long> g_launches{0};
static long long g_ws_limit = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
  va_end(ap);
  std::lock_guard<std::mutex> lk(g_err_mu);
  memcpy(g_error, tl_error, sizeof(g_error));
}

// ---- optional phase timing (CUDA events on the launching stream) ----------------------------------------
enum Phase { PH_COV = 0, PH_POTRF, PH_TRTRI, PH_TRMV, PH_CROSS, PH_MAIN, PH_COUNT };
static int g_profile = 0;
static double g_phase_ms[PH_COUNT] = {0};
static long long g_phase_calls[PH_COUNT] = {0};
static double g_phase_host_ms[PH_COUNT] = {0};   // host wall time spent issuing each phase
static double g_main_flops = 0.0;          // algorithmic flops of the timed gradient_main launches
static std::mutex g_prof_mu;

struct PhaseTimer {
  struct Rec { int phase; cudaEvent_t e0, e1; double h0, h1; };
  static double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  }
  std::vector<Rec> recs;
  cudaStream_t st;
  explicit PhaseTimer(cudaStream_t s) : st(s) {}
  cudaStream_t cur = nullptr;
  void begin(int phase, cudaStream_t on) {
    if (!g_profile) return;
    Rec r;
    r.phase = phase;
    cur = on;
    cudaEventCreate(&r.e0);
    cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, on);
    r.h0 = r.h1 = now_ms();
    recs.push_back(r);
  }
  void end() {
    if (!g_profile || recs.empty()) return;
    cudaEventRecord(recs.back().e1, cur);
    recs.back().h1 = now_ms();
  }
  void collect() {   // call after the stream has been synchronised
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : recs) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
        g_phase_ms[r.phase] += ms;
        g_phase_calls[r.phase] += 1;
        g_phase_host_ms[r.phase] += r.h1 - r.h0;
      }
      cudaEventDestroy(r.e0);
      cudaEventDestroy(r.e1);
    }
    recs.clear();
  }
  ~PhaseTimer() { collect(); }
};

// host wall-clock phases of spw_gradient_run (SPW_TIMING=1 prints them)
static double wall_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// ---- per-device context -----------------------------------------------------------------------------
struct DeviceCtx {
  cudaStream_t stream = nullptr;
  std::map<std::tuple<int, int, int>, LinalgPlan*> plans;
  std::map<std::pair<int, int>, DagPlan*> dag_plans;
  std::map<std::pair<int, int>, LaPlan*> la_plans;
  LaStreams la_streams;        // side / bulk streams and events of the look-ahead Cholesky (potrf_la.cu)
  cudaGraphExec_t hlm_exec = nullptr;   // iteration graph of the last hlm call (updated in place by the next one)
  void* ws = nullptr;          // grow-only workspace
  size_t ws_bytes = 0;
  void* ws2 = nullptr;         // grow-only scratch of the summary kernels
  size_t ws2_bytes = 0;
  void* io = nullptr;          // grow-only device staging of the host-pointer entry points (inputs, draws, moments)
  size_t io_bytes = 0;
  void* io2 = nullptr;         // grow-only gather / R-layout buffers on the root device
  size_t io2_bytes = 0;
  Lookahead la;                // second stream + events of the single-matrix Cholesky's lookahead schedule
  cudaStream_t prep = nullptr; // high-priority side stream of the gradient pipeline (factorisations of the next batch)
  cudaEvent_t ev_prep[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr}, ev_start = nullptr;
};
static DeviceCtx g_ctx[64];
static std::mutex g_ctx_mu;

static int get_ctx(DeviceCtx** out) {
  int dev = 0;
  SPW_CUDA(cudaGetDevice(&dev));
  DeviceCtx& c = g_ctx[dev & 63];
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  if (!c.stream) SPW_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  *out = &c;
  return SPW_OK;
}

// lookahead resources for factorising one matrix of order n (nullptr when disabled by SPW_LOOKAHEAD=0)
static int get_lookahead(DeviceCtx& c, int n, Lookahead** out) {
  *out = nullptr;
  // SPW_LOOKAHEAD: 0 plain right-looking schedule; 1 bulk of each trailing update concurrent with the next panel;
  // 2 / 3 (default 3) pieces of the bulk in the shadow of the diagonal-block kernels (see potrf_batched)
  const char* env = getenv("SPW_LOOKAHEAD");
  const int mode = (env && env[0] >= '0' && env[0] <= '3') ? env[0] - '0' : 3;
  if (mode == 0) return SPW_OK;
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  c.la.mode = mode;
  if (!c.la.bulk) {
    int lo = 0, hi = 0;
    SPW_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    SPW_CUDA(cudaStreamCreateWithPriority(&c.la.bulk, cudaStreamNonBlocking, lo));   // lowest priority: bulk work
  }
  const size_t need = (size_t)ceil_div(n, c.la.mode >= 2 ? NB : OUTER_W) + 1;
  while (c.la.ev_chain.size() < need) {
    cudaEvent_t a, b;
    SPW_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    SPW_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    c.la.ev_chain.push_back(a);
    c.la.ev_bulk.push_back(b);
  }
  while (c.la.ev_aux.size() < (size_t)ceil_div(n, OUTER_W) + 1) {
    cudaEvent_t a;
    SPW_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    c.la.ev_aux.push_back(a);
  }
  *out = &c.la;
  return SPW_OK;
}

static int get_plan(DeviceCtx& c, int n, int extra, int single, const LinalgPlan** out) {
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  auto key = std::make_tuple(n, extra, single);
  auto it = c.plans.find(key);
  if (it == c.plans.end()) {
    LinalgPlan* p = new LinalgPlan();
    int rc = plan_build(*p, n, extra, single);
    if (rc != SPW_OK) {
      delete p;
      return rc;
    }
    it = c.plans.emplace(key, p).first;
  }
  *out = it->second;
  return SPW_OK;
}

// Persistent tile-DAG Cholesky plan for one matrix of order n (nullptr: use the multi-launch schedule).
// SPW_DAG=0 disables it, SPW_DAG_MIN sets the smallest order it is used for.
static int get_dag_plan(DeviceCtx& c, int n, int extra, const DagPlan** out) {
  *out = nullptr;
  const char* env = getenv("SPW_DAG");
  if (!(env && env[0] == '1')) return SPW_OK;      // work in progress: opt-in
  const char* env_min = getenv("SPW_DAG_MIN");
  const int nmin = env_min ? atoi(env_min) : 256;
  if (n < nmin) return SPW_OK;
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  auto key = std::make_pair(n, extra);
  auto it = c.dag_plans.find(key);
  if (it == c.dag_plans.end()) {
    DagPlan* p = new DagPlan();
    int rc = dag_plan_build(*p, n, extra);
    if (rc != SPW_OK) {
      delete p;
      return rc;
    }
    it = c.dag_plans.emplace(key, p).first;
  }
  *out = it->second;
  return SPW_OK;
}

// Look-ahead schedule for one matrix of order n (nullptr: use potrf_batched): the default from order 1024 on, where
// it is 4-8 % faster than the shadow schedule.  SPW_LA=0 disables it, SPW_LA_MIN sets the smallest order it is used for.
static int get_la_plan(DeviceCtx& c, int n, int extra, const LaPlan** out, LaStreams** streams) {
  *out = nullptr;
  *streams = nullptr;
  const char* env = getenv("SPW_LA");
  if (env && env[0] == '0') return SPW_OK;
  const char* env_min = getenv("SPW_LA_MIN");
  const int nmin = env_min ? atoi(env_min) : 1024;
  if (n < nmin) return SPW_OK;
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  auto key = std::make_pair(n, extra);
  auto it = c.la_plans.find(key);
  if (it == c.la_plans.end()) {
    LaPlan* p = new LaPlan();
    int rc = la_plan_build(*p, n, extra);
    if (rc != SPW_OK) {
      delete p;
      return rc;
    }
    it = c.la_plans.emplace(key, p).first;
  }
  SPW_TRY(la_streams_reserve(c.la_streams, it->second->nblk));
  *out = it->second;
  *streams = &c.la_streams;
  return SPW_OK;
}

static int get_workspace(DeviceCtx& c, size_t bytes, void** out) {
  if (bytes > c.ws_bytes) {
    if (c.ws) SPW_CUDA(cudaFree(c.ws));
    c.ws = nullptr;
    c.ws_bytes = 0;
    SPW_CUDA(cudaMalloc(&c.ws, bytes));
    c.ws_bytes = bytes;
  }
  *out = c.ws;
  return SPW_OK;
}

static int grow_buffer(void** p, size_t* have, size_t bytes) {
  if (bytes > *have) {
    if (*p) SPW_CUDA(cudaFree(*p));
    *p = nullptr;
    *have = 0;
    SPW_CUDA(cudaMalloc(p, bytes));
    *have = bytes;
  }
  return SPW_OK;
}

static int get_scratch(DeviceCtx& c, size_t bytes, void** out) {
  if (bytes > c.ws2_bytes) {
    if (c.ws2) SPW_CUDA(cudaFree(c.ws2));
    c.ws2 = nullptr;
    c.ws2_bytes = 0;
    SPW_CUDA(cudaMalloc(&c.ws2, bytes));
    c.ws2_bytes = bytes;
  }
  *out = c.ws2;
  return SPW_OK;
}

static size_t workspace_budget() {
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return size_t(8) << 30;
  int dev = 0;
  cudaGetDevice(&dev);
  free_b += g_ctx[dev & 63].ws_bytes;   // our cached workspace is reusable
  size_t budget = (size_t)(0.6 * (double)free_b);
  if (g_ws_limit > 0 && (size_t)g_ws_limit < budget) budget = (size_t)g_ws_limit;
  return budget;
}

struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base((char*)p) {}
  template <class T>
  T* take(size_t count) {
    off = (off + 255) / 256 * 256;
    T* r = (T*)(base + off);
    off += count * sizeof(T);
    return r;
  }
};

static int ncomp_of(int type) { return type == SPW_COV_MATERN1 ? 2 : 5; }
static long long pad_ld(long long n) { return round_up_ll(n, 16); }

// ---- the per-shard spatial_gradient driver (device pointers) -------------------------------------------
// Batches of B samples go through two stages: "prep" (covariance assembly, batched Cholesky, triangular inverse,
// v = Linv z -- many short, latency-bound kernels) and "main" (cross-covariances + the fused TRMM / Gram / draw kernel).
// Optionally (SPW_OVERLAP=1) the stages run on two streams with double-buffered factors, so that the prep of batch k+1
// runs while the main kernel of batch k owns the SMs; see the note at `overlap` below for why this is not the default.
// Status words live in one array for the whole shard and are read once at the end: no host sync per batch.
static int gradient_shard(int type, int n, const double* cx, const double* cy, int G, const double* gx,
                          const double* gy, int S, const double* phi, const double* sig2, const double* z, int ldz,
                          const double* eps, double* out_draws, double* out_mean, double* out_var, int* info,
                          cudaStream_t st) {
  if (info) info[0] = info[1] = info[2] = info[3] = 0;
  if (type < 0 || type > 2 || n < 1 || G < 1 || S < 0) { set_error("spw_gradient: bad arguments"); return SPW_ERR_ARG; }
  if (out_draws && !eps) { set_error("spw_gradient: eps is required when draws are requested"); return SPW_ERR_ARG; }
  if (S == 0) return SPW_OK;
  DeviceCtx* ctx;
  SPW_TRY(get_ctx(&ctx));
  // Measured on B200 (C5, 64 samples per call): 2506 ms with the two-stream pipeline against 2515 ms without -- the
  // main kernel's CTAs run 8.5 ms each and retire in waves, so a dependent chain of ~600 short prep kernels only finds
  // free SMs once per wave; and since the fused kernel already keeps the FP64 pipe 96 % busy, the most an overlap can
  // recover is the idle share of the factorisations (< 3 % of a step).  Off by default (SPW_OVERLAP=1 enables it).
  static const bool overlap = getenv("SPW_OVERLAP") && getenv("SPW_OVERLAP")[0] == '1';
  const int nset = overlap ? 2 : 1;
  const int c = ncomp_of(type);
  const long long ld = pad_ld(n);
  const int nblk = ceil_div(n, NB);
  // workspace sizing: batch B samples, grid chunk gc
  const size_t mat = (size_t)n * ld * sizeof(double);
  const size_t per_sample_fixed =
      nset * (2 * mat + (size_t)ld * sizeof(double) + (size_t)nblk * NB * NB * sizeof(double) + 64);
  const size_t at_per_g = (size_t)c * ld * sizeof(double);
  // the (B, gc) decision is cached per problem shape so that repeated calls skip cudaMemGetInfo
  struct ShapeKey { int dev, type, n, G, S; size_t limit; int B, gc; };
  static thread_local ShapeKey last = {-1, -1, 0, 0, 0, 0, 0, 0};
  int cur_dev = 0;
  SPW_CUDA(cudaGetDevice(&cur_dev));
  // reuse only on the same device and only while the cached choice still fits the workspace this context already owns
  // (otherwise the budget is consulted again: other allocations may have grown since)
  const size_t status_bytes = sizeof(int) * 2 * (size_t)S + 256;
  const bool reuse = (last.dev == cur_dev && last.type == type && last.n == n && last.G == G && last.S == S &&
                      last.limit == (size_t)g_ws_limit &&
                      ctx->ws_bytes >= (per_sample_fixed + at_per_g * (size_t)last.gc) * (size_t)last.B + status_bytes + 4096);
  const size_t budget = reuse ? 0 : workspace_budget();
  int gc = std::min(G, 65535);
  int B = (int)std::min<size_t>((size_t)std::min(S, 1024), budget / (per_sample_fixed + at_per_g * gc));
  // beyond ~24 GB a larger batch buys no efficiency and only makes the allocation slow
  const size_t soft_cap = (size_t)24 << 30;
  if (B > 16 && (per_sample_fixed + at_per_g * gc) * (size_t)B > soft_cap)
    B = std::max<int>(16, (int)(soft_cap / (per_sample_fixed + at_per_g * gc)));
  if (reuse) {
    B = last.B;
    gc = last.gc;
  } else if (B < std::min(S, 16)) {   // not enough room for the whole grid at a useful batch size: chunk the grid
    B = std::min(S, 16);
    while (B > 1 && per_sample_fixed * B > budget / 2) B /= 2;
    const size_t left = budget > per_sample_fixed * B ? budget - per_sample_fixed * B : 0;
    gc = (int)std::min<size_t>((size_t)gc, left / (at_per_g * B));
    gc = gc / 16 * 16;
    if (gc < 16) { set_error("spw_gradient: workspace too small (n = %d)", n); return SPW_ERR_ARG; }
  }
  // at least two batches, so that the second one's factorisations hide behind the first one's main kernel
  if (!reuse && overlap && S >= 8 && B > (S + 1) / 2) B = (S + 1) / 2;
  last = ShapeKey{cur_dev, type, n, G, S, (size_t)g_ws_limit, B, gc};
  const LinalgPlan* plan;
  SPW_TRY(get_plan(*ctx, n, 0, B < 4 ? 1 : 0, &plan));
  const size_t total = (per_sample_fixed + at_per_g * gc) * B + status_bytes + 4096;
  void* wsp;
  SPW_TRY(get_workspace(*ctx, total, &wsp));
  Carver cv(wsp);
  double *Lb[2], *Li[2], *vb[2], *dinv[2];
  for (int i = 0; i < nset; ++i) {
    Lb[i] = cv.take<double>((size_t)B * n * ld);
    Li[i] = cv.take<double>((size_t)B * n * ld);
    vb[i] = cv.take<double>((size_t)B * ld);
    dinv[i] = cv.take<double>((size_t)B * nblk * NB * NB);
  }
  double* At = cv.take<double>((size_t)B * gc * c * ld);
  int* status = cv.take<int>(2 * (size_t)S);      // [0, S): chol(K); [S, 2S): the 5 x 5 factorisations
  MatView AT{At, ld, (long long)gc * c * ld};
  // side stream and events of the two-stage pipeline
  cudaStream_t ps = st;
  if (overlap) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (!ctx->prep) {
      int lo = 0, hi = 0;
      SPW_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      SPW_CUDA(cudaStreamCreateWithPriority(&ctx->prep, cudaStreamNonBlocking, hi));
      for (int i = 0; i < 2; ++i) {
        SPW_CUDA(cudaEventCreateWithFlags(&ctx->ev_prep[i], cudaEventDisableTiming));
        SPW_CUDA(cudaEventCreateWithFlags(&ctx->ev_free[i], cudaEventDisableTiming));
      }
      SPW_CUDA(cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming));
    }
    ps = ctx->prep;
  }
  SPW_CUDA(cudaMemsetAsync(status, 0, sizeof(int) * 2 * (size_t)S, st));
  if (overlap) {
    SPW_CUDA(cudaEventRecord(ctx->ev_start, st));          // the side stream starts after the caller's earlier work
    SPW_CUDA(cudaStreamWaitEvent(ps, ctx->ev_start, 0));
  }
  PhaseTimer pt(st);
  int k = 0;
  for (int s0 = 0; s0 < S; s0 += B, ++k) {
    const int nb = std::min(B, S - s0), set = overlap ? (k & 1) : 0;
    MatView L{Lb[set], ld, (long long)n * ld}, Linv{Li[set], ld, (long long)n * ld};
    // ---- prep: K = cov(phi_s, sig2 = 1), no nugget; chol(K); its triangular inverse; v = Linv z
    //      R/spatial_gradient.R:215-216,264-265,303-304 ----
    if (overlap && k >= 2) SPW_CUDA(cudaStreamWaitEvent(ps, ctx->ev_free[set], 0));   // main of batch k-2 is done with this set
    pt.begin(PH_COV, ps);
    SPW_TRY(cov_assemble_batched(type, n, cx, cy, phi + s0, nullptr, nullptr, 1, L, nb, 1, ps));
    pt.end();
    pt.begin(PH_POTRF, ps);
    SPW_TRY(potrf_batched(*plan, L, dinv[set], &Linv, nb, status + s0, ps));
    pt.end();
    pt.begin(PH_TRTRI, ps);
    SPW_TRY(trtri_batched(*plan, L, Linv, nb, ps));
    pt.end();
    pt.begin(PH_TRMV, ps);
    SPW_TRY(trmv_lower_batched(n, Linv, z + (long long)s0 * ldz, ldz, vb[set], ld, nb, ps));
    pt.end();
    if (overlap) {
      SPW_CUDA(cudaEventRecord(ctx->ev_prep[set], ps));
      SPW_CUDA(cudaStreamWaitEvent(st, ctx->ev_prep[set], 0));
    }
    // ---- main: cross-covariances and the fused kernel for every grid chunk ----
    for (int g0 = 0; g0 < G; g0 += gc) {
      const int ng = std::min(gc, G - g0);
      pt.begin(PH_CROSS, st);
      SPW_TRY(cross_cov_batched(type, n, cx, cy, g0, ng, gx, gy, phi + s0, AT, nb, st));
      pt.end();
      GradParams P;
      P.linv = Linv;
      P.at = AT;
      P.v = vb[set];
      P.ldv = ld;
      P.phi = phi + s0;
      P.sig2 = sig2 + s0;
      P.eps = eps;
      P.out_draws = out_draws;
      P.out_mean = out_mean;
      P.out_var = out_var;
      P.status = status + S + s0;
      P.n = n;
      P.g_begin = g0;
      P.g_count = ng;
      P.g_total = G;
      P.s_begin = s0;
      pt.begin(PH_MAIN, st);
      SPW_TRY(gradient_tma_launch(type, P, nb, st));
      pt.end();
      if (g_profile) {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_main_flops += (double)nb * ng * ((double)c * n * ((double)n + 1.0) + (double)(c * (c + 1) + 2 * c) * n);
      }
    }
    if (overlap) SPW_CUDA(cudaEventRecord(ctx->ev_free[set], st));
  }
  std::vector<int> h_status(2 * (size_t)S);
  SPW_CUDA(cudaMemcpyAsync(h_status.data(), status, sizeof(int) * 2 * (size_t)S, cudaMemcpyDeviceToHost, st));
  SPW_CUDA(cudaStreamSynchronize(st));
  pt.collect();
  for (int s = 0; s < S; ++s) {
    if (h_status[s]) {
      if (info) { info[0] = s + 1; info[1] = h_status[s]; info[2] = 1; }
      set_error("spatial_gradient: sample %d: the leading minor of order %d of K is not positive", s + 1, h_status[s]);
      return SPW_ERR_NOT_PD;
    }
    if (h_status[S + s]) {
      if (info) { info[0] = s + 1; info[1] = h_status[S + s]; info[2] = 2; }
      set_error("spatial_gradient: sample %d, grid point %d: var.grad is not positive definite", s + 1, h_status[S + s]);
      return SPW_ERR_NOT_PD;
    }
  }
  return SPW_OK;
}

// ---- spsample: Gibbs recovery of z for a batch of (phi, sig2, tau2) draws (device pointers) -----------------
// R/spsample_beta.R:50-100 without the p x p regression block (done by the host):
//   R.inv.Z = chol2inv(chol(R.Z))            -> Linv^T Linv from the triangular inverse        (:70)
//   Sig.Z.in = R.inv.Z/sig2 + I/tau2         -> scale_add_diag                                 (:83)
//   Sig.Z = chol2inv(chol(Sig.Z.in))         -> second factorisation + inverse + Gram          (:84,91)
//   mu.Z = crossprod(Sig.Z, w)               -> gemv (w = (y - X beta)/tau2 from the host)     (:92)
//   z = crossprod(chol(Sig.Z), eps) + mu.Z   -> third factorisation, lower-triangular matvec   (:93)
// jitter[b] is added to the diagonal of R.Z (0, or 1e-4 on the retry of :71).
// status_k[b] != 0: chol(R.Z) failed; status_bad[b] != 0: chol(Sig.Z.in) failed ("Bad Iteration", :85).
static int spsample_shard(int type, int n, const double* cx, const double* cy, int S, const double* phi,
                          const double* inv_sig2, const double* inv_tau2, const double* jitter, const double* w,
                          const double* eps, long long ldv, double* z_out, int* h_status_k, int* h_status_bad,
                          cudaStream_t st) {
  if (S == 0) return SPW_OK;
  DeviceCtx* ctx;
  SPW_TRY(get_ctx(&ctx));
  const long long ld = pad_ld(n);
  const int nblk = ceil_div(n, NB);
  const size_t per_sample = 2 * (size_t)n * ld * sizeof(double) + (size_t)ld * sizeof(double) +
                            (size_t)nblk * NB * NB * sizeof(double) + 64;
  const size_t budget = workspace_budget();
  int B = (int)std::min<size_t>((size_t)std::min(S, 512), budget / per_sample);
  if (B < 1) { set_error("spw_spsample: workspace too small (n = %d)", n); return SPW_ERR_ARG; }
  const LinalgPlan* plan;
  SPW_TRY(get_plan(*ctx, n, 0, B < 4 ? 1 : 0, &plan));
  void* wsp;
  SPW_TRY(get_workspace(*ctx, per_sample * B + 4096, &wsp));
  Carver cv(wsp);
  double* Lb = cv.take<double>((size_t)B * n * ld);
  double* Li = cv.take<double>((size_t)B * n * ld);
  double* mu = cv.take<double>((size_t)B * ld);
  double* dinv = cv.take<double>((size_t)B * nblk * NB * NB);
  int* status = cv.take<int>((size_t)3 * B);
  MatView L{Lb, ld, (long long)n * ld}, Linv{Li, ld, (long long)n * ld};
  std::vector<int> hs(3 * (size_t)B);
  for (int s0 = 0; s0 < S; s0 += B) {
    const int nb = std::min(B, S - s0);
    SPW_CUDA(cudaMemsetAsync(status, 0, sizeof(int) * 3 * B, st));
    static const bool timing = getenv("SPW_TIMING") && getenv("SPW_TIMING")[0] == '1';
    double tt[12];
    int ti = 0;
    auto tick = [&]() { if (timing) { cudaStreamSynchronize(st); tt[ti++] = wall_ms(); } };
    tick();
    SPW_TRY(cov_assemble_batched(type, n, cx, cy, phi + s0, nullptr, jitter + s0, 1, L, nb, 1, st));
    tick();
    SPW_TRY(potrf_batched(*plan, L, dinv, &Linv, nb, status, st));
    tick();
    SPW_TRY(trtri_batched(*plan, L, Linv, nb, st));
    tick();
    SPW_TRY(gram_of_inverse_batched(*plan, Linv, L, nb, st));                       // R.inv.Z
    tick();
    SPW_TRY(scale_add_diag_batched(n, L, inv_sig2 + s0, inv_tau2 + s0, nb, st));    // Sig.Z.in
    tick();
    SPW_TRY(potrf_batched(*plan, L, dinv, &Linv, nb, status + B, st));
    tick();
    SPW_TRY(trtri_batched(*plan, L, Linv, nb, st));
    tick();
    SPW_TRY(gram_of_inverse_batched(*plan, Linv, L, nb, st));                       // Sig.Z
    tick();
    SPW_TRY(gemv_batched(n, L, w + (long long)s0 * ldv, ldv, nullptr, 0, mu, ld, 0, nb, st));   // mu.Z
    tick();
    SPW_TRY(potrf_batched(*plan, L, dinv, nullptr, nb, status + 2 * B, st));        // chol(Sig.Z)
    tick();
    if (timing)
      fprintf(stderr, "[spw] spsample batch of %d (n = %d): cov %.2f potrf %.2f trtri %.2f gram %.2f scale %.2f potrf %.2f trtri "
              "%.2f gram %.2f gemv %.2f potrf %.2f ms\n", nb, n, tt[1] - tt[0], tt[2] - tt[1], tt[3] - tt[2], tt[4] - tt[3],
              tt[5] - tt[4], tt[6] - tt[5], tt[7] - tt[6], tt[8] - tt[7], tt[9] - tt[8], tt[10] - tt[9]);
    // z = t(chol(Sig.Z)) eps + mu.Z: the upper factor's transpose is this library's lower factor
    SPW_TRY(gemv_batched(n, L, eps + (long long)s0 * ldv, ldv, mu, ld, z_out + (long long)s0 * ldv, ldv, 1, nb, st));
    SPW_CUDA(cudaMemcpyAsync(hs.data(), status, sizeof(int) * 3 * B, cudaMemcpyDeviceToHost, st));
    SPW_CUDA(cudaStreamSynchronize(st));
    for (int b = 0; b < nb; ++b) {
      h_status_k[s0 + b] = hs[b];
      h_status_bad[s0 + b] = hs[b] ? 0 : (hs[B + b] ? hs[B + b] : hs[2 * B + b]);
    }
  }
  return SPW_OK;
}

// ---- NCCL, loaded at run time (dlopen) so that the library links against libcudart only and never clashes with
//      the copy of NCCL a host process (PyTorch) may already have mapped ------------------------------------------
struct NcclApi {
  typedef struct ncclComm* comm_t;
  int (*CommInitAll)(comm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(comm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
static NcclApi g_nccl;
static int g_last_gather_nccl = -1;   // 1: the last multi-GPU gather went through NCCL, 0: peer copies, -1: none yet
static std::vector<NcclApi::comm_t> g_nccl_comms;   // communicator clique over devices 0 .. n-1 (cached)
static std::mutex g_nccl_mu;

static bool nccl_load() {
  static bool tried = false;
  if (tried) return g_nccl.ok;
  tried = true;
  const char* env = getenv("SPW_NCCL");
  if (env && env[0] == '0') return false;
  // SPW_NCCL_PATH: an explicit library (the Python host points it at the copy PyTorch bundles, so that a later
  // `import torch` in the same process finds the NCCL it was built against under the shared soname); then whatever is
  // already mapped; then the system library
  void* h = nullptr;
  const char* path = getenv("SPW_NCCL_PATH");
  if (path && path[0]) h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!h) return false;
  g_nccl.CommInitAll = (int (*)(NcclApi::comm_t*, int, const int*))dlsym(h, "ncclCommInitAll");
  g_nccl.CommDestroy = (int (*)(NcclApi::comm_t))dlsym(h, "ncclCommDestroy");
  g_nccl.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
  g_nccl.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
  g_nccl.Send = (int (*)(const void*, size_t, int, int, NcclApi::comm_t, cudaStream_t))dlsym(h, "ncclSend");
  g_nccl.Recv = (int (*)(void*, size_t, int, int, NcclApi::comm_t, cudaStream_t))dlsym(h, "ncclRecv");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  g_nccl.ok = g_nccl.CommInitAll && g_nccl.CommDestroy && g_nccl.GroupStart && g_nccl.GroupEnd && g_nccl.Send &&
              g_nccl.Recv;
  return g_nccl.ok;
}

// The path's one exchange step as an NCCL collective over NVLink: every device d > 0 sends its [S_d][c][g] block,
// device 0 receives them into the sample-major gather buffer (grouped ncclSend / ncclRecv, datatype ncclFloat64 = 8).
// Returns SPW_OK, or a negative value when NCCL is unavailable (the caller then uses peer copies).
// communicator clique over devices 0 .. ngpu-1 (cached); false when NCCL is unavailable.  Called from a side thread
// at the start of spw_gradient_run so that the set-up overlaps the shards' compute.
static bool nccl_prepare(int ngpu) {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (!nccl_load()) return false;
  if ((int)g_nccl_comms.size() != ngpu) {
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto c : g_nccl_comms)
      if (c) g_nccl.CommDestroy(c);
    g_nccl_comms.assign(ngpu, nullptr);
    std::vector<int> devs(ngpu);
    for (int d = 0; d < ngpu; ++d) devs[d] = d;
    const int rc = g_nccl.CommInitAll(g_nccl_comms.data(), ngpu, devs.data());
    cudaSetDevice(cur);
    if (rc != 0) {
      g_nccl_comms.clear();
      return false;
    }
  }
  return true;
}

static int nccl_gather(int ngpu, const std::vector<const double*>& src, const std::vector<size_t>& count,
                       const std::vector<double*>& dst_on_root) {
  if (!nccl_prepare(ngpu)) return -1;
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  int rc = g_nccl.GroupStart();
  for (int d = 1; d < ngpu && rc == 0; ++d) {
    if (count[d] == 0) continue;
    cudaSetDevice(0);
    rc = g_nccl.Recv(dst_on_root[d], count[d], 8 /* ncclFloat64 */, d, g_nccl_comms[0], g_ctx[0].stream);
    if (rc != 0) break;
    cudaSetDevice(d);
    rc = g_nccl.Send(src[d], count[d], 8, 0, g_nccl_comms[d], g_ctx[d].stream);
  }
  const int rc2 = g_nccl.GroupEnd();
  cudaSetDevice(0);
  if (rc != 0 || rc2 != 0) {
    set_error("NCCL gather failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc ? rc : rc2) : "?");
    return SPW_ERR_CUDA;
  }
  for (int d = 1; d < ngpu; ++d) {       // the sends run on the senders' streams
    cudaSetDevice(d);
    if (cudaStreamSynchronize(g_ctx[d].stream) != cudaSuccess) return SPW_ERR_CUDA;
  }
  cudaSetDevice(0);
  return SPW_OK;
}

// small RAII helper for temporary device buffers of the host-pointer entry points
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t bytes) {
    SPW_CUDA(cudaMalloc(&p, bytes ? bytes : 8));
    return SPW_OK;
  }
  template <class T> T* as() { return (T*)p; }
};

}  // namespace spw

using namespace spw;

extern "C" {

int spw_version(void) { return 100; }

const char* spw_last_error(void) {
  if (tl_error[0]) return tl_error;
  return g_error;
}

int spw_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
  }
  return n;
}

int spw_set_device(int device) {
  SPW_CUDA(cudaSetDevice(device));
  return SPW_OK;
}

int spw_set_workspace_limit(long long bytes) {
  g_ws_limit = bytes;
  return SPW_OK;
}

long long spw_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

int spw_last_gather_used_nccl(void) { return g_last_gather_nccl; }

int spw_profile_enable(int on) {
  g_profile = on;
  return SPW_OK;
}

int spw_profile_read(double* ms, long long* calls, double* main_flops, int reset) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < PH_COUNT; ++i) {
    if (ms) ms[i] = g_phase_ms[i];
    if (calls) calls[i] = g_phase_calls[i];
    if (getenv("SPW_PROFILE_VERBOSE") && g_phase_calls[i])
      fprintf(stderr, "[spw] phase %d: gpu %.3f ms, host issue %.3f ms, %lld groups\n", i, g_phase_ms[i],
              g_phase_host_ms[i], g_phase_calls[i]);
    if (reset) {
      g_phase_ms[i] = 0.0;
      g_phase_calls[i] = 0;
      g_phase_host_ms[i] = 0.0;
    }
  }
  if (main_flops) *main_flops = g_main_flops;
  if (reset) g_main_flops = 0.0;
  return SPW_OK;
}

// diagnostic: cycle counters of the last diagonal-block kernel: [0] load, [1] end of the factor + inverse sweep,
// [4] end of kernel
int spw_debug_diag_timing(long long* out16) {
  if (!g_diag_dbg) {
    SPW_CUDA(cudaGetDevice(&g_diag_dbg_device));      // the counters live on (and are only handed to kernels of) this device
    SPW_CUDA(cudaMalloc(&g_diag_dbg, 16 * sizeof(long long)));
    SPW_CUDA(cudaMemset(g_diag_dbg, 0, 16 * sizeof(long long)));
    return SPW_OK;
  }
  SPW_CUDA(cudaDeviceSynchronize());
  SPW_CUDA(cudaMemcpy(out16, g_diag_dbg, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  return SPW_OK;
}

// diagnostic (SPW_DAG_TRACE=1): per-task trace of the last persistent-Cholesky run of order n (extra appended rows) on the
// current device: 4 words per task (packed id, start ns, end ns, 2 * SM + sub-worker).  Returns the words written.
long long spw_debug_dag_trace(int n, int extra, unsigned long long* out, long long cap_words) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  DeviceCtx& c = g_ctx[dev & 63];
  auto it = c.dag_plans.find(std::make_pair(n, extra));
  if (it == c.dag_plans.end()) return 0;
  cudaDeviceSynchronize();
  std::vector<unsigned long long> v;
  if (dag_trace_read(*it->second, v) != SPW_OK) return -1;
  const long long cnt = std::min<long long>((long long)v.size(), cap_words);
  if (out && cnt > 0) memcpy(out, v.data(), sizeof(unsigned long long) * cnt);
  return cnt;
}

// diagnostic: average time (ms, CUDA events) of one single-matrix factorisation of order n (+ extra appended rows) under the
// schedule the environment selects (SPW_DAG / SPW_LA / SPW_LOOKAHEAD), replayed `reps` times as a CUDA graph on a
// synthetic positive definite matrix.  info[0] = status of the last factorisation.
__global__ void debug_fill_kernel(double* __restrict__ a, long long ld, int rows, int n) {
  const int i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows || j >= n) return;
  const int d = i > j ? i - j : j - i;
  a[(long long)i * ld + j] = (i < n ? exp(-0.02 * d) + (d == 0 ? 1.0 : 0.0) : 1.0);
}

// diagnostic (SPW_LA_TRACE=1): 8 words per diagonal step of the last look-ahead factorisation of order n (globaltimer
// ns): chain wait begins, step begins, step published, PANEL2(j) first start / last end, LLS(j) first start / last end.
// Returns the words written.
long long spw_debug_la_trace(int n, int extra, unsigned long long* out, long long cap_words) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  DeviceCtx& c = g_ctx[dev & 63];
  auto it = c.la_plans.find(std::make_pair(n, extra));
  if (it == c.la_plans.end()) return 0;
  cudaDeviceSynchronize();
  std::vector<unsigned long long> v;
  if (la_trace_read(*it->second, v) != SPW_OK) return -1;
  const long long cnt = std::min<long long>((long long)v.size(), cap_words);
  if (out && cnt > 0) memcpy(out, v.data(), sizeof(unsigned long long) * cnt);
  return cnt;
}

int spw_debug_potrf_time(int n, int extra, int reps, double* ms_out, int* info) {
  DeviceCtx* ctx;
  SPW_TRY(get_ctx(&ctx));
  cudaStream_t st = ctx->stream;
  const LinalgPlan* plan;
  SPW_TRY(get_plan(*ctx, n, extra, 1, &plan));
  const long long ld = pad_ld(n);
  const int rows = n + extra;
  DevBuf dm, d0, dd, ds;
  SPW_TRY(dm.alloc(sizeof(double) * rows * ld));
  SPW_TRY(d0.alloc(sizeof(double) * rows * ld));
  SPW_TRY(dd.alloc(sizeof(double) * plan->nblk * NB * NB));
  SPW_TRY(ds.alloc(sizeof(int)));
  debug_fill_kernel<<<dim3(ceil_div(n, 256), rows), 256, 0, st>>>(d0.as<double>(), ld, rows, n);
  SPW_LAUNCHED();
  MatView L{dm.as<double>(), ld, (long long)rows * ld};
  const DagPlan* dag;
  SPW_TRY(get_dag_plan(*ctx, n, extra, &dag));
  const LaPlan* lap;
  LaStreams* las;
  SPW_TRY(get_la_plan(*ctx, n, extra, &lap, &las));
  Lookahead* la;
  SPW_TRY(get_lookahead(*ctx, n, &la));
  const bool eager = reps < 0;             // reps < 0: plain stream launches instead of a CUDA graph
  if (eager) reps = -reps;
  auto factor = [&]() -> int {
    if (dag) return potrf_dag(*dag, L, dd.as<double>(), ds.as<int>(), st);
    if (lap) return potrf_lookahead(*lap, *las, L, dd.as<double>(), ds.as<int>(), st);
    return potrf_batched(*plan, L, dd.as<double>(), nullptr, 1, ds.as<int>(), st, la);
  };
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  if (!eager) {
    SPW_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = factor();
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc != SPW_OK) return rc;
    SPW_CUDA(ce);
    SPW_CUDA(cudaGraphInstantiate(&exec, graph, 0));
  }
  std::vector<cudaEvent_t> ev(2 * (size_t)reps);
  for (auto& e : ev) SPW_CUDA(cudaEventCreate(&e));
  for (int r = -2; r < reps; ++r) {
    SPW_CUDA(cudaMemcpyAsync(dm.p, d0.p, sizeof(double) * rows * ld, cudaMemcpyDeviceToDevice, st));
    SPW_CUDA(cudaMemsetAsync(ds.p, 0, sizeof(int), st));
    if (r >= 0) SPW_CUDA(cudaEventRecord(ev[2 * r], st));
    if (eager) SPW_TRY(factor());
    else SPW_CUDA(cudaGraphLaunch(exec, st));
    if (r >= 0) SPW_CUDA(cudaEventRecord(ev[2 * r + 1], st));
    if (eager) SPW_CUDA(cudaStreamSynchronize(st));
  }
  SPW_CUDA(cudaStreamSynchronize(st));
  double total = 0.0;
  for (int r = 0; r < reps; ++r) {
    float ms = 0.f;
    SPW_CUDA(cudaEventElapsedTime(&ms, ev[2 * r], ev[2 * r + 1]));
    total += ms;
  }
  for (auto& e : ev) cudaEventDestroy(e);
  if (exec) cudaGraphExecDestroy(exec);
  if (graph) cudaGraphDestroy(graph);
  if (ms_out) *ms_out = total / std::max(1, reps);
  if (info) SPW_CUDA(cudaMemcpy(info, ds.p, sizeof(int), cudaMemcpyDeviceToHost));
  return SPW_OK;
}

int spw_shutdown(void) {
  int cur = 0, ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) return SPW_OK;
  cudaGetDevice(&cur);
  for (int d = 0; d < ndev && d < 64; ++d) {
    DeviceCtx& c = g_ctx[d];
    if (!c.stream && !c.ws && !c.io && !c.io2 && c.plans.empty() && c.dag_plans.empty()) continue;
    cudaSetDevice(d);
    for (auto& kv : c.plans) {
      plan_free(*kv.second);
      delete kv.second;
    }
    c.plans.clear();
    for (auto& kv : c.dag_plans) {
      dag_plan_free(*kv.second);
      delete kv.second;
    }
    c.dag_plans.clear();
    for (auto& kv : c.la_plans) {
      la_plan_free(*kv.second);
      delete kv.second;
    }
    c.la_plans.clear();
    if (c.la_streams.side) cudaStreamDestroy(c.la_streams.side);
    if (c.la_streams.side2) cudaStreamDestroy(c.la_streams.side2);
    if (c.la_streams.ev_join2) cudaEventDestroy(c.la_streams.ev_join2);
    if (c.la_streams.side2b) cudaStreamDestroy(c.la_streams.side2b);
    if (c.la_streams.ev_join2b) cudaEventDestroy(c.la_streams.ev_join2b);
    if (c.la_streams.side3) cudaStreamDestroy(c.la_streams.side3);
    if (c.la_streams.ev_join3) cudaEventDestroy(c.la_streams.ev_join3);
    for (int i = 0; i < LaStreams::NBULK; ++i) {
      if (c.la_streams.bulk[i]) cudaStreamDestroy(c.la_streams.bulk[i]);
      if (c.la_streams.ev_bjoin[i]) cudaEventDestroy(c.la_streams.ev_bjoin[i]);
    }
    for (auto* v : {&c.la_streams.ev_p, &c.la_streams.ev_g, &c.la_streams.ev_s})
      for (auto e : *v) cudaEventDestroy(e);
    if (c.la_streams.ev_fork) cudaEventDestroy(c.la_streams.ev_fork);
    if (c.la_streams.ev_join) cudaEventDestroy(c.la_streams.ev_join);
    c.la_streams = LaStreams();
    if (c.hlm_exec) cudaGraphExecDestroy(c.hlm_exec);
    c.hlm_exec = nullptr;
    if (c.ws) cudaFree(c.ws);
    c.ws = nullptr;
    c.ws_bytes = 0;
    if (c.ws2) cudaFree(c.ws2);
    c.ws2 = nullptr;
    c.ws2_bytes = 0;
    if (c.io) cudaFree(c.io);
    c.io = nullptr;
    c.io_bytes = 0;
    if (c.io2) cudaFree(c.io2);
    c.io2 = nullptr;
    c.io2_bytes = 0;
    if (c.stream) cudaStreamDestroy(c.stream);
    c.stream = nullptr;
    if (c.la.bulk) cudaStreamDestroy(c.la.bulk);
    c.la.bulk = nullptr;
    if (c.prep) {
      cudaStreamDestroy(c.prep);
      c.prep = nullptr;
      for (int i = 0; i < 2; ++i) {
        cudaEventDestroy(c.ev_prep[i]);
        cudaEventDestroy(c.ev_free[i]);
      }
      cudaEventDestroy(c.ev_start);
    }
    for (auto e : c.la.ev_chain) cudaEventDestroy(e);
    for (auto e : c.la.ev_bulk) cudaEventDestroy(e);
    for (auto e : c.la.ev_aux) cudaEventDestroy(e);
    c.la.ev_aux.clear();
    c.la.ev_chain.clear();
    c.la.ev_bulk.clear();
  }
  {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.ok)
      for (auto c : g_nccl_comms)
        if (c) g_nccl.CommDestroy(c);
    g_nccl_comms.clear();
  }
  tensor_map_cache_clear();
  cudaSetDevice(cur);
  return SPW_OK;
}

int spw_cov(int type, int n, const double* coords, double phi, double sig2, double* out) {
  if (type < 0 || type > 2 || n < 1 || !coords || !out) { set_error("spw_cov: bad arguments"); return SPW_ERR_ARG; }
  DeviceCtx* ctx;
  SPW_TRY(get_ctx(&ctx));
  cudaStream_t st = ctx->stream;
  const long long ld = pad_ld(n);
  DevBuf dc, dp, dm;
  SPW_TRY(dc.alloc(sizeof(double) * 2 * n));
  SPW_TRY(dp.alloc(sizeof(double) * 2));
  SPW_TRY(dm.alloc(sizeof(double) * n * ld));
  const double par[2] = {phi, sig2};
  SPW_CUDA(cudaMemcpyAsync(dc.p, coords, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, st));
  SPW_CUDA(cudaMemcpyAsync(dp.p, par, sizeof(par), cudaMemcpyHostToDevice, st));
  MatView M{dm.as<double>(), ld, (long long)n * ld};
  SPW_TRY(cov_assemble_batched(type, n, dc.as<double>(), dc.as<double>() + n, dp.as<double>(), dp.as<double>() + 1,
                               nullptr, 0, M, 1, 0, st));
  SPW_CUDA(cudaMemcpy2DAsync(out, sizeof(double) * n, dm.p, sizeof(double) * ld, sizeof(double) * n, n,
                             cudaMemcpyDeviceToHost, st));
  SPW_CUDA(cudaStreamSynchronize(st));
  return SPW_OK;
}

int spw_cov_delta(int type, long long len, const double* delta, double phi, double sig2, double* out) {
  if (type < 0 || type > 2 || len < 0 || (len && (!delta || !out))) { set_error("spw_cov_delta: bad arguments"); return SPW_ERR_ARG; }
  if (len == 0) return SPW_OK;
  DeviceCtx* ctx;
  SPW_TRY(get_ctx(&ctx));
  cudaStream_t st = ctx->stream;
  DevBuf din, dout;
  SPW_TRY(din.alloc(sizeof(double) * len));
  SPW_TRY(dout.alloc(sizeof(double) * len));
  SPW_CUDA(cudaMemcpyAsync(din.p, delta, sizeof(double) * len, cudaMemcpyHostToDevice, st));
  SPW_TRY(cov_delta(type, len, din.as<double>(), phi, sig2, dout.as<double>(), st));
  SPW_CUDA(cudaMemcpyAsync(out, dout.p, sizeof(double) * len, cudaMemcpyDeviceToHost, st));
  SPW_CUDA(cudaStreamSynchronize(st));
  return SPW_OK;
}

int spw_cross_cov(int type, int n, const double* coords, int g, const double* grid, double phi, double* out) {
  if (type < 0 || type > 2 || n < 1 || g < 1 || !coords || !grid || !out) { set_error("spw_cross_cov: bad arguments"); return SPW_ERR_ARG; }
  DeviceCtx* ctx;
  SPW_TRY(get_ctx(&ctx));
  cudaStream_t st = ctx->stream;
  const int c = ncomp_of(type);
  const long long ld = pad_ld(n);
  DevBuf dc, dg, dp, da;
  SPW_TRY(dc.alloc(sizeof(double) * 2 * n));
  SPW_TRY(dg.alloc(sizeof(double) * 2 * g));
  SPW_TRY(dp.alloc(sizeof(double)));
  SPW_CUDA(cudaMemcpyAsync(dc.p, coords, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, st));
  SPW_CUDA(cudaMemcpyAsync(dg.p, grid, sizeof(double) * 2 * g, cudaMemcpyHostToDevice, st));
  SPW_CUDA(cudaMemcpyAsync(dp.p, &phi, sizeof(double), cudaMemcpyHostToDevice, st));
  const int gc = std::min(g, 8192);
  SPW_TRY(da.alloc(sizeof(double) * (size_t)gc * c * ld));
  MatView AT{da.as<double>(), ld, (long long)gc * c * ld};
  for (int g0 = 0; g0 < g; g0 += gc) {
    const int ng = std::min(gc, g - g0);
    SPW_TRY(cross_cov_batched(type, n, dc.as<double>(), dc.as<double>() + n, g0, ng, dg.as<double>(),
                              dg.as<double>() + g, dp.as<double>(), AT, 1, st));
    SPW_CUDA(cudaMemcpy2DAsync(out + (size_t)g0 * c * n, sizeof(double) * n, da.p, sizeof(double) * ld,
                               sizeof(double) * n, (size_t)ng * c, cudaMemcpyDeviceToHost, st));
    SPW_CUDA(cudaStreamSynchronize(st));
  }
  return SPW_OK;
}

int spw_chol(int n, const double* a, double* u, double* logdet_half, int* info) {
  if (n < 1 || !a || !u) { set_error("spw_chol: bad arguments"); return SPW_ERR_ARG; }
  DeviceCtx* ctx;
  SPW_TRY(get_ctx(&ctx));
  cudaStream_t st = ctx->stream;
  const LinalgPlan* plan;
  SPW_TRY(get_plan(*ctx, n, 0, 1, &plan));
  const long long ld = pad_ld(n);
  DevBuf dm, dd, ds, dl;
  SPW_TRY(dm.alloc(sizeof(double) * n * ld));
  SPW_TRY(dd.alloc(sizeof(double) * plan->nblk * NB * NB));
  SPW_TRY(ds.alloc(sizeof(int)));
  SPW_TRY(dl.alloc(sizeof(double)));
  SPW_CUDA(cudaMemsetAsync(ds.p, 0, sizeof(int), st));
  // R's column-major upper triangle a[i + n*j] (i <= j) is this library's row-major lower triangle L[j][i]
  SPW_CUDA(cudaMemcpy2DAsync(dm.p, sizeof(double) * ld, a, sizeof(double) * n, sizeof(double) * n, n,
                             cudaMemcpyHostToDevice, st));
  MatView L{dm.as<double>(), ld, (long long)n * ld};
  const DagPlan* dag;
  SPW_TRY(get_dag_plan(*ctx, n, 0, &dag));
  const LaPlan* lap;
  LaStreams* las;
  SPW_TRY(get_la_plan(*ctx, n, 0, &lap, &las));
  if (dag) {
    SPW_TRY(potrf_dag(*dag, L, dd.as<double>(), ds.as<int>(), st));
  } else if (lap) {
    SPW_TRY(potrf_lookahead(*lap, *las, L, dd.as<double>(), ds.as<int>(), st));
  } else {
    Lookahead* la;
    SPW_TRY(get_lookahead(*ctx, n, &la));
    SPW_TRY(potrf_batched(*plan, L, dd.as<double>(), nullptr, 1, ds.as<int>(), st, la));
  }
  SPW_TRY(logdet_quad_batched(n, L, dl.as<double>(), nullptr, 1, st));
  std::vector<double> tmp((size_t)n * n);
  int status = 0;
  double ld_half = 0.0;
  SPW_CUDA(cudaMemcpy2DAsync(tmp.data(), sizeof(double) * n, dm.p, sizeof(double) * ld, sizeof(double) * n, n,
                             cudaMemcpyDeviceToHost, st));
  SPW_CUDA(cudaMemcpyAsync(&status, ds.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  SPW_CUDA(cudaMemcpyAsync(&ld_half, dl.p, sizeof(double), cudaMemcpyDeviceToHost, st));
  SPW_CUDA(cudaStreamSynchronize(st));
  if (info) *info = status;
  if (status == LA_STALLED) {
    set_error("chol: the kernels of the look-ahead schedule did not run at the same time (a profiler that replays kernels "
              "one by one, or a time-sliced GPU); set SPW_LA=0");
    return SPW_ERR_CUDA;
  }
  if (status) {
    set_error("the leading minor of order %d is not positive", status);
    return SPW_ERR_NOT_PD;
  }
  for (int j = 0; j < n; ++j)        // column j of U: rows i <= j hold L[j][i]; zeros below the diagonal
    for (int i = 0; i < n; ++i) u[i + (size_t)n * j] = (i <= j) ? tmp[(size_t)j * n + i] : 0.0;
  if (logdet_half) *logdet_half = ld_half;
  return SPW_OK;
}

int spw_hlm_run(int n, const double* coords, const double* y, int type, int niter, int nburn, int report,
                double a_sigma, double b_sigma, double a_tau, double b_tau, double lower_phi, double upper_phi,
                const double* norm, const double* unif, int want_trgt, double* out_sig2, double* out_tau2,
                double* out_phi, double* out_trgt, double* out_accept, double* out_tuning, int* info) {
  (void)nburn;   // the R host slices (nburn+1):niter; the full chains are returned
  if (type < 0 || type > 2 || n < 1 || niter < 0 || report < 1 || !coords || !y || !norm || !unif || !out_sig2 ||
      !out_tau2 || !out_phi || (want_trgt && !out_trgt)) {
    set_error("spw_hlm_run: bad arguments");
    return SPW_ERR_ARG;
  }
  DeviceCtx* ctx;
  SPW_TRY(get_ctx(&ctx));
  cudaStream_t st = ctx->stream;
  const LinalgPlan* plan;
  SPW_TRY(get_plan(*ctx, n, 1, 1, &plan));
  const long long ld = pad_ld(n);
  const int nrep = niter / report;
  DevBuf dc, dy, dnorm, dunif, dL, dd, ds, dstate, dout;
  SPW_TRY(dc.alloc(sizeof(double) * 2 * n));
  SPW_TRY(dy.alloc(sizeof(double) * n));
  SPW_TRY(dnorm.alloc(sizeof(double) * 3 * (size_t)std::max(niter, 1)));
  SPW_TRY(dunif.alloc(sizeof(double) * 3 * (size_t)std::max(niter, 1)));
  SPW_TRY(dL.alloc(sizeof(double) * (size_t)(n + 1) * ld));
  SPW_TRY(dd.alloc(sizeof(double) * plan->nblk * NB * NB));
  SPW_TRY(ds.alloc(sizeof(int)));
  SPW_TRY(dstate.alloc(hlm_state_bytes()));
  const size_t n_out = 3 * (size_t)niter + (size_t)niter + 1 + 6 * (size_t)std::max(nrep, 1);
  SPW_TRY(dout.alloc(sizeof(double) * n_out));
  SPW_CUDA(cudaMemsetAsync(dout.p, 0, sizeof(double) * n_out, st));
  double* o_sig2 = dout.as<double>();
  double* o_tau2 = o_sig2 + niter;
  double* o_phi = o_tau2 + niter;
  double* o_trgt = o_phi + niter;
  double* o_acc = o_trgt + niter + 1;
  double* o_tun = o_acc + 3 * (size_t)std::max(nrep, 1);
  SPW_CUDA(cudaMemcpyAsync(dc.p, coords, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, st));
  SPW_CUDA(cudaMemcpyAsync(dy.p, y, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  if (niter > 0) {
    SPW_CUDA(cudaMemcpyAsync(dnorm.p, norm, sizeof(double) * 3 * (size_t)niter, cudaMemcpyHostToDevice, st));
    SPW_CUDA(cudaMemcpyAsync(dunif.p, unif, sizeof(double) * 3 * (size_t)niter, cudaMemcpyHostToDevice, st));
  }
  HlmWorkspace ws;
  ws.L = dL.as<double>();
  ws.ld = ld;
  ws.dinv = dd.as<double>();
  ws.status = ds.as<int>();
  ws.state = (HlmState*)dstate.p;
  ws.plan = plan;
  SPW_TRY(get_lookahead(*ctx, n, &ws.la));
  SPW_TRY(get_dag_plan(*ctx, n, 1, &ws.dag));
  SPW_TRY(get_la_plan(*ctx, n, 1, &ws.lap, &ws.las));
  {
    const char* env = getenv("SPW_HLM_GRAPH_CACHE");     // 0: instantiate the iteration graph anew in every call
    ws.exec_cache = (env && env[0] == '0') ? nullptr : &ctx->hlm_exec;
  }
  int rc = hlm_run_dev(ws, type, n, dc.as<double>(), dc.as<double>() + n, dy.as<double>(), niter, report, a_sigma,
                       b_sigma, a_tau, b_tau, lower_phi, upper_phi, dnorm.as<double>(), dunif.as<double>(), want_trgt,
                       o_sig2, o_tau2, o_phi, o_trgt, o_acc, o_tun, info, st);
  if (rc != SPW_OK && rc != SPW_ERR_NOT_PD) return rc;
  if (niter > 0) {
    SPW_CUDA(cudaMemcpyAsync(out_sig2, o_sig2, sizeof(double) * niter, cudaMemcpyDeviceToHost, st));
    SPW_CUDA(cudaMemcpyAsync(out_tau2, o_tau2, sizeof(double) * niter, cudaMemcpyDeviceToHost, st));
    SPW_CUDA(cudaMemcpyAsync(out_phi, o_phi, sizeof(double) * niter, cudaMemcpyDeviceToHost, st));
  }
  if (want_trgt) SPW_CUDA(cudaMemcpyAsync(out_trgt, o_trgt, sizeof(double) * ((size_t)niter + 1), cudaMemcpyDeviceToHost, st));
  if (nrep > 0 && out_accept) SPW_CUDA(cudaMemcpyAsync(out_accept, o_acc, sizeof(double) * 3 * nrep, cudaMemcpyDeviceToHost, st));
  if (nrep > 0 && out_tuning) SPW_CUDA(cudaMemcpyAsync(out_tuning, o_tun, sizeof(double) * 3 * nrep, cudaMemcpyDeviceToHost, st));
  SPW_CUDA(cudaStreamSynchronize(st));
  return rc;
}

int spw_gradient_shard_dev(int type, int n, const double* coords_dev, int g, const double* grid_dev, int S,
                           const double* phi_dev, const double* sig2_dev, const double* z_dev, int ldz,
                           const double* eps_dev, double* out_draws_dev, double* out_mean_dev, double* out_var_dev,
                           int* info, void* stream) {
  DeviceCtx* ctx;
  SPW_TRY(get_ctx(&ctx));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  return gradient_shard(type, n, coords_dev, coords_dev + n, g, grid_dev, grid_dev + g, S, phi_dev, sig2_dev, z_dev,
                        ldz, eps_dev, out_draws_dev, out_mean_dev, out_var_dev, info, st);
}

// one device's share of spw_gradient_run: samples [s_lo, s_hi).  The shard's draws stay on the device
// (layout [Sl][c][g], in the context's grow-only staging buffer or in place in the gather buffer) for the gather;
// moments go straight back to the host.
static int gradient_run_one(int device, int type, int n, const double* coords, int g, const double* grid, int S,
                            int s_lo, int s_hi, const double* phi, const double* sig2, const double* z,
                            const double* eps, bool want_draws, double* draws_inplace, double** shard_draws,
                            double* out_mean, double* out_var, int* info, double* t_setup_ms) {
  const double t0 = wall_ms();
  SPW_CUDA(cudaSetDevice(device));
  DeviceCtx* ctx;
  SPW_TRY(get_ctx(&ctx));
  cudaStream_t st = ctx->stream;
  const int c = ncomp_of(type), Sl = s_hi - s_lo;
  *shard_draws = nullptr;
  if (Sl <= 0) return SPW_OK;
  const long long ldz = pad_ld(n);
  const size_t n_units = (size_t)Sl * g * c;
  // one grow-only staging allocation per device instead of ~10 cudaMalloc / cudaFree pairs per call
  size_t need = 4096 + sizeof(double) * (2 * (size_t)n + 2 * (size_t)g + 2 * (size_t)Sl + (size_t)Sl * n + (size_t)Sl * ldz) + 8 * 256;
  if (want_draws) need += sizeof(double) * n_units * (draws_inplace ? 1 : 2) + 2 * 256;
  if (out_mean) need += sizeof(double) * n_units + 256;
  if (out_var) need += sizeof(double) * n_units * c + 256;
  SPW_TRY(grow_buffer(&ctx->io, &ctx->io_bytes, need));
  Carver cv(ctx->io);
  double* dc = cv.take<double>(2 * (size_t)n);
  double* dg = cv.take<double>(2 * (size_t)g);
  double* dphi = cv.take<double>(Sl);
  double* dsig = cv.take<double>(Sl);
  double* dzr = cv.take<double>((size_t)Sl * n);
  double* dz = cv.take<double>((size_t)Sl * ldz);
  double* deps = want_draws ? cv.take<double>(n_units) : nullptr;
  double* ddraw = nullptr;
  if (want_draws) ddraw = draws_inplace ? draws_inplace : cv.take<double>(n_units);
  double* dmean = out_mean ? cv.take<double>(n_units) : nullptr;
  double* dvar = out_var ? cv.take<double>(n_units * c) : nullptr;
  SPW_CUDA(cudaMemcpyAsync(dc, coords, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, st));
  SPW_CUDA(cudaMemcpyAsync(dg, grid, sizeof(double) * 2 * g, cudaMemcpyHostToDevice, st));
  SPW_CUDA(cudaMemcpyAsync(dphi, phi + s_lo, sizeof(double) * Sl, cudaMemcpyHostToDevice, st));
  SPW_CUDA(cudaMemcpyAsync(dsig, sig2 + s_lo, sizeof(double) * Sl, cudaMemcpyHostToDevice, st));
  // model$z is S x n column-major: bring rows s_lo..s_hi of every column, then transpose to [Sl][ldz]
  SPW_CUDA(cudaMemcpy2DAsync(dzr, sizeof(double) * Sl, z + s_lo, sizeof(double) * S, sizeof(double) * Sl, n,
                             cudaMemcpyHostToDevice, st));
  SPW_TRY(transpose_batched(n, Sl, dzr, Sl, 0, dz, ldz, 0, 1, st));
  if (want_draws)
    SPW_CUDA(cudaMemcpyAsync(deps, eps + (size_t)s_lo * g * c, sizeof(double) * n_units, cudaMemcpyHostToDevice, st));
  if (t_setup_ms) *t_setup_ms = wall_ms() - t0;
  int rc = gradient_shard(type, n, dc, dc + n, g, dg, dg + g, Sl, dphi, dsig, dz, (int)ldz, deps, ddraw, dmean, dvar,
                          info, st);
  if (rc != SPW_OK) {
    if (info && info[0] > 0) info[0] += s_lo;
    return rc;
  }
  if (out_mean)
    SPW_CUDA(cudaMemcpyAsync(out_mean + (size_t)s_lo * g * c, dmean, sizeof(double) * n_units, cudaMemcpyDeviceToHost, st));
  if (out_var)
    SPW_CUDA(cudaMemcpyAsync(out_var + (size_t)s_lo * g * c * c, dvar, sizeof(double) * n_units * c,
                             cudaMemcpyDeviceToHost, st));
  SPW_CUDA(cudaStreamSynchronize(st));
  *shard_draws = ddraw;
  return SPW_OK;
}

// restores the caller's current device on every exit path
struct DeviceRestore {
  int dev = -1;
  DeviceRestore() { if (cudaGetDevice(&dev) != cudaSuccess) dev = -1; }
  ~DeviceRestore() { if (dev >= 0) cudaSetDevice(dev); }
};

int spw_gradient_run(int type, int n, const double* coords, int g, const double* grid, int S, const double* phi,
                     const double* sig2, const double* z, const double* eps, int ngpu, int nbatch, double prob,
                     double* out_draws, double* out_summary, double* out_mean, double* out_var, int* info) {
  if (type < 0 || type > 2 || n < 1 || g < 1 || S < 0 || !coords || !grid || (S && (!phi || !sig2 || !z))) {
    set_error("spw_gradient_run: bad arguments");
    return SPW_ERR_ARG;
  }
  const bool want_draws = out_draws || out_summary;
  if (want_draws && !eps) { set_error("spw_gradient_run: eps is required when draws are requested"); return SPW_ERR_ARG; }
  if (out_summary && nbatch < 1) { set_error("spw_gradient_run: nbatch must be >= 1 for the summary"); return SPW_ERR_ARG; }
  if (info) info[0] = info[1] = info[2] = info[3] = 0;
  if (S == 0) return SPW_OK;
  int ndev = spw_device_count();
  if (ndev < 1) { set_error("spw_gradient_run: no CUDA device"); return SPW_ERR_CUDA; }
  if (ngpu < 1) ngpu = 1;
  if (ngpu > ndev) { set_error("spw_gradient_run: ngpu = %d but only %d devices", ngpu, ndev); return SPW_ERR_ARG; }
  if (ngpu > S) ngpu = S;
  static const bool timing = getenv("SPW_TIMING") && getenv("SPW_TIMING")[0] == '1';
  const double t_begin = wall_ms();
  DeviceRestore restore;       // the caller's device is current again when this function returns
  int root = restore.dev < 0 ? 0 : restore.dev;
  if (ngpu > 1) root = 0;      // shards run on devices 0 .. ngpu-1, the gather goes to device 0
  const int c = ncomp_of(type);
  const size_t n_all = (size_t)S * g * c;
  // gather buffer on the root device: all draws, sample-major [S][c][g] (+ the R-layout copy and the summary)
  SPW_CUDA(cudaSetDevice(root));
  DeviceCtx* rctx;
  SPW_TRY(get_ctx(&rctx));
  cudaStream_t rst = rctx->stream;
  double *dall = nullptr, *dr = nullptr, *dsum = nullptr;
  if (want_draws) {
    size_t need = sizeof(double) * n_all + 4096;
    if (out_draws) need += sizeof(double) * n_all + 256;
    if (out_summary) need += sizeof(double) * 3 * (size_t)g * c + 256;
    SPW_TRY(grow_buffer(&rctx->io2, &rctx->io2_bytes, need));
    Carver cv(rctx->io2);
    dall = cv.take<double>(n_all);
    if (out_draws) dr = cv.take<double>(n_all);
    if (out_summary) dsum = cv.take<double>(3 * (size_t)g * c);
  }
  auto lo_of = [&](int d) { return (int)((long long)S * d / ngpu); };
  // the NCCL clique is created while the shards compute (communicator set-up takes seconds on 8 GPUs)
  std::thread nccl_init;
  if (ngpu > 1 && want_draws) nccl_init = std::thread([ngpu]() { nccl_prepare(ngpu); });
  // contiguous sample shards, one host thread per device (R/spatial_gradient.R:29-32,200)
  std::vector<int> rcs(ngpu, SPW_OK);
  std::vector<std::vector<int>> infos(ngpu, std::vector<int>(4, 0));
  std::vector<std::string> errs(ngpu);
  std::vector<double*> shard(ngpu, nullptr);
  std::vector<double> t_setup(ngpu, 0.0);
  auto work = [&](int d) {
    const int dev = (ngpu > 1) ? d : root;
    double* inplace = (want_draws && dev == root) ? dall + (size_t)lo_of(d) * g * c : nullptr;
    rcs[d] = gradient_run_one(dev, type, n, coords, g, grid, S, lo_of(d), lo_of(d + 1), phi, sig2, z, eps, want_draws,
                              inplace, &shard[d], out_mean, out_var, infos[d].data(), &t_setup[d]);
    if (rcs[d] != SPW_OK) errs[d] = tl_error;
  };
  if (ngpu == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int d = 0; d < ngpu; ++d) th.emplace_back(work, d);
    for (auto& t : th) t.join();
  }
  if (nccl_init.joinable()) nccl_init.join();
  const double t_compute = wall_ms();
  SPW_CUDA(cudaSetDevice(root));
  for (int d = 0; d < ngpu; ++d)
    if (rcs[d] != SPW_OK) {
      if (info) memcpy(info, infos[d].data(), sizeof(int) * 4);
      set_error("%s", errs[d].c_str());
      return rcs[d];
    }
  if (!want_draws) return SPW_OK;
  // the one gather (R/spatial_gradient.R:378-390 rbind): a single grouped NCCL send/recv over NVLink when NCCL can
  // be loaded, otherwise peer copies (device to device over NVLink once peer access is enabled)
  bool gathered = (ngpu == 1);
  if (!gathered) {
    std::vector<const double*> src(ngpu, nullptr);
    std::vector<size_t> cnt(ngpu, 0);
    std::vector<double*> dst(ngpu, nullptr);
    for (int d = 1; d < ngpu; ++d) {
      src[d] = shard[d];
      cnt[d] = (size_t)(lo_of(d + 1) - lo_of(d)) * g * c;
      dst[d] = dall + (size_t)lo_of(d) * g * c;
    }
    const int nrc = nccl_gather(ngpu, src, cnt, dst);
    if (nrc == SPW_OK) gathered = true;
    else if (nrc > 0) return nrc;
    g_last_gather_nccl = gathered ? 1 : 0;
  }
  for (int d = 0; d < ngpu && !gathered; ++d) {
    const int dev = (ngpu > 1) ? d : root;
    if (dev == root) continue;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, root, dev) == cudaSuccess && can) {
      cudaError_t pe = cudaDeviceEnablePeerAccess(dev, 0);
      if (pe != cudaSuccess) cudaGetLastError();   // already enabled (or unsupported): the copy still works
    }
    const size_t cnt = (size_t)(lo_of(d + 1) - lo_of(d)) * g * c;
    SPW_CUDA(cudaMemcpyPeerAsync(dall + (size_t)lo_of(d) * g * c, root, shard[d], dev, sizeof(double) * cnt, rst));
  }
  SPW_CUDA(cudaStreamSynchronize(rst));
  const double t_gather = wall_ms();
  if (out_summary) {
    void* scratch;
    SPW_TRY(get_scratch(*rctx, hpd_scratch_bytes(S, g, c, nbatch), &scratch));
    SPW_TRY(hpd_summary_dev(S, g, c, nbatch, prob, dall, dsum, scratch, rst));
    SPW_CUDA(cudaMemcpyAsync(out_summary, dsum, sizeof(double) * 3 * (size_t)g * c, cudaMemcpyDeviceToHost, rst));
  }
  double t_summary = t_gather;
  if (timing) {
    SPW_CUDA(cudaStreamSynchronize(rst));
    t_summary = wall_ms();
  }
  if (out_draws) {
    // [S][c][g] -> the c matrices S x g, column-major: out_draws[s + S*(j + g*a)]; transposed and copied back in
    // column chunks so that the copy of one chunk overlaps the transpose of the next
    const size_t cols = (size_t)c * g;
    const size_t chunk = std::max<size_t>(1, std::min<size_t>(cols, ((size_t)64 << 20) / (sizeof(double) * (size_t)S)));
    for (size_t c0 = 0; c0 < cols; c0 += chunk) {
      const size_t nc = std::min(chunk, cols - c0);
      SPW_TRY(transpose_batched(S, (int)nc, dall + c0, (long long)cols, 0, dr + c0 * S, S, 0, 1, rst));
      SPW_CUDA(cudaMemcpyAsync(out_draws + c0 * S, dr + c0 * S, sizeof(double) * nc * S, cudaMemcpyDeviceToHost, rst));
    }
  }
  SPW_CUDA(cudaStreamSynchronize(rst));
  if (timing) {
    const double t_end = wall_ms();
    double ts = 0.0;
    for (double v : t_setup) ts = std::max(ts, v);
    fprintf(stderr, "[spw] spw_gradient_run S=%d g=%d n=%d ngpu=%d: total %.1f ms = shards %.1f (of which set-up + H2D <= %.1f) + "
            "gather %.1f (%s) + summary %.1f + draws to R layout and D2H %.1f\n", S, g, n, ngpu, t_end - t_begin,
            t_compute - t_begin, ts, t_gather - t_compute, ngpu == 1 ? "none" : (g_last_gather_nccl == 1 ? "nccl" : "peer copies"),
            t_summary - t_gather, t_end - t_summary);
  }
  return SPW_OK;
}

// bayes_cwomb(type = "rectilinear"): host buffers in, the two wombling measures per (segment, sample) out.
int spw_wombling_rect_run(int type, int n, const double* coords, int nseg, const double* curve, int rule_len, int S,
                          const double* phi, const double* sig2, const double* z, const double* eps, double* out_w,
                          double* out_mean, double* out_var, int* info) {
  if ((type != SPW_COV_GAUSSIAN && type != SPW_COV_MATERN2) || n < 1 || nseg < 1 || rule_len < 2 || S < 0 || !coords ||
      !curve || (S && (!phi || !sig2 || !z || !eps || !out_w))) {
    set_error("spw_wombling_rect_run: bad arguments (cov type must be gaussian or matern2: the matern1 kernels of the "
              "reference divide 0 by 0)");
    return SPW_ERR_ARG;
  }
  if (info) info[0] = info[1] = info[2] = info[3] = 0;
  if (S == 0) return SPW_OK;
  DeviceCtx* ctx;
  SPW_TRY(get_ctx(&ctx));
  cudaStream_t st = ctx->stream;
  const long long ld = pad_ld(n);
  const int nblk = ceil_div(n, NB);
  const size_t per_sample = 2 * (size_t)n * ld * sizeof(double) + (size_t)ld * sizeof(double) +
                            (size_t)nblk * NB * NB * sizeof(double) + (size_t)nseg * 2 * ld * sizeof(double) + 64;
  const size_t budget = workspace_budget();
  int B = (int)std::min<size_t>((size_t)std::min(S, 1024), budget / per_sample);
  if (B < 1) { set_error("spw_wombling_rect_run: workspace too small (n = %d)", n); return SPW_ERR_ARG; }
  const LinalgPlan* plan;
  SPW_TRY(get_plan(*ctx, n, 0, B < 4 ? 1 : 0, &plan));
  void* wsp;
  SPW_TRY(get_workspace(*ctx, per_sample * B + 4096, &wsp));
  Carver cv(wsp);
  double* Lb = cv.take<double>((size_t)B * n * ld);
  double* Li = cv.take<double>((size_t)B * n * ld);
  double* At = cv.take<double>((size_t)B * nseg * 2 * ld);
  double* vb = cv.take<double>((size_t)B * ld);
  double* dinv = cv.take<double>((size_t)B * nblk * NB * NB);
  // inputs / outputs of the whole call in the context's staging buffer
  const size_t n_units = (size_t)S * nseg;
  size_t need = 4096 + sizeof(double) * (2 * (size_t)n + 2 * ((size_t)nseg + 1) + 2 * (size_t)S + (size_t)S * n + (size_t)S * ld +
                                         n_units * (2 + 5 + 2 + 2 + 4)) + sizeof(int) * 3 * (size_t)S + 16 * 256;
  SPW_TRY(grow_buffer(&ctx->io, &ctx->io_bytes, need));
  Carver io(ctx->io);
  double* dc = io.take<double>(2 * (size_t)n);
  double* dcurve = io.take<double>(2 * ((size_t)nseg + 1));
  double* dphi = io.take<double>(S);
  double* dsig = io.take<double>(S);
  double* dzr = io.take<double>((size_t)S * n);
  double* dz = io.take<double>((size_t)S * ld);
  double* deps = io.take<double>(n_units * 2);
  double* dgram = io.take<double>(n_units * 5);
  double* dw = io.take<double>(n_units * 2);
  double* dmean = io.take<double>(n_units * 2);
  double* dvar = io.take<double>(n_units * 4);
  int* status = io.take<int>(3 * (size_t)S);            // chol(K) | the fused kernel's own (unused) 2 x 2 | mvrnorm
  SPW_CUDA(cudaMemcpyAsync(dc, coords, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, st));
  SPW_CUDA(cudaMemcpyAsync(dcurve, curve, sizeof(double) * 2 * (nseg + 1), cudaMemcpyHostToDevice, st));
  SPW_CUDA(cudaMemcpyAsync(dphi, phi, sizeof(double) * S, cudaMemcpyHostToDevice, st));
  SPW_CUDA(cudaMemcpyAsync(dsig, sig2, sizeof(double) * S, cudaMemcpyHostToDevice, st));
  SPW_CUDA(cudaMemcpyAsync(dzr, z, sizeof(double) * (size_t)S * n, cudaMemcpyHostToDevice, st));      // S x n column-major
  SPW_TRY(transpose_batched(n, S, dzr, S, 0, dz, ld, 0, 1, st));
  SPW_CUDA(cudaMemcpyAsync(deps, eps, sizeof(double) * n_units * 2, cudaMemcpyHostToDevice, st));
  SPW_CUDA(cudaMemsetAsync(status, 0, sizeof(int) * 3 * (size_t)S, st));
  MatView L{Lb, ld, (long long)n * ld}, Linv{Li, ld, (long long)n * ld}, AT{At, ld, (long long)nseg * 2 * ld};
  for (int s0 = 0; s0 < S; s0 += B) {
    const int nb = std::min(B, S - s0);
    // K = cov(phi, sig2 = 1): the reference's sig2 * K enters as 1 / sig2 factors in womb_finish_kernel
    SPW_TRY(cov_assemble_batched(type, n, dc, dc + n, dphi + s0, nullptr, nullptr, 1, L, nb, 1, st));
    SPW_TRY(potrf_batched(*plan, L, dinv, &Linv, nb, status + s0, st));
    SPW_TRY(trtri_batched(*plan, L, Linv, nb, st));
    SPW_TRY(trmv_lower_batched(n, Linv, dz + (long long)s0 * ld, ld, vb, ld, nb, st));
    SPW_TRY(womb_gamma_batched(type, n, dc, dc + n, dcurve, nseg, rule_len, dphi + s0, AT, nb, st));
    GradParams P;
    P.linv = Linv;
    P.at = AT;
    P.v = vb;
    P.ldv = ld;
    P.phi = dphi + s0;
    P.sig2 = dsig + s0;
    P.eps = nullptr;
    P.out_draws = nullptr;
    P.out_mean = nullptr;
    P.out_var = nullptr;
    P.out_gram = dgram;
    P.status = status + S + s0;
    P.n = n;
    P.g_begin = 0;
    P.g_count = nseg;
    P.g_total = nseg;
    P.s_begin = s0;
    // two components per segment: the fused kernel in its two-component instantiation returns M = G K^-1 G' and G K^-1 z
    SPW_TRY(gradient_tma_launch(SPW_COV_MATERN1, P, nb, st));
    SPW_TRY(womb_finish_batched(type, nb, s0, S, nseg, rule_len, dcurve, dphi + s0, dsig + s0, dgram, deps, dw, dmean, dvar,
                                status + 2 * (size_t)S + s0, st));
  }
  std::vector<int> hs(3 * (size_t)S);
  SPW_CUDA(cudaMemcpyAsync(hs.data(), status, sizeof(int) * 3 * (size_t)S, cudaMemcpyDeviceToHost, st));
  SPW_CUDA(cudaMemcpyAsync(out_w, dw, sizeof(double) * n_units * 2, cudaMemcpyDeviceToHost, st));
  if (out_mean) SPW_CUDA(cudaMemcpyAsync(out_mean, dmean, sizeof(double) * n_units * 2, cudaMemcpyDeviceToHost, st));
  if (out_var) SPW_CUDA(cudaMemcpyAsync(out_var, dvar, sizeof(double) * n_units * 4, cudaMemcpyDeviceToHost, st));
  SPW_CUDA(cudaStreamSynchronize(st));
  for (int s = 0; s < S; ++s) {
    if (hs[s]) {
      if (info) { info[0] = s + 1; info[1] = hs[s]; info[2] = 1; }
      set_error("bayes_cwomb: sample %d: the leading minor of order %d of Sig.Z is not positive", s + 1, hs[s]);
      return SPW_ERR_NOT_PD;
    }
    if (hs[2 * (size_t)S + s]) {
      if (info) { info[0] = s + 1; info[1] = hs[2 * (size_t)S + s]; info[2] = 3; }
      set_error("bayes_cwomb: sample %d, segment %d: 'Sigma' is not positive definite (mvrnorm)", s + 1, hs[2 * (size_t)S + s]);
      return SPW_ERR_NOT_PD;
    }
  }
  return SPW_OK;
}

int spw_spsample_run(int type, int n, const double* coords, int S, const double* phi, const double* sig2,
                     const double* tau2, const double* w, const double* eps, double* out_z, int* bad, int* info) {
  if (type < 0 || type > 2 || n < 1 || S < 0 || !coords || (S && (!phi || !sig2 || !tau2 || !w || !eps || !out_z))) {
    set_error("spw_spsample_run: bad arguments");
    return SPW_ERR_ARG;
  }
  if (info) info[0] = info[1] = info[2] = info[3] = 0;
  if (S == 0) return SPW_OK;
  DeviceCtx* ctx;
  SPW_TRY(get_ctx(&ctx));
  cudaStream_t st = ctx->stream;
  const long long ldv = pad_ld(n);
  std::vector<double> h_par(4 * (size_t)S);
  for (int s = 0; s < S; ++s) {
    h_par[s] = phi[s];
    h_par[S + s] = 1.0 / sig2[s];
    h_par[2 * (size_t)S + s] = 1.0 / tau2[s];
    h_par[3 * (size_t)S + s] = 0.0;
  }
  DevBuf dc, dpar, dwr, dw, der, de, dz, dzr;
  SPW_TRY(dc.alloc(sizeof(double) * 2 * n));
  SPW_TRY(dpar.alloc(sizeof(double) * 4 * S));
  SPW_TRY(dwr.alloc(sizeof(double) * (size_t)S * n));
  SPW_TRY(dw.alloc(sizeof(double) * (size_t)S * ldv));
  SPW_TRY(der.alloc(sizeof(double) * (size_t)S * n));
  SPW_TRY(de.alloc(sizeof(double) * (size_t)S * ldv));
  SPW_TRY(dz.alloc(sizeof(double) * (size_t)S * ldv));
  SPW_TRY(dzr.alloc(sizeof(double) * (size_t)S * n));
  SPW_CUDA(cudaMemcpyAsync(dc.p, coords, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, st));
  SPW_CUDA(cudaMemcpyAsync(dpar.p, h_par.data(), sizeof(double) * 4 * S, cudaMemcpyHostToDevice, st));
  // w and eps are n x S column-major == [S][n] row-major: repack to the padded leading dimension
  SPW_CUDA(cudaMemcpyAsync(dwr.p, w, sizeof(double) * (size_t)S * n, cudaMemcpyHostToDevice, st));
  SPW_CUDA(cudaMemcpyAsync(der.p, eps, sizeof(double) * (size_t)S * n, cudaMemcpyHostToDevice, st));
  SPW_CUDA(cudaMemcpy2DAsync(dw.p, sizeof(double) * ldv, dwr.p, sizeof(double) * n, sizeof(double) * n, S,
                             cudaMemcpyDeviceToDevice, st));
  SPW_CUDA(cudaMemcpy2DAsync(de.p, sizeof(double) * ldv, der.p, sizeof(double) * n, sizeof(double) * n, S,
                             cudaMemcpyDeviceToDevice, st));
  std::vector<int> sk(S), sb(S);
  double* par = dpar.as<double>();
  SPW_TRY(spsample_shard(type, n, dc.as<double>(), dc.as<double>() + n, S, par, par + S, par + 2 * (size_t)S,
                         par + 3 * (size_t)S, dw.as<double>(), de.as<double>(), ldv, dz.as<double>(), sk.data(),
                         sb.data(), st));
  // R/spsample_beta.R:70-71: samples whose chol(R.Z) failed are redone once with R.Z + 1e-4 I
  for (int s = 0; s < S; ++s) {
    if (!sk[s]) continue;
    const double jit = 1e-4;
    SPW_CUDA(cudaMemcpyAsync(par + 3 * (size_t)S + s, &jit, sizeof(double), cudaMemcpyHostToDevice, st));
    int k1 = 0, b1 = 0;
    SPW_TRY(spsample_shard(type, n, dc.as<double>(), dc.as<double>() + n, 1, par + s, par + S + s, par + 2 * (size_t)S + s,
                           par + 3 * (size_t)S + s, dw.as<double>() + (size_t)s * ldv, de.as<double>() + (size_t)s * ldv, ldv,
                           dz.as<double>() + (size_t)s * ldv, &k1, &b1, st));
    if (k1) {   // the reference would stop here: chol() of the jittered matrix failed too
      if (info) { info[0] = s + 1; info[1] = k1; info[2] = 1; }
      set_error("spsample: sample %d: chol(R.Z + 1e-4 I) failed (leading minor %d)", s + 1, k1);
      return SPW_ERR_NOT_PD;
    }
    sb[s] = b1;
  }
  for (int s = 0; s < S; ++s) bad[s] = sb[s];
  // [S][ldv] -> S x n column-major (out_z[s + S*k])
  SPW_TRY(transpose_batched(S, n, dz.as<double>(), ldv, 0, dzr.as<double>(), S, 0, 1, st));
  SPW_CUDA(cudaMemcpyAsync(out_z, dzr.p, sizeof(double) * (size_t)S * n, cudaMemcpyDeviceToHost, st));
  SPW_CUDA(cudaStreamSynchronize(st));
  return SPW_OK;
}

int spw_hpd_summary_dev(int S, int g, int ncomp, int nbatch, double prob, const double* draws_dev, double* out_dev,
                        void* stream) {
  DeviceCtx* ctx;
  SPW_TRY(get_ctx(&ctx));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  void* scratch;
  SPW_TRY(get_scratch(*ctx, hpd_scratch_bytes(S, g, ncomp, nbatch), &scratch));
  return hpd_summary_dev(S, g, ncomp, nbatch, prob, draws_dev, out_dev, scratch, st);
}

int spw_draws_to_r_layout_dev(int S, int g, int ncomp, const double* draws_dev, double* out_dev, void* stream) {
  DeviceCtx* ctx;
  SPW_TRY(get_ctx(&ctx));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  return transpose_batched(S, ncomp * g, draws_dev, (long long)ncomp * g, 0, out_dev, S, 0, 1, st);
}

int spw_hpd_summary(int S, int g, int nbatch, double prob, const double* draws, double* out) {
  if (S < 1 || g < 1 || !draws || !out) { set_error("spw_hpd_summary: bad arguments"); return SPW_ERR_ARG; }
  DeviceCtx* ctx;
  SPW_TRY(get_ctx(&ctx));
  cudaStream_t st = ctx->stream;
  DevBuf din, dt, dout;
  SPW_TRY(din.alloc(sizeof(double) * (size_t)S * g));
  SPW_TRY(dt.alloc(sizeof(double) * (size_t)S * g));
  SPW_TRY(dout.alloc(sizeof(double) * 3 * (size_t)g));
  SPW_CUDA(cudaMemcpyAsync(din.p, draws, sizeof(double) * (size_t)S * g, cudaMemcpyHostToDevice, st));
  // R layout [g][S] (S contiguous) -> [S][1][g]
  SPW_TRY(transpose_batched(g, S, din.as<double>(), S, 0, dt.as<double>(), g, 0, 1, st));
  void* scratch;
  SPW_TRY(get_scratch(*ctx, hpd_scratch_bytes(S, g, 1, nbatch), &scratch));
  SPW_TRY(hpd_summary_dev(S, g, 1, nbatch, prob, dt.as<double>(), dout.as<double>(), scratch, st));
  SPW_CUDA(cudaMemcpyAsync(out, dout.p, sizeof(double) * 3 * (size_t)g, cudaMemcpyDeviceToHost, st));
  SPW_CUDA(cudaStreamSynchronize(st));
  return SPW_OK;
}

int spw_surface_analysis_dev(int S, int g, const double* draws_dev, double* out_dev, void* stream) {
  DeviceCtx* ctx;
  SPW_TRY(get_ctx(&ctx));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  return surface_analysis_dev(S, g, draws_dev, out_dev, st);
}

int spw_surface_analysis(int S, int g, const double* draws, int nbatch, double prob, double* out_summary,
                         double* out_derived) {
  if (S < 1 || g < 1 || !draws || (!out_summary && !out_derived)) { set_error("spw_surface_analysis: bad arguments"); return SPW_ERR_ARG; }
  DeviceCtx* ctx;
  SPW_TRY(get_ctx(&ctx));
  cudaStream_t st = ctx->stream;
  const size_t cnt = (size_t)S * g * 5;
  DevBuf din, dt, dd, dsum;
  SPW_TRY(din.alloc(sizeof(double) * cnt));
  SPW_TRY(dt.alloc(sizeof(double) * cnt));
  SPW_TRY(dd.alloc(sizeof(double) * cnt));
  SPW_CUDA(cudaMemcpyAsync(din.p, draws, sizeof(double) * cnt, cudaMemcpyHostToDevice, st));
  // R layout [5][g][S] (S contiguous) -> [S][5][g]
  SPW_TRY(transpose_batched(5 * g, S, din.as<double>(), S, 0, dt.as<double>(), (long long)5 * g, 0, 1, st));
  SPW_TRY(surface_analysis_dev(S, g, dt.as<double>(), dd.as<double>(), st));
  if (out_summary) {
    SPW_TRY(dsum.alloc(sizeof(double) * 15 * (size_t)g));
    void* scratch;
    SPW_TRY(get_scratch(*ctx, hpd_scratch_bytes(S, g, 5, nbatch), &scratch));
    SPW_TRY(hpd_summary_dev(S, g, 5, nbatch, prob, dd.as<double>(), dsum.as<double>(), scratch, st));
    SPW_CUDA(cudaMemcpyAsync(out_summary, dsum.p, sizeof(double) * 15 * (size_t)g, cudaMemcpyDeviceToHost, st));
  }
  if (out_derived) {
    SPW_TRY(transpose_batched(S, 5 * g, dd.as<double>(), (long long)5 * g, 0, din.as<double>(), S, 0, 1, st));
    SPW_CUDA(cudaMemcpyAsync(out_derived, din.p, sizeof(double) * cnt, cudaMemcpyDeviceToHost, st));
  }
  SPW_CUDA(cudaStreamSynchronize(st));
  return SPW_OK;
}

int spw_cov_p(const int* type, const int* n, const double* coords, const double* phi, const double* sig2, double* out,
              int* status) {
  const int rc = spw_cov(*type, *n, coords, *phi, *sig2, out);
  if (status) *status = rc;
  return rc;
}

int spw_hlm_run_p(const int* n, const double* coords, const double* y, const int* type, const int* niter,
                  const int* nburn, const int* report, const double* hyper, const double* norm, const double* unif,
                  const int* want_trgt, double* out_sig2, double* out_tau2, double* out_phi, double* out_trgt,
                  double* out_accept, double* out_tuning, int* info, int* status) {
  const int rc = spw_hlm_run(*n, coords, y, *type, *niter, *nburn, *report, hyper[0], hyper[1], hyper[2], hyper[3],
                             hyper[4], hyper[5], norm, unif, *want_trgt, out_sig2, out_tau2, out_phi, out_trgt,
                             out_accept, out_tuning, info);
  if (status) *status = rc;
  return rc;
}

int spw_gradient_run_p(const int* type, const int* n, const double* coords, const int* g, const double* grid,
                       const int* S, const double* phi, const double* sig2, const double* z, const double* eps,
                       const int* ngpu, const int* nbatch, const double* prob, const int* want_draws, double* out_draws,
                       double* out_summary, int* info, int* status) {
  const int rc = spw_gradient_run(*type, *n, coords, *g, grid, *S, phi, sig2, z, eps, *ngpu, *nbatch, *prob,
                                  *want_draws ? out_draws : nullptr, out_summary, nullptr, nullptr, info);
  if (status) *status = rc;
  return rc;
}

int spw_hpd_summary_p(const int* S, const int* g, const int* nbatch, const double* prob, const double* draws,
                      double* out, int* status) {
  const int rc = spw_hpd_summary(*S, *g, *nbatch, *prob, draws, out);
  if (status) *status = rc;
  return rc;
}

}  // extern "C"
