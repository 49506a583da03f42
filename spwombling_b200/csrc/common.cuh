// Shared device helpers for libspwombling (sm_100a, fp64).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>

namespace spw {

// ---- error plumbing -------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;   // kernels launched by this library (spw_launch_count)

#define SPW_CUDA(call)                                                                        \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      spw::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return SPW_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define SPW_LAUNCHED()                                                                        \
  do {                                                                                        \
    ++spw::g_launches;                                                                        \
    cudaError_t e__ = cudaGetLastError();                                                     \
    if (e__ != cudaSuccess) {                                                                 \
      spw::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return SPW_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define SPW_TRY(call)                    \
  do {                                   \
    int rc__ = (call);                   \
    if (rc__ != SPW_OK) return rc__;     \
  } while (0)

// ---- FP64 tensor pipe: mma.sync m8n8k4 (SASS DMMA.8x8x4) ---------------------------------------
// fragment ownership (lane = 4*g + q, g = lane>>2, q = lane&3):
//   A (8x4, row):  a  = A[g][q]
//   B (4x8, col):  b  = B[q][g]
//   C (8x8):       c0 = C[g][2q], c1 = C[g][2q+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- flags in global memory between CTAs / kernels that run at the same time -----------------------
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
// generic-proxy writes of other SMs are read here through the async proxy (TMA) and vice versa
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}

// ---- cp.async (LDGSTS), 16-byte, L2-only (.cg) with zero-fill of the bytes past src_bytes -------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 1/sqrt(x) to full double precision on a short dependency chain: hardware approximation (relative error 2^-22)
// followed by ONE cubically convergent (Halley) step, y (1 + e/2 + 3 e^2/8) with e = 1 - x y^2 -- four dependent
// FP64 operations (the library rsqrt() costs ~10).  Non-positive / non-finite x propagate NaN or the IEEE limits
// through the same formula, which the callers' pivot checks catch.
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;\n" : "=d"(y) : "d"(x));
  const double e = fma(-(x * y), y, 1.0);
  const double p = fma(0.375, e, 0.5);
  return fma(y * e, p, y);
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ long long round_up_ll(long long a, long long b) { return (a + b - 1) / b * b; }

// ---- covariance functions of the reference (R/cov_gaussian.R:14, R/cov_matern1.R:14, R/cov_matern2.R:14)
// rho(d) with sig2 = 1; the gaussian kernel squares the (square-rooted) distance as the reference does.
// Every product and sum is an explicitly rounded operation (no FMA contraction whatever the flags of the translation
// unit), so the assembled matrix, the single-CTA hlm chain and spw_cov see bit-identical covariances and
// K[i][j] == K[j][i].
template <int TYPE>
__device__ __forceinline__ double cov_rho(double phi, double d) {
  if constexpr (TYPE == 0) {
    return exp(-__dmul_rn(phi, __dmul_rn(d, d)));
  } else if constexpr (TYPE == 1) {
    const double ad = __dmul_rn(__dmul_rn(phi, 1.7320508075688772), d);   // phi * sqrt(3) * Delta
    return __dmul_rn(__dadd_rn(1.0, ad), exp(-ad));
  } else {
    const double ad = __dmul_rn(__dmul_rn(phi, 2.2360679774997898), d);   // phi * sqrt(5) * Delta
    const double q = __ddiv_rn(__dmul_rn(__dmul_rn(__dmul_rn(phi, phi), 5.0), __dmul_rn(d, d)), 3.0);
    return __dmul_rn(__dadd_rn(__dadd_rn(1.0, ad), q), exp(-ad));
  }
}

}  // namespace spw
