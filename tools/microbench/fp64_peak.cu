// FP64 pipe micro-benchmarks for B200 (sm_100a).
// Measures: DFMA throughput, DMMA (mma.sync m8n8k4 / m16n8k8 / m16n8k16 f64) throughput
// as a function of independent accumulator chains and warps per SM.
// Output: one JSON line per measurement on stdout.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int CH>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = threadIdx.x * 1e-9 + c;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = fma(acc[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += acc[c];
  if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int CH>
__global__ void dmma884_kernel(double* out, int iters, double a, double b) {
  double c0[CH], c1[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) { c0[c] = threadIdx.x * 1e-9; c1[c] = c; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CH; ++c) dmma884(c0[c], c1[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += c0[c] + c1[c];
  if (s == 123.456) out[0] = s;
}

// register-blocked: MF a-fragments x NF b-fragments (as in a real GEMM inner loop), fragments constant
template <int MF, int NF>
__global__ void dmma884_blocked_kernel(double* out, int iters, double a0, double b0) {
  double c0[MF][NF], c1[MF][NF], a[MF], b[NF];
#pragma unroll
  for (int i = 0; i < MF; ++i) a[i] = a0 + i * 1e-3 + threadIdx.x * 1e-6;
#pragma unroll
  for (int j = 0; j < NF; ++j) b[j] = b0 + j * 1e-3;
#pragma unroll
  for (int i = 0; i < MF; ++i)
#pragma unroll
    for (int j = 0; j < NF; ++j) { c0[i][j] = 0; c1[i][j] = 0; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < MF; ++i)
#pragma unroll
      for (int j = 0; j < NF; ++j) dmma884(c0[i][j], c1[i][j], a[i], b[j]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < MF; ++i)
#pragma unroll
    for (int j = 0; j < NF; ++j) s += c0[i][j] + c1[i][j];
  if (s == 123.456) out[0] = s;
}

template <int CH>
__global__ void dmma1688_kernel(double* out, int iters, double av, double bv) {
  double c[CH][4], a[4] = {av, av + 1, av + 2, av + 3}, b[2] = {bv, bv + 1};
#pragma unroll
  for (int k = 0; k < CH; ++k) for (int i = 0; i < 4; ++i) c[k][i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < CH; ++k) dmma1688(c[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < CH; ++k) for (int i = 0; i < 4; ++i) s += c[k][i];
  if (s == 123.456) out[0] = s;
}

template <int CH>
__global__ void dmma16816_kernel(double* out, int iters, double av, double bv) {
  double c[CH][4], a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = av + i;
  for (int i = 0; i < 4; ++i) b[i] = bv + i;
#pragma unroll
  for (int k = 0; k < CH; ++k) for (int i = 0; i < 4; ++i) c[k][i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < CH; ++k) dmma16816(c[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < CH; ++k) for (int i = 0; i < 4; ++i) s += c[k][i];
  if (s == 123.456) out[0] = s;
}

template <typename F>
static float time_launch(F launch) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 2; ++i) launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, 8));
  const int iters = 8192;
  const int warps_list[] = {4, 8, 16, 32};
  for (int wi = 0; wi < 4; ++wi) {
    int warps = warps_list[wi];
    int threads = warps * 32 > 1024 ? 1024 : warps * 32;
    int blocks_per_sm = (warps * 32) / threads;
    int grid = sms * blocks_per_sm;
#define RUN_DFMA(CH) { float ms = time_launch([&]{ dfma_kernel<CH><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }); \
      double fl = 2.0 * CH * iters * (double)grid * threads; \
      printf("{\"bench\": \"dfma\", \"chains\": %d, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", CH, warps, ms, fl / ms * 1e-9); }
    RUN_DFMA(4) RUN_DFMA(8) RUN_DFMA(16)
#define RUN_884(CH) { float ms = time_launch([&]{ dmma884_kernel<CH><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }); \
      double fl = 2.0 * 256 * CH * iters * (double)grid * (threads / 32); \
      printf("{\"bench\": \"dmma884\", \"chains\": %d, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", CH, warps, ms, fl / ms * 1e-9); }
    RUN_884(1) RUN_884(2) RUN_884(4) RUN_884(8) RUN_884(16)
#define RUN_BLK(MF, NF) { float ms = time_launch([&]{ dmma884_blocked_kernel<MF, NF><<<grid, threads>>>(out, iters / 4, 1.0000001, 1e-9); }); \
      double fl = 2.0 * 256 * MF * NF * (iters / 4) * (double)grid * (threads / 32); \
      printf("{\"bench\": \"dmma884_blocked\", \"mf\": %d, \"nf\": %d, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", MF, NF, warps, ms, fl / ms * 1e-9); }
    if (warps <= 16) { RUN_BLK(4, 4) RUN_BLK(4, 5) RUN_BLK(8, 4) RUN_BLK(8, 5) }
#define RUN_1688(CH) { float ms = time_launch([&]{ dmma1688_kernel<CH><<<grid, threads>>>(out, iters / 4, 1.0000001, 1e-9); }); \
      double fl = 2.0 * 16 * 8 * 8 * CH * (iters / 4) * (double)grid * (threads / 32); \
      printf("{\"bench\": \"dmma1688\", \"chains\": %d, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", CH, warps, ms, fl / ms * 1e-9); }
    RUN_1688(2) RUN_1688(4) RUN_1688(8)
#define RUN_16816(CH) { float ms = time_launch([&]{ dmma16816_kernel<CH><<<grid, threads>>>(out, iters / 8, 1.0000001, 1e-9); }); \
      double fl = 2.0 * 16 * 8 * 16 * CH * (iters / 8) * (double)grid * (threads / 32); \
      printf("{\"bench\": \"dmma16816\", \"chains\": %d, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", CH, warps, ms, fl / ms * 1e-9); }
    RUN_16816(2) RUN_16816(4) RUN_16816(8)
    fflush(stdout);
  }
  return 0;
}
