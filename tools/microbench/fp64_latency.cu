// Dependent-issue latency and single-warp issue rate of the FP64 operations on the Cholesky pivot chain (one SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I spwombling_b200/csrc -o tools/microbench/fp64_latency tools/microbench/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"

template <int ILP>
__device__ __forceinline__ long long dfma_ilp(double seed, int iters, double* sink) {
  double x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) x[j] = seed + j + threadIdx.x * 1e-9;
  const double y = 0.999999;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], y, 1e-9);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += x[j];
  *sink += s;
  return t1 - t0;
}

__global__ void lat_kernel(double* out, long long* cyc, double seed, int iters) {
  __shared__ __align__(16) double sh[1024];
  double sink = 0;
  cyc[0] = dfma_ilp<1>(seed, iters, &sink);
  cyc[1] = dfma_ilp<2>(seed, iters, &sink);
  cyc[2] = dfma_ilp<4>(seed, iters, &sink);
  cyc[3] = dfma_ilp<8>(seed, iters, &sink);
  cyc[4] = dfma_ilp<16>(seed, iters, &sink);
  double r = seed + 1.0 + threadIdx.x;
  long long t1 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) r = spw::fast_rsqrt(r) + 1.5;
  }
  long long t2 = clock64();
  cyc[5] = t2 - t1;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = seed + i;
  __syncthreads();
  double m = seed;
  long long t5 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) { sh[threadIdx.x] = m; __syncthreads(); m = sh[(threadIdx.x + 1) % blockDim.x] + 1.0; }
  }
  long long t6 = clock64();
  cyc[6] = t6 - t5;
  // 8 independent 128-bit shared loads + dependent add, conflict-free (18-double stride)
  double acc = 0;
  int idx = (threadIdx.x & 31) * 18;
  long long t7 = clock64();
  for (int i = 0; i < iters; ++i) {
    double2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const double2*>(&sh[idx + 2 * u]);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y;
    idx = (idx + (acc > 1e300 ? 18 : 0)) % 900;
  }
  long long t8 = clock64();
  cyc[7] = t8 - t7;
  // back-to-back clock reads
  long long c0 = clock64();
  long long c1 = clock64();
  cyc[8] = c1 - c0;
  out[threadIdx.x] = sink + r + m + acc;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 16 * 8);
  const int iters = 1000;
  for (int threads : {32, 32, 128, 288}) {
    lat_kernel<<<1, threads>>>(out, cyc, 1.25, iters);
    cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, cyc, 128, cudaMemcpyDeviceToHost);
    printf("threads %3d: cycles per DFMA at ILP 1/2/4/8/16: %.2f %.2f %.2f %.2f %.2f | fast_rsqrt+dadd %.1f | sts+bar+lds+dadd %.1f | "
           "8 x lds.128 + 16 dadd %.1f | clock64 pair %lld\n", threads,
           h[0] / (8.0 * iters), h[1] / (16.0 * iters), h[2] / (32.0 * iters), h[3] / (64.0 * iters), h[4] / (128.0 * iters),
           h[5] / (4.0 * iters), h[6] / (4.0 * iters), h[7] / (1.0 * iters), h[8]);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
