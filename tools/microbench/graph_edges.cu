// What does a two-chain look-ahead schedule cost as a CUDA graph?  Stand-in kernels that spin for a fixed time:
//   D(k): 1 CTA, d_us        (diagonal block + look-ahead row)
//   P(k): ncta CTAs, p_us    (panel product)      U(k): ncta CTAs, u_us   (near-column update)
//   pattern 0: one stream, D P U D P U ...                         (today's critical path)
//   pattern 1: D chain on stream a; P(k) after D(k) and U(k-1); U(k) after P(k); D(k+2) after U(k)   (look-ahead 2)
//   pattern 2: pattern 1 + a low-priority stream with bulk kernels (bulk_cta CTAs of b_us) every 4 steps, joined 8
//              steps later
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench/graph_edges tools/microbench/graph_edges.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

extern __shared__ double dyn_smem[];
__global__ void spin_kernel(unsigned ns, int* sink) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < ns);
  if (ns == 0xffffffffu) *sink = int(dyn_smem[threadIdx.x]);
}

int main(int argc, char** argv) {
  const int steps = 79;
  const double d_us = argc > 1 ? atof(argv[1]) : 18, p_us = argc > 2 ? atof(argv[2]) : 5, u_us = argc > 3 ? atof(argv[3]) : 6;
  const int ncta = argc > 4 ? atoi(argv[4]) : 150;
  const double b_us = argc > 5 ? atof(argv[5]) : 13;
  const int bulk_cta = argc > 6 ? atoi(argv[6]) : 5500;
  const size_t sm3 = 70 * 1024;                                     // three CTAs per SM, like the 32 x 64 tile engine
  CK(cudaFuncSetAttribute(spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm3)));
  int* sink; CK(cudaMalloc(&sink, 4));
  int lo, hi; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  cudaStream_t sa, sb, sc;
  CK(cudaStreamCreateWithPriority(&sa, cudaStreamNonBlocking, hi));
  CK(cudaStreamCreateWithPriority(&sb, cudaStreamNonBlocking, hi + (lo > hi + 1 ? 1 : 0)));
  CK(cudaStreamCreateWithPriority(&sc, cudaStreamNonBlocking, lo));
  auto ns = [](double us) { return unsigned(us * 1000); };
  for (int pattern = 0; pattern < 4; ++pattern) {
    std::vector<cudaEvent_t> evD(steps), evU(steps), evP(steps), evB(steps);
    for (int k = 0; k < steps; ++k) {
      CK(cudaEventCreateWithFlags(&evD[k], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&evU[k], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&evP[k], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&evB[k], cudaEventDisableTiming));
    }
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(sa, cudaStreamCaptureModeGlobal));
    if (pattern == 0) {
      for (int k = 0; k < steps; ++k) {
        spin_kernel<<<1, 384, 0, sa>>>(ns(d_us - 5), sink);       // today's diagonal kernel has no look-ahead work
        spin_kernel<<<ncta, 256, sm3, sa>>>(ns(p_us), sink);
        spin_kernel<<<ncta, 256, sm3, sa>>>(ns(u_us), sink);
      }
    } else {
      cudaEvent_t fork; CK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
      CK(cudaEventRecord(fork, sa)); CK(cudaStreamWaitEvent(sb, fork)); CK(cudaStreamWaitEvent(sc, fork));
      for (int k = 0; k < steps; ++k) {
        if (k >= 2) CK(cudaStreamWaitEvent(sa, evU[k - 2]));
        if (pattern >= 2 && k >= 8 && k % 4 == 0) CK(cudaStreamWaitEvent(sa, evB[k - 8]));
        spin_kernel<<<1, 384, 0, sa>>>(ns(d_us), sink);
        CK(cudaEventRecord(evD[k], sa));
        CK(cudaStreamWaitEvent(sb, evD[k]));
        spin_kernel<<<ncta, 256, sm3, sb>>>(ns(p_us), sink);
        CK(cudaEventRecord(evP[k], sb));
        spin_kernel<<<ncta, 256, sm3, sb>>>(ns(u_us), sink);
        CK(cudaEventRecord(evU[k], sb));
        if (pattern >= 2 && k % 4 == 3) {
          CK(cudaStreamWaitEvent(sc, evP[k]));
          const int nb = pattern == 3 ? 4 : 1;                     // pattern 3: bulk in 4 smaller launches
          const double f = double(steps - k) / steps;
          for (int q = 0; q < nb; ++q) spin_kernel<<<int(bulk_cta * f * f / nb) + 1, 256, sm3, sc>>>(ns(b_us), sink);
          CK(cudaEventRecord(evB[k - 3], sc));
        }
      }
      CK(cudaStreamWaitEvent(sa, evU[steps - 1]));
      if (pattern >= 2) { cudaEvent_t j; CK(cudaEventCreateWithFlags(&j, cudaEventDisableTiming)); CK(cudaEventRecord(j, sc)); CK(cudaStreamWaitEvent(sa, j)); }
    }
    CK(cudaStreamEndCapture(sa, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    cudaEvent_t t0, t1; CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
    for (int w = 0; w < 3; ++w) CK(cudaGraphLaunch(ge, sa));
    CK(cudaStreamSynchronize(sa));
    const int reps = 20;
    CK(cudaEventRecord(t0, sa));
    for (int r = 0; r < reps; ++r) CK(cudaGraphLaunch(ge, sa));
    CK(cudaEventRecord(t1, sa));
    CK(cudaStreamSynchronize(sa));
    float ms; CK(cudaEventElapsedTime(&ms, t0, t1));
    const double ideal = pattern == 0 ? steps * (d_us - 5 + p_us + u_us) : steps * d_us;
    printf("pattern %d: %.1f us per graph (%d steps; sum of the critical kernels %.1f us; %.2f us overhead per step)\n", pattern,
           ms * 1000 / reps, steps, ideal, (ms * 1000 / reps - ideal) / steps);
    CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
  }
  return 0;
}
