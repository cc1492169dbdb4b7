// dmma_probe.cu -- does the fp64 tensor path (mma.sync m8n8k4 f64, "DMMA") of B200 run beside the fp64 FMA pipe, or
// on it?  (tuning aid for the LST-sweep kernel, which is bound by fp64 issue: if DMMA is a separate pipe, the three
// K=3 products of the position chain -- rotation, p.eh, p.v -- can move there.)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// mode 0: DFMA only (NF independent chains); 1: DMMA only (NM independent accumulator pairs); 2: both in every warp;
// 3: even warps DFMA, odd warps DMMA
template <int MODE>
__global__ void __launch_bounds__(256) probe(double* out, int iters, double x) {
  double f[8];
  double m0[4], m1[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) { m0[i] = i; m1[i] = -i; }
  const double a = x + threadIdx.x * 1e-6, b = x - threadIdx.x * 1e-6;
  const bool odd = (threadIdx.x >> 5) & 1;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2 || (MODE == 3 && !odd)) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fma(f[i], a, b);
    }
    if (MODE == 1 || MODE == 2 || (MODE == 3 && odd)) {
#pragma unroll
      for (int i = 0; i < 4; ++i) dmma(m0[i], m1[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += f[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += m0[i] + m1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
float run(double* out, int blocks, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE><<<blocks, 256>>>(out, iters, 1.0000001);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe<MODE><<<blocks, 256>>>(out, iters, 1.0000001);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  const int blocks = sms * 4, iters = 20000;
  double* out;
  cudaMalloc(&out, sizeof(double) * blocks * 256);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double warps = blocks * 8.0;
  const float t0 = run<0>(out, blocks, iters), t1 = run<1>(out, blocks, iters), t2 = run<2>(out, blocks, iters),
              t3 = run<3>(out, blocks, iters);
  // per SM per clock (nominal max clock): DFMA lane-ops, DMMA FMAs (256 per warp instruction)
  const double clk = clk_khz * 1e3;
  printf("sms %d clock %.0f MHz\n", sms, clk / 1e6);
  printf("DFMA only : %.3f ms  %.1f DFMA lanes/clk/SM\n", t0, warps * iters * 8 * 32 / (t0 * 1e-3) / clk / sms);
  printf("DMMA only : %.3f ms  %.1f FMA/clk/SM (%.2f DMMA warp-instr/clk/SM)\n", t1, warps * iters * 4 * 256 / (t1 * 1e-3) / clk / sms,
         warps * iters * 4 / (t1 * 1e-3) / clk / sms);
  printf("both/warp : %.3f ms  (sum of the two alone %.3f ms, max %.3f ms)\n", t2, t0 + t1, t0 > t1 ? t0 : t1);
  printf("split warps: %.3f ms (half the warps each; alone-at-half-work %.3f / %.3f ms)\n", t3, t0 / 2, t1 / 2);
  return 0;
}
