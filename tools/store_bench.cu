// store_bench.cu -- microbenchmark of the K2 output write pattern on B200 (tuning aid, not product).
// out[plane(4)][f(nf)][col(n)] complex128; block = tile of TC columns x chunk of FC frequencies.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bench store_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int OP> __device__ __forceinline__ void st16(double2* p, double a, double b) {
  if (OP == 0) *p = make_double2(a, b);
  else if (OP == 1) __stcs(p, make_double2(a, b));
  else if (OP == 2) __stwt(p, make_double2(a, b));
  else if (OP == 3) asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
  else if (OP == 4) __stcg(p, make_double2(a, b));
}

__global__ void k_linear(double2* out, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) __stcs(out + i, make_double2(1.0, 0.0));
}

// CPT columns per thread (strided by blockDim so every warp store stays 512 B contiguous)
template <int OP, int CPT>
__global__ void __launch_bounds__(256) k_pattern(double2* out, int64_t n, int nf, int fchunk, int nfchunks, int do_exp, int tile_fast = 0, int ntiles = 1) {
  int64_t tile = blockIdx.x / nfchunks;
  int chunk = blockIdx.x - tile * nfchunks;
  if (tile_fast) { chunk = blockIdx.x / ntiles; tile = blockIdx.x - (int64_t)chunk * ntiles; }
  const int f0 = chunk * fchunk;
  const int nfc = min(fchunk, nf - f0);
  const int64_t plane = (int64_t)nf * n;
  int64_t col[CPT];
  double a[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) { col[j] = tile * 256 * CPT + j * 256 + threadIdx.x; a[j] = -0.8 + 1e-6 * (double)(col[j] & 1023); }
#pragma unroll 4
  for (int k = 0; k < nfc; ++k) {
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      if (col[j] < n) {
        double v = do_exp ? exp(a[j] * (0.001 * (double)(f0 + k))) : a[j];
        double2* d = out + (int64_t)(f0 + k) * n + col[j];
        st16<OP>(d, v, 0.0);
        st16<OP>(d + plane, v, -v);
        st16<OP>(d + 2 * plane, v, v);
        st16<OP>(d + 3 * plane, v, 0.0);
      }
    }
  }
}

// two adjacent columns per thread, one 256-bit store per plane per frequency (STG.256)
__global__ void __launch_bounds__(256) k_pattern_v4(double2* out, int64_t n, int nf, int fchunk, int nfchunks, int tile_fast, int ntiles) {
  int64_t tile = blockIdx.x / nfchunks;
  int chunk = blockIdx.x - tile * nfchunks;
  if (tile_fast) { chunk = blockIdx.x / ntiles; tile = blockIdx.x - (int64_t)chunk * ntiles; }
  const int f0 = chunk * fchunk;
  const int nfc = min(fchunk, nf - f0);
  const int64_t plane = (int64_t)nf * n;
  const int64_t col = (tile * 256 + threadIdx.x) * 2;
  if (col >= n) return;
  const double a0 = -0.8 + 1e-6 * (double)(col & 1023), a1 = -0.8 + 1e-6 * (double)((col + 1) & 1023);
#pragma unroll 4
  for (int k = 0; k < nfc; ++k) {
    const double v0 = exp(a0 * (0.001 * (double)(f0 + k))), v1 = exp(a1 * (0.001 * (double)(f0 + k)));
    double2* d = out + (int64_t)(f0 + k) * n + col;
    asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(d), "d"(v0), "d"(0.0), "d"(v1), "d"(0.0) : "memory");
    asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(d + plane), "d"(v0), "d"(-v0), "d"(v1), "d"(-v1) : "memory");
    asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(d + 2 * plane), "d"(v0), "d"(v0), "d"(v1), "d"(v1) : "memory");
    asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(d + 3 * plane), "d"(v0), "d"(0.0), "d"(v1), "d"(0.0) : "memory");
  }
}


// the K2 pattern with a few reads mixed in: every `period` channels a thread fetches 16 doubles (coalesced rows of src,
// fresh addresses, i.e. from DRAM) before it goes on storing -- what the banded subband kernel does at a change of break
// interval (period ~54 on C4: 2.4 B read per 64 B written)
__global__ void __launch_bounds__(256) k_pattern_reads(double2* out, const double* __restrict__ src, int64_t n, int nf, int fchunk,
                                                       int nfchunks, int period) {
  const int64_t tile = blockIdx.x / nfchunks;
  const int chunk = blockIdx.x - tile * nfchunks;
  const int f0 = chunk * fchunk;
  const int nfc = min(fchunk, nf - f0);
  const int64_t plane = (int64_t)nf * n;
  const int64_t col = tile * 256 + threadIdx.x;
  if (col >= n) return;
  double a = -0.8 + 1e-6 * (double)(col & 1023);
#pragma unroll 4
  for (int k = 0; k < nfc; ++k) {
    if (period > 0 && (f0 + k) % period == 0) {
      const double* r = src + (int64_t)((f0 + k) / period) * 16 * n + col;
      double acc = 0.0;
#pragma unroll
      for (int q = 0; q < 16; ++q) acc += __ldg(r + (int64_t)q * n);
      a = -0.8 + 1e-9 * acc;
    }
    const double v = exp(a * (0.001 * (double)(f0 + k)));
    double2* d = out + (int64_t)(f0 + k) * n + col;
    st16<1>(d, v, 0.0);
    st16<1>(d + plane, v, -v);
    st16<1>(d + 2 * plane, v, v);
    st16<1>(d + 3 * plane, v, 0.0);
  }
}

// plane-major variant: block loops planes outermost (recompute), to see if fewer concurrent streams help
template <int OP>
__global__ void __launch_bounds__(256) k_plane_outer(double2* out, int64_t n, int nf, int fchunk, int nfchunks) {
  const int64_t tile = blockIdx.x / nfchunks;
  const int chunk = blockIdx.x - tile * nfchunks;
  const int f0 = chunk * fchunk;
  const int nfc = min(fchunk, nf - f0);
  const int64_t plane = (int64_t)nf * n;
  const int64_t col = tile * 256 + threadIdx.x;
  if (col >= n) return;
  const double a = -0.8 + 1e-6 * (double)(col & 1023);
  for (int p = 0; p < 4; ++p) {
#pragma unroll 4
    for (int k = 0; k < nfc; ++k) {
      double v = exp(a * (0.001 * (double)(f0 + k)));
      st16<OP>(out + p * plane + (int64_t)(f0 + k) * n + col, v, p == 1 ? -v : (p == 2 ? v : 0.0));
    }
  }
}

template <typename F> float timeit(F f, int reps = 5) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

template <int OP, int CPT> void run_pattern(double2* out, int64_t n, int nf, int fchunk, int do_exp, const char* opname, int smem = 0, int tile_fast = 0) {
  const int nfchunks = (nf + fchunk - 1) / fchunk;
  const int64_t tiles = (n + 256 * CPT - 1) / (256 * CPT);
  const double bytes = 64.0 * nf * n;
  if (smem > 0) CK(cudaFuncSetAttribute(k_pattern<OP, CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  float ms = timeit([&] { k_pattern<OP, CPT><<<(unsigned)(tiles * nfchunks), 256, smem>>>(out, n, nf, fchunk, nfchunks, do_exp, tile_fast, (int)tiles); });
  printf("pattern op=%-8s cpt=%d fchunk=%4d exp=%d n=%lld smem=%6d tilefast=%d : %7.3f ms  %7.1f GB/s\n", opname, CPT, fchunk, do_exp, (long long)n, smem, tile_fast, ms, bytes / ms / 1e6);
}

int main(int argc, char** argv) {
  int64_t n = argc > 1 ? atoll(argv[1]) : 150000;
  int nf = argc > 2 ? atoi(argv[2]) : 1024;
  const int64_t total = 4ll * nf * n;
  double2* out; CK(cudaMalloc(&out, total * sizeof(double2)));
  double2* src; CK(cudaMalloc(&src, total * sizeof(double2) / 2));
  const double bytes = 16.0 * total;
  float ms = timeit([&] { CK(cudaMemsetAsync(out, 0, total * sizeof(double2))); });
  printf("cudaMemset                      : %7.3f ms  %7.1f GB/s\n", ms, bytes / ms / 1e6);
  ms = timeit([&] { CK(cudaMemcpyAsync(out, src, total * sizeof(double2) / 2, cudaMemcpyDeviceToDevice)); });
  printf("cudaMemcpy D2D (r+w bytes)      : %7.3f ms  %7.1f GB/s\n", ms, bytes / ms / 1e6);
  for (int blocks : {148 * 8, 148 * 16, 148 * 32}) {
    ms = timeit([&] { k_linear<<<blocks, 256>>>(out, total); });
    printf("linear store grid=%5d          : %7.3f ms  %7.1f GB/s\n", blocks, ms, bytes / ms / 1e6);
  }
  if (argc > 3 && argv[3][0] == 'r') {  // a few reads mixed into the write stream
    for (int fc : {64}) for (int period : {0, 54, 27, 13, 4}) {
      const int nfchunks = (nf + fc - 1) / fc; const int64_t tiles = (n + 255) / 256;
      ms = timeit([&] { k_pattern_reads<<<(unsigned)(tiles * nfchunks), 256>>>(out, (const double*)src, n, nf, fc, nfchunks, period); });
      const double rd = period > 0 ? 128.0 * ((nf + period - 1) / period) * n : 0.0;
      printf("reads mixed in: period=%3d (%.2f B read per 64 B written) : %7.3f ms  write %7.1f GB/s  total %7.1f GB/s\n", period,
             rd / (1.0 * nf * n), ms, 64.0 * nf * n / ms / 1e6, (64.0 * nf * n + rd) / ms / 1e6);
    }
    return 0;
  }
  if (argc > 3 && argv[3][0] == 'v') {  // 256-bit stores, two adjacent columns per thread
    for (int tf : {0, 1}) for (int fc : {16, 32, 64, 128}) for (int smem : {0, 72 * 1024, 112 * 1024}) {
      const int nfchunks = (nf + fc - 1) / fc; const int64_t tiles = (n + 511) / 512;
      if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_pattern_v4, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      ms = timeit([&] { k_pattern_v4<<<(unsigned)(tiles * nfchunks), 256, smem>>>(out, n, nf, fc, nfchunks, tf, (int)tiles); });
      printf("v4 pattern fchunk=%4d smem=%6d tilefast=%d : %7.3f ms  %7.1f GB/s\n", fc, smem, tf, ms, 64.0 * nf * n / ms / 1e6);
    }
    for (int fc : {32, 64}) for (int smem : {0, 72 * 1024}) run_pattern<1, 1>(out, n, nf, fc, 1, "cs", smem, 0);
    return 0;
  }
  if (argc > 3 && argv[3][0] == 'w') {  // how few warps per SM still saturate the stores? (1 block of 256 per SM = 8 warps)
    for (int fc : {64, 256, 1024}) for (int smem : {0, 112 * 1024, 200 * 1024})
      run_pattern<1, 1>(out, n, nf, fc, 1, "cs", smem, 0);
    return 0;
  }
  if (argc > 3 && argv[3][0] == 'g') {  // grid order x chunk x occupancy sweep on a 256B-aligned width
    for (int tf : {0, 1}) for (int fc : {8, 16, 32, 64, 128, 256}) for (int smem : {0, 56 * 1024, 112 * 1024})
      run_pattern<1, 1>(out, n, nf, fc, 1, "cs", smem, tf);
    for (int tf : {0, 1}) for (int fc : {16, 64}) { run_pattern<1, 2>(out, n, nf, fc, 1, "cs", 0, tf); run_pattern<1, 4>(out, n, nf, fc, 1, "cs", 0, tf); }
    return 0;
  }
  if (argc > 3) {  // occupancy sensitivity: dynamic smem limits resident blocks per SM (227 KB / smem)
    for (int fc : {64, 205}) for (int smem : {0, 36 * 1024, 54 * 1024, 72 * 1024, 110 * 1024})
      run_pattern<1, 1>(out, n, nf, fc, 1, "cs", smem);
    for (int64_t nn : {n - 1, n - 2, n - 4, n - 6, n - 8, n - 16}) for (int fc : {64, 205}) run_pattern<1, 1>(out, nn, nf, fc, 1, "cs");
    for (int64_t nn : {n - 1, n - 2}) for (int fc : {64}) { run_pattern<0, 1>(out, nn, nf, fc, 1, "default"); run_pattern<1, 1>(out, nn, nf, fc, 1, "cs", 110 * 1024); }
    return 0;
  }
  for (int fc : {32, 64, 128, 205, 256, 1024}) run_pattern<1, 1>(out, n, nf, fc, 1, "cs");
  for (int fc : {64, 205, 1024}) run_pattern<1, 1>(out, n, nf, fc, 0, "cs");
  for (int fc : {64, 205}) { run_pattern<0, 1>(out, n, nf, fc, 1, "default"); run_pattern<2, 1>(out, n, nf, fc, 1, "wt"); run_pattern<3, 1>(out, n, nf, fc, 1, "noalloc"); run_pattern<4, 1>(out, n, nf, fc, 1, "cg"); }
  for (int fc : {32, 64, 128, 256}) { run_pattern<1, 2>(out, n, nf, fc, 1, "cs"); run_pattern<1, 4>(out, n, nf, fc, 1, "cs"); run_pattern<1, 8>(out, n, nf, fc, 1, "cs"); }
  for (int fc : {64, 256}) { run_pattern<0, 4>(out, n, nf, fc, 1, "default"); }
  for (int fc : {64, 256, 1024}) {
    const int nfchunks = (nf + fc - 1) / fc; const int64_t tiles = (n + 255) / 256;
    ms = timeit([&] { k_plane_outer<1><<<(unsigned)(tiles * nfchunks), 256>>>(out, n, nf, fc, nfchunks); });
    printf("plane-outer cs fchunk=%4d        : %7.3f ms  %7.1f GB/s\n", fc, ms, 64.0 * nf * n / ms / 1e6);
  }
  return 0;
}
