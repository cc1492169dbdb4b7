// ddio_probe.cu -- can the host consume Hermitian-packed slabs out of the last-level cache instead of DRAM?
// T worker threads, each with its own stream: D2H copy of S bytes into one of two pinned staging buffers, then expand the
// other one (8 bytes in -> 16 bytes out, non-temporal stores, the traffic pattern of rsk_unpack_coherency_host) into a
// large output region.  If DMA writes land in the LLC (DDIO) and are consumed before they are evicted, small S sustains a
// higher aggregate D2H rate than large S, whose staging data makes a round trip through DRAM.  (tuning aid)
// nvcc -O3 -std=c++17 -Xcompiler -march=native,-pthread -o ddio_probe ddio_probe.cu
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include <immintrin.h>

static int g_wc = 0;  // staging in write-combined pinned memory, read with streaming loads

static void expand(const double* in, double* out, size_t n) {  // n doubles in -> 2n doubles out
  size_t i = 0;
#if defined(__AVX512F__)
  const __m512d zero = _mm512_setzero_pd();
  for (; i + 8 <= n; i += 8) {
    const __m512d v = g_wc ? _mm512_castsi512_pd(_mm512_stream_load_si512((void*)(in + i))) : _mm512_loadu_pd(in + i);
    _mm512_stream_pd(out + 2 * i, _mm512_mask_expand_pd(zero, 0x55, v));
    _mm512_stream_pd(out + 2 * i + 8, _mm512_mask_expand_pd(zero, 0x55, _mm512_castpd256_pd512(_mm512_extractf64x4_pd(v, 1))));
  }
#endif
  for (; i < n; ++i) _mm_stream_pd(out + 2 * i, _mm_set_pd(0.0, in[i]));
  _mm_sfence();
}

int main(int argc, char** argv) {
  const int T = argc > 1 ? atoi(argv[1]) : 12;
  const double secs = argc > 2 ? atof(argv[2]) : 0.6;
  g_wc = argc > 3 ? atoi(argv[3]) : 0;
  const size_t out_bytes = 256ull << 20;  // per worker, wrapped
  const size_t sizes[] = {1ull << 20, 4ull << 20, 16ull << 20, 64ull << 20};
  for (size_t S : sizes) {
    std::atomic<long long> copied{0};
    std::atomic<bool> stop{false};
    std::vector<std::thread> pool;
    for (int t = 0; t < T; ++t) {
      pool.emplace_back([&, t]() {
        cudaSetDevice(0);
        cudaStream_t st;
        cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        double *dev, *stage[2], *out;
        cudaMalloc(&dev, S);
        cudaMemset(dev, 1, S);
        cudaHostAlloc(&stage[0], S, g_wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
        cudaHostAlloc(&stage[1], S, g_wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
        out = static_cast<double*>(aligned_alloc(4096, out_bytes));
        memset(out, 0, out_bytes);
        if (!g_wc) {
          memset(stage[0], 0, S);
          memset(stage[1], 0, S);
        }
        cudaEvent_t ev[2];
        cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
        size_t pos = 0;
        int b = 0;
        long long mine = 0;
        bool first = true;
        while (!stop.load(std::memory_order_relaxed)) {
          cudaMemcpyAsync(stage[b], dev, S, cudaMemcpyDeviceToHost, st);
          cudaEventRecord(ev[b], st);
          if (!first) {
            cudaEventSynchronize(ev[1 - b]);
            if (pos + 2 * S > out_bytes) pos = 0;
            expand(stage[1 - b], out + pos / 8, S / 8);
            pos += 2 * S;
            mine += (long long)S;
          }
          first = false;
          b ^= 1;
        }
        cudaStreamSynchronize(st);
        copied += mine;
        cudaFree(dev); cudaFreeHost(stage[0]); cudaFreeHost(stage[1]); free(out);
      });
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(150));  // warm-up (allocation, first touches)
    const long long c0 = 0;
    (void)c0;
    const auto t0 = std::chrono::steady_clock::now();
    std::this_thread::sleep_for(std::chrono::duration<double>(secs));
    stop = true;
    for (auto& th : pool) th.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() + 0.15;
    printf("wc %d threads %2d  slab %8zu KB  staging total %7.1f MB  D2H+expand %.1f GB/s in (%.1f GB/s out)\n", g_wc, T, S >> 10,
           T * 2.0 * S / 1048576.0, copied.load() / dt / 1e9, 2.0 * copied.load() / dt / 1e9);
    fflush(stdout);
  }
  return 0;
}
