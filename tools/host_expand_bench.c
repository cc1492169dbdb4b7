// host_expand_bench.c -- can the host unpack a Hermitian-packed coherency slab (4 doubles -> 2x2 complex =
// 8 doubles per eval) fast enough to beat sending the full 64 B/eval over PCIe?  (tuning aid)
// gcc -O3 -march=native -fopenmp -o host_expand_bench host_expand_bench.c
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <immintrin.h>

int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : 150000;   // columns
  const long nf = argc > 2 ? atol(argv[2]) : 1024;
  const long rows = nf;
  double* packed = (double*)aligned_alloc(64, sizeof(double) * 4 * rows * n);  // [4][nf][n]: h00,h11,re01,im10
  double* out = (double*)aligned_alloc(64, sizeof(double) * 8 * rows * n);     // [2][2][nf][n] complex
  if (!packed || !out) return 1;
#pragma omp parallel for
  for (long i = 0; i < 4 * rows * n; ++i) packed[i] = (double)(i & 1023);
#pragma omp parallel for
  for (long i = 0; i < 8 * rows * n; ++i) out[i] = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    const double t0 = omp_get_wtime();
#pragma omp parallel for schedule(static)
    for (long r = 0; r < rows; ++r) {
      const double* h00 = packed + (0 * rows + r) * n;
      const double* h11 = packed + (1 * rows + r) * n;
      const double* re = packed + (2 * rows + r) * n;
      const double* im = packed + (3 * rows + r) * n;
      double* o00 = out + 2 * ((0 * rows + r) * n);
      double* o01 = out + 2 * ((1 * rows + r) * n);
      double* o10 = out + 2 * ((2 * rows + r) * n);
      double* o11 = out + 2 * ((3 * rows + r) * n);
      for (long c = 0; c < n; ++c) {
        _mm_stream_pd(o00 + 2 * c, _mm_set_pd(0.0, h00[c]));
        _mm_stream_pd(o01 + 2 * c, _mm_set_pd(-im[c], re[c]));
        _mm_stream_pd(o10 + 2 * c, _mm_set_pd(im[c], re[c]));
        _mm_stream_pd(o11 + 2 * c, _mm_set_pd(0.0, h11[c]));
      }
    }
    const double dt = omp_get_wtime() - t0;
    printf("threads=%d expand %.1f ms  out %.1f GB/s  (read+write %.1f GB/s)\n", omp_get_max_threads(), dt * 1e3,
           64.0 * rows * n / dt / 1e9, 96.0 * rows * n / dt / 1e9);
  }
  // plain memcpy of the same output volume for reference
  double* out2 = (double*)aligned_alloc(64, sizeof(double) * 8 * rows * n);
  for (int rep = 0; rep < 2; ++rep) {
    const double t0 = omp_get_wtime();
#pragma omp parallel for schedule(static)
    for (long r = 0; r < 4 * rows; ++r) memcpy(out2 + 2 * r * n, out + 2 * r * n, 16 * n);
    const double dt = omp_get_wtime() - t0;
    printf("parallel memcpy %.1f ms  %.1f GB/s copied\n", dt * 1e3, 64.0 * rows * n / dt / 1e9);
  }
  return 0;
}
