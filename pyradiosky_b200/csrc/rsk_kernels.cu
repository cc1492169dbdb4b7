// rsk_kernels.cu -- sm_100a kernels and the C-ABI entry points that launch them (see include/rsk.h).
//
// Everything here is fp64 streaming work bounded by HBM bandwidth (nothing is a dense contraction, so
// tensor cores are not used).  Layout rules: structure-of-arrays with the component axis fastest,
// one thread per component, 128-bit streaming stores for the complex128 output, the frequency axis
// (log-frequencies, spline interval + local coordinate, or channel map) staged in shared memory once
// per block, and grids linearised so that blocks that share a component tile run back to back and
// re-read its few hundred input bytes from L2 instead of HBM.  Output rows should start on 256-byte
// boundaries (ld_out % 16 == 0): measured +25-30 % HBM write throughput on B200.
#include <cuda_pipeline.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/rsk.h"
#include "rsk_device.cuh"

namespace {

constexpr int kThreads = 256;       // threads per block for the component-parallel kernels
constexpr int kMaxFChunk = 256;     // frequencies staged in shared memory per block
// K2 launch shape, from sweeps of the real kernel on B200 (profiles/r01_k2_launch_shape_sweep_c2.log,
// r01_k2_launch_shape_sweep_unpolarized.log, r01_k2_register_cap_sweep.log): 64 frequencies
// per block, frequency chunk fastest in the grid, no residency cap.  RSK_K2_FCHUNK / RSK_K2_TILEFAST /
// RSK_K2_SMEM override them for tuning runs.
constexpr int kDefaultFChunk = 64;
constexpr int kFullCmFChunk = 32;            // FULL mode reading the component-major copy (measured)
#ifndef RSK_FULL_BATCH
#define RSK_FULL_BATCH 8
#endif
// FULL mode: channels whose loads a thread issues together.  Measured on the stored (4, 1024, 300k) cube, component-major
// reads: 1 / 2 / 4 / 8 / 16 -> 0.857 / 0.865 / 0.871 / 0.929 / 0.855 of the 96 B/eval roofline (8: 110 registers, two
// blocks per SM; 16: 176 registers, one); a cp.async ring (0.63) and L2 prefetches of the chunk (0.84) are slower.
constexpr int kFullBatch = RSK_FULL_BATCH;
#ifndef RSK_CM_DENSITY8
#define RSK_CM_DENSITY8 8  // blocks whose columns are at least this / 8 dense in the catalogue read the (4, Nf, N) cube
#endif
#ifndef RSK_FULL_BATCH_PLAIN
#define RSK_FULL_BATCH_PLAIN 4
#endif
constexpr int kFullBatchPlain = RSK_FULL_BATCH_PLAIN;  // the same for the gathered (4, Nf, N) reads (four 8-byte loads per channel)
constexpr int kSubbandFChunk = 256;          // longer chunks: the per-block staging cost is amortised (measured)
constexpr int kMaxDynSmem = 200 * 1024;      // dynamic shared memory budget per block
constexpr int kDefaultTileFast = 0;
constexpr int kDefaultOccSmem = 0;  // > 0: dynamic smem request that caps resident blocks per SM
constexpr int kCompactItems = 8;    // mask bytes per thread in the compaction kernels
constexpr int kCompactTile = kThreads * kCompactItems;  // 2048, see rsk_workspace_bytes
#ifndef RSK_SWEEP_TIMES
#define RSK_SWEEP_TIMES 64
#endif
constexpr int kSweepTimes = RSK_SWEEP_TIMES;  // time steps per block in the sweep kernel
constexpr int kSubbandBanded = 4;   // internal mode: RSK_SUBBAND evaluated from precomputed spline coefficients

struct CoordSrc {
  const double* lon;
  const double* lat;
  int64_t nside;
  int nested;
  const int64_t* hpx_inds;
  int64_t hpx_first;
  double* rot;  // optional [3][n] rotation coefficients (written by K1, read by K2)
};

// sin/cos of the catalogue-frame longitude and latitude of component i
__device__ __forceinline__ void frame_trig(const CoordSrc& c, int64_t i, double& sl, double& cl, double& sb,
                                           double& cb) {
  if (c.nside > 0) {
    const int64_t pix = c.hpx_inds ? __ldg(c.hpx_inds + i) : c.hpx_first + i;
    double z, st, phi;
    rsk_hpx_center(c.nside, c.nested, pix, z, st, phi);
    rsk_sincos(phi, &sl, &cl);
    sb = z;
    cb = st;
  } else {
    rsk_sincos(__ldg(c.lon + i), &sl, &cl);
    rsk_sincos(__ldg(c.lat + i), &sb, &cb);
  }
}

// ---------------------------------------------------------------------------------------------------
// K1: positions (one time step)
// ---------------------------------------------------------------------------------------------------
struct Ravg {
  double m[9];
  int valid;
};

// The per-component body of K1, shared by the two entry kernels below.
__device__ __forceinline__ void positions_body(const CoordSrc& cs, const int64_t n, const int64_t i, const RskCtx& ctx,
                                               const double* ravg_m, const bool have_ravg, double* __restrict__ alt_az,
                                               double* __restrict__ lmn, uint8_t* __restrict__ above) {
  double sl, cl, sb, cb;
  frame_trig(cs, i, sl, cl, sb, cb);
  const RskAltAz o = rsk_chain(ctx, cl * cb, sl * cb, sb);
  if (have_ravg) {
    // Polarisation-rotation coefficients for K2.  sin/cos of alt and az come from the direction cosines
    // (l, m, n) = (sin az cos alt, cos az cos alt, sin alt), so no further trigonometry is needed.
    const double se = o.n;
    const double ce = rsk_sqrt(o.l * o.l + o.m * o.m);
    const double cinv = ce > 0.0 ? rsk_div(1.0, ce) : 0.0;
    const double sa = ce > 0.0 ? o.l * cinv : 0.0;
    const double ca = ce > 0.0 ? o.m * cinv : 1.0;
    const RskRot r = rsk_basis_rotation(ravg_m, sl, cl, sb, cb, sa, ca, se, ce);
    cs.rot[i] = r.n2;
    cs.rot[n + i] = r.c2;
    cs.rot[2 * n + i] = r.s2;
  }
  __stcs(alt_az + i, o.alt);
  __stcs(alt_az + n + i, o.az);
  if (lmn) {
    __stcs(lmn + i, o.l);
    __stcs(lmn + n + i, o.m);
    __stcs(lmn + 2 * n + i, o.n);
  }
  if (above) above[i] = o.alt > 0.0 ? 1 : 0;  // skymodel.py:2479
}

__global__ void __launch_bounds__(kThreads) k_positions(const CoordSrc cs, const int64_t n,
                                                        const __grid_constant__ RskCtx ctx,
                                                        const __grid_constant__ Ravg ravg,
                                                        double* __restrict__ alt_az, double* __restrict__ lmn,
                                                        uint8_t* __restrict__ above) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  positions_body(cs, n, i, ctx, ravg.m, ravg.valid != 0, alt_az, lmn, above);
}

// The same with the context and R_avg read from device memory (RSK_STEP_LEN doubles: packed context, R_avg, flag):
// the launch then has no time-dependent argument, so a whole time step can be captured in a CUDA graph once and
// replayed after refreshing that small buffer.
__global__ void __launch_bounds__(kThreads) k_positions_ctxdev(const CoordSrc cs, const int64_t n,
                                                               const double* __restrict__ step, double* __restrict__ alt_az,
                                                               double* __restrict__ lmn, uint8_t* __restrict__ above) {
  __shared__ double s_step[RSK_STEP_LEN];
  if (threadIdx.x < RSK_STEP_LEN) s_step[threadIdx.x] = __ldg(step + threadIdx.x);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const RskCtx& ctx = *reinterpret_cast<const RskCtx*>(s_step);
  positions_body(cs, n, i, ctx, s_step + RSK_CTX_LEN, s_step[RSK_CTX_LEN + 9] != 0.0 && cs.rot != nullptr, alt_az, lmn, above);
}

// minimum resident blocks per SM asked of ptxas for the sweep kernel (register cap); -DRSK_SWEEP_MINB=n to tune
#ifndef RSK_SWEEP_MINB
#define RSK_SWEEP_MINB 3  // register cap: 3 blocks of 256 threads per SM measured best for the round-2 loop (6.38 ms against
                          // 6.47 with 4 and 6.73 with 2 for 1M x 1000; round 1's longer loop preferred 4, profiles/r01_sweep_tuning.log)
#endif

// 1: the sweep's folded, no-lmn path takes alt / az from rsk_sweep_angles (30 instead of 43 fp64 instructions, below
// 1e-13 rad); 0: the full-margin routines of the single-time kernel.  -DRSK_SWEEP_FAST=0 to compare.
#ifndef RSK_SWEEP_UNROLL
#define RSK_SWEEP_UNROLL 2  // time steps per trip of the sweep loop (1 and 4 measured slower, profiles/r02_summary.md)
#endif
constexpr int kSweepUnroll = RSK_SWEEP_UNROLL;
#ifndef RSK_SWEEP_FAST
#define RSK_SWEEP_FAST 1
#endif

// The time loop of the sweep kernel.  MASK: the horizon mask is wanted; with CPT == 1 threads past the end of the
// catalogue have left already, so the loop body carries no bounds test.
template <int CPT, bool WANT_LMN, int FOLD, bool MASK>
__device__ __forceinline__ void sweep_times(const RskCtx* s_ctx, const int ntc, const int64_t n, const int64_t* idx,
                                            const double* px, const double* py, const double* pz, double** pa,
                                            double** pl, uint8_t** pm) {
#pragma unroll kSweepUnroll
  for (int k = 0; k < ntc; ++k) {
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      // (px, py, pz) is a product of sines and cosines, rotated by an orthogonal matrix at most: unit length
      const RskAltAz o = rsk_chain_t<WANT_LMN, true, FOLD, true, RSK_SWEEP_FAST != 0>(s_ctx[k], px[j], py[j], pz[j]);
      if (CPT == 1 || idx[j] < n) {
        __stcs(pa[j], o.alt);
        __stcs(pa[j] + n, o.az);
        if (WANT_LMN) {
          __stcs(pl[j], o.l);
          __stcs(pl[j] + n, o.m);
          __stcs(pl[j] + 2 * n, o.n);
        }
        // alt > 0 read off the high word (alt is never a positive denormal; a NaN is not "above"): integer compares,
        // not an fp64 one
        if (MASK) *pm[j] = (unsigned)(__double2hiint(o.alt) - 1) < 0x7fefffffu ? 1 : 0;
      }
      pa[j] += 2 * n;
      if (WANT_LMN) pl[j] += 3 * n;
      if (MASK) pm[j] += n;
    }
  }
}

// K1 sweep: T contexts per launch.  Block = (CPT x kThreads) components x kSweepTimes time steps; the
// contexts are staged in shared memory; the per-component trigonometry is hoisted out of the time loop;
// CPT components per thread share every context load and give the fp64 pipe independent chains.
template <int CPT, bool WANT_LMN>
__global__ void __launch_bounds__(kThreads, RSK_SWEEP_MINB) k_positions_sweep(const CoordSrc cs, const int64_t n,
                                                              const double* __restrict__ ctx_g, const int64_t nt,
                                                              const int ntchunks, double* __restrict__ alt_az,
                                                              double* __restrict__ lmn, uint8_t* __restrict__ above) {
  __shared__ RskCtx s_ctx[kSweepTimes];
  const int64_t tile = blockIdx.x / ntchunks;
  const int tchunk = (int)(blockIdx.x - tile * ntchunks);
  const int64_t t0 = (int64_t)tchunk * kSweepTimes;
  const int ntc = (int)min((int64_t)kSweepTimes, nt - t0);
  {
    double* dst = reinterpret_cast<double*>(s_ctx);
    const double* src = ctx_g + t0 * RSK_CTX_LEN;
    for (int k = threadIdx.x; k < ntc * RSK_CTX_LEN; k += kThreads) dst[k] = __ldg(src + k);
  }
  __syncthreads();
  int64_t idx[CPT];
  double px[CPT], py[CPT], pz[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    idx[j] = (tile * CPT + j) * kThreads + threadIdx.x;
    px[j] = py[j] = 0.0; pz[j] = 1.0;
    if (idx[j] < n) {
      double sl, cl, sb, cb;
      frame_trig(cs, idx[j], sl, cl, sb, cb);
      px[j] = cl * cb; py[j] = sl * cb; pz[j] = sb;
      if (s_ctx[0].m_ident == 0.0) {  // catalogue frame -> ICRS: the same matrix in every context, applied once
        const double* m = s_ctx[0].m;
        const double x = m[0] * px[j] + m[1] * py[j] + m[2] * pz[j];
        const double y = m[3] * px[j] + m[4] * py[j] + m[5] * pz[j];
        const double z = m[6] * px[j] + m[7] * py[j] + m[8] * pz[j];
        px[j] = x; py[j] = y; pz[j] = z;
      }
    }
  }
  // running output pointers (one 64-bit add per plane and time step instead of a 64-bit multiply per store)
  double* pa[CPT];
  double* pl[CPT];
  uint8_t* pm[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    pa[j] = alt_az + (2 * t0) * n + idx[j];
    pl[j] = WANT_LMN ? lmn + (3 * t0) * n + idx[j] : nullptr;
    pm[j] = above ? above + t0 * n + idx[j] : nullptr;
  }
  // every context of this block folded (the case for contexts built for the reference's AltAz frame)?  Decided
  // once here, one context per thread and a block-wide vote, so neither the time loop nor each thread's prologue
  // carries a per-step test (as a per-thread loop over the contexts it was a tenth of the kernel's instructions).
  bool mine = true;
  for (int k = threadIdx.x; k < ntc; k += kThreads) mine = mine && rsk_ctx_folded(s_ctx[k]);
  const bool folded = __syncthreads_and(mine) != 0;
  if (CPT == 1 && idx[0] >= n) return;  // (after the last barrier of the kernel)
  if (above) {
    if (folded) sweep_times<CPT, WANT_LMN, 1, true>(s_ctx, ntc, n, idx, px, py, pz, pa, pl, pm);
    else sweep_times<CPT, WANT_LMN, 0, true>(s_ctx, ntc, n, idx, px, py, pz, pa, pl, pm);
  } else {
    if (folded) sweep_times<CPT, WANT_LMN, 1, false>(s_ctx, ntc, n, idx, px, py, pz, pa, pl, pm);
    else sweep_times<CPT, WANT_LMN, 0, false>(s_ctx, ntc, n, idx, px, py, pz, pa, pl, pm);
  }
}

__global__ void __launch_bounds__(kThreads) k_healpix_lonlat(const int64_t nside, const int nested,
                                                             const int64_t* __restrict__ inds, const int64_t first,
                                                             const int64_t n, double* __restrict__ lon,
                                                             double* __restrict__ lat) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const int64_t pix = inds ? __ldg(inds + i) : first + i;
  double z, st, phi;
  rsk_hpx_center(nside, nested, pix, z, st, phi);
  lon[i] = phi;
  lat[i] = rsk_atan2_t<true>(z, st);  // st >= 0, (z, st) on the unit circle
}

// inverse of k_healpix_lonlat: the pixel each (lon, lat) falls in (astropy_healpix lonlat_to_healpix, called by
// SkyModel.assign_to_healpix at skymodel.py:1916)
__global__ void __launch_bounds__(kThreads) k_lonlat_to_healpix(const int64_t nside, const int nested,
                                                                const double* __restrict__ lon,
                                                                const double* __restrict__ lat, const int64_t n,
                                                                int64_t* __restrict__ pix) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  double sb, cb;
  rsk_sincos(__ldg(lat + i), &sb, &cb);
  pix[i] = rsk_hpx_pixel(nside, nested, sb, fabs(cb), __ldg(lon + i));
}

// x *= factor or x /= factor, correctly rounded (pixel-area scaling of healpix_to_point / assign_to_healpix,
// skymodel.py:1722, :1787, :1990-2003)
__global__ void __launch_bounds__(kThreads) k_scale(double* __restrict__ x, const int64_t count, const double factor,
                                                    const int divide) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= count) return;
  x[i] = divide ? __ddiv_rn(x[i], factor) : __dmul_rn(x[i], factor);
}

// Sum (or root-sum-square) of the components that fall in one pixel, in catalogue order (skymodel.py:1967-1979):
// out[r][j] = sum_{k in [start[j], start[j+1])} in[r][order[k]] for each of the nrows = 4 * Nfreqs rows.
__global__ void __launch_bounds__(kThreads) k_segment_sum(const double* __restrict__ in, const int64_t nrows,
                                                          const int64_t n_in, const int64_t* __restrict__ order,
                                                          const int64_t* __restrict__ start, const int64_t nseg,
                                                          const int quadrature, double* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t r = blockIdx.y;
  if (j >= nseg) return;
  const int64_t k0 = __ldg(start + j), k1 = __ldg(start + j + 1);
  const double* row = in + r * n_in;
  double acc = 0.0;
  for (int64_t k = k0; k < k1; ++k) {
    const double v = __ldg(row + __ldg(order + k));
    acc = quadrature ? __dadd_rn(acc, __dmul_rn(v, v)) : __dadd_rn(acc, v);
  }
  out[r * nseg + j] = quadrature ? __dsqrt_rn(acc) : acc;
}

// ---------------------------------------------------------------------------------------------------
// Component down-selection (SkyModel.select, cut_nonrising, calculate_rise_set_lsts)
// ---------------------------------------------------------------------------------------------------
// longitude and latitude of component i in the catalogue frame, as SkyModel.get_lon_lat gives them
__device__ __forceinline__ void frame_lonlat(const CoordSrc& c, int64_t i, double& lon, double& lat) {
  if (c.nside > 0) {
    const int64_t pix = c.hpx_inds ? __ldg(c.hpx_inds + i) : c.hpx_first + i;
    double z, st, phi;
    rsk_hpx_center(c.nside, c.nested, pix, z, st, phi);
    lon = phi;
    lat = rsk_atan2_t<true>(z, st);  // as k_healpix_lonlat
  } else {
    lon = __ldg(c.lon + i);
    lat = __ldg(c.lat + i);
  }
}

// One thread per candidate component: every enabled predicate of RskSelect, one byte out.  The Stokes reads are
// coalesced across components (the component axis is the fastest); only the planes a predicate needs are touched.
__global__ void __launch_bounds__(kThreads) k_select_mask(const CoordSrc cs, const double* __restrict__ stokes,
                                                          const int64_t nf, const int64_t n,
                                                          const int64_t* __restrict__ inds, const int64_t m,
                                                          const RskSelect p, uint8_t* __restrict__ keep) {
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (j >= m) return;
  const int64_t i = inds ? __ldg(inds + j) : j;
  bool k = i >= 0 && i < n;
  if (k && (p.use_lat || p.use_lon || p.cut_nonrising)) {
    double lon, lat;
    frame_lonlat(cs, i, lon, lat);
    if (p.use_lat) k = k && lat >= p.lat_lo && lat <= p.lat_hi;
    if (p.use_lon) {
      // skymodel.py:3036-3044: a range whose end is below its start wraps through lon = 0
      if (p.lon_hi < p.lon_lo) k = k && (lon >= p.lon_lo || lon <= p.lon_hi);
      else k = k && lon >= p.lon_lo && lon <= p.lon_hi;
    }
    if (p.cut_nonrising) k = k && !(p.tan_tel_lat * tan(lat) < -1.0);  // skymodel.py:3380-3381
  }
  if (k && (p.use_min || p.use_max || p.non_negative || p.non_nan)) {
    double mn = INFINITY, mx = -INFINITY;
    bool nan_used = false, any_neg = false, any_nan = false, all_nan = true;
    const bool cuts = p.use_min || p.use_max;
    const double* __restrict__ sp = stokes + i;
    // four channels per trip, every load of the trip issued before the first use (the planes are nf * n apart, so
    // each load is its own coalesced stream; what limits this kernel is the number of loads in flight)
    for (int64_t f0 = 0; f0 < nf; f0 += 4) {
      double si[4], sq[4], su[4], sv[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const int64_t f = f0 + d < nf ? f0 + d : nf - 1;  // (the tail repeats the last channel: idempotent tests)
        si[d] = __ldg(sp + f * n);
        if (p.non_nan) {
          sq[d] = __ldg(sp + (nf + f) * n);
          su[d] = __ldg(sp + (2 * nf + f) * n);
          sv[d] = __ldg(sp + (3 * nf + f) * n);
        }
      }
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const int64_t f = f0 + d < nf ? f0 + d : nf - 1;
        if (cuts && (!p.freq_use || __ldg(p.freq_use + f))) {
          nan_used = nan_used || isnan(si[d]);
          mn = si[d] < mn ? si[d] : mn;
          mx = si[d] > mx ? si[d] : mx;
        }
        any_neg = any_neg || si[d] < 0.0;
        if (p.non_nan) {
          const bool nanf = isnan(si[d]) || isnan(sq[d]) || isnan(su[d]) || isnan(sv[d]);
          any_nan = any_nan || nanf;
          all_nan = all_nan && nanf;
        }
      }
    }
    if (p.use_min) k = k && !nan_used && mn >= p.min_brightness;  // np.min(...) >= min (:3096-3098)
    if (p.use_max) k = k && !nan_used && mx <= p.max_brightness;  // np.max(...) <= max (:3120-3122)
    if (p.non_negative) k = k && !any_neg;                        // :3232
    if (p.non_nan == 1) k = k && !any_nan;                        // :3215-3218
    if (p.non_nan == 2) k = k && !(all_nan && nf > 0);            // :3222-3224
  }
  keep[j] = k ? 1 : 0;
}

// out[r][j] = in[r][idx[j]]: the index is read once per thread and reused for RPT rows (gridDim.y strides the rows)
template <typename T, int RPT>
__global__ void __launch_bounds__(kThreads) k_gather_rows(const T* __restrict__ in, const int64_t nrows,
                                                          const int64_t n_in, const int64_t* __restrict__ idx,
                                                          const int64_t n_out, T* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (j >= n_out) return;
  const int64_t i = __ldg(idx + j);
  for (int64_t r0 = (int64_t)blockIdx.y * RPT; r0 < nrows; r0 += (int64_t)gridDim.y * RPT) {
    T v[RPT];
#pragma unroll
    for (int q = 0; q < RPT; ++q)
      if (r0 + q < nrows) v[q] = __ldg(in + (r0 + q) * n_in + i);
#pragma unroll
    for (int q = 0; q < RPT; ++q)
      if (r0 + q < nrows) out[(r0 + q) * n_out + j] = v[q];
  }
}

// out[r] = a[r] followed by b[r]; a missing side (NULL) is filled with ``fill`` (NaN for the numeric parameters
// only one of two concatenated objects has, skymodel.py:2960-2999)
template <typename T>
__global__ void __launch_bounds__(kThreads) k_concat_rows(const T* __restrict__ a, const int64_t n_a,
                                                          const T* __restrict__ b, const int64_t n_b,
                                                          const int64_t nrows, const T fill, T* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t n = n_a + n_b;
  if (j >= n) return;
  for (int64_t r = blockIdx.y; r < nrows; r += gridDim.y) {
    T v = fill;
    if (j < n_a) {
      if (a) v = __ldg(a + r * n_a + j);
    } else if (b) {
      v = __ldg(b + r * n_b + (j - n_a));
    }
    out[r * n + j] = v;
  }
}

// (lon, lat) -> (lon', lat') under a rotation of the unit sphere: SkyCoord.transform_to between frames that differ by
// a fixed matrix (icrs, fk5(equinox), galactic), skymodel.py:1419-1446
struct Rot9 {
  double m[9];
};
__global__ void __launch_bounds__(kThreads) k_rotate_lonlat(const double* __restrict__ lon,
                                                            const double* __restrict__ lat, const int64_t n,
                                                            const Rot9 r, double* __restrict__ lon_out,
                                                            double* __restrict__ lat_out) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  double sl, cl, sb, cb;
  rsk_sincos(__ldg(lon + i), &sl, &cl);
  rsk_sincos(__ldg(lat + i), &sb, &cb);
  const double px = cl * cb, py = sl * cb, pz = sb;
  const double x = r.m[0] * px + r.m[1] * py + r.m[2] * pz;
  const double y = r.m[3] * px + r.m[4] * py + r.m[5] * pz;
  const double z = r.m[6] * px + r.m[7] * py + r.m[8] * pz;
  lon_out[i] = rsk_azimuth(y, x);  // [0, 2 pi], as astropy's Longitude wraps
  lat_out[i] = rsk_atan2_t<true>(z, rsk_sqrt(x * x + y * y));
}

// utils.coherency_to_stokes (utils.py:64-103): I = Re(C00 + C11), Q = Re(C00 - C11), U = Re(C01 + C10),
// V = -Im(C01 - C10); count = elements per plane, 64 B read + 32 B written each
__global__ void __launch_bounds__(kThreads) k_coherency_to_stokes(const double2* __restrict__ coh, const int64_t count,
                                                                  double* __restrict__ stokes) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= count) return;
  const double2 c00 = __ldcs(coh + i), c01 = __ldcs(coh + count + i);
  const double2 c10 = __ldcs(coh + 2 * count + i), c11 = __ldcs(coh + 3 * count + i);
  __stcs(stokes + i, c00.x + c11.x);
  __stcs(stokes + count + i, c00.x - c11.x);
  __stcs(stokes + 2 * count + i, c01.x + c10.x);
  __stcs(stokes + 3 * count + i, -(c01.y - c10.y));
}

// ---- HEALPix bilinear interpolation: SkyModel.healpix_interp_transform (skymodel.py:1448-1538) ----------------
// HEALPix ring geometry (get_ring_info2): first pixel, pixels in the ring, colatitude, half-pixel shift
__device__ __forceinline__ void hpx_ring_info(const int64_t nside, const int64_t ring, int64_t& start, int64_t& nr,
                                              double& theta, bool& shifted) {
  const int64_t npix = 12 * nside * nside, ncap = 2 * nside * (nside - 1);
  const int64_t north = ring > 2 * nside ? 4 * nside - ring : ring;
  if (north < nside) {
    const double tmp = (double)(north * north) * (4.0 / (double)npix);
    theta = atan2(sqrt(tmp * (2.0 - tmp)), 1.0 - tmp);
    nr = 4 * north;
    shifted = true;
    start = 2 * north * (north - 1);
  } else {
    theta = acos(fmin(1.0, fmax(-1.0, (double)(2 * nside - north) * (2.0 / (3.0 * (double)nside)))));
    nr = 4 * nside;
    shifted = ((north - nside) & 1) == 0;
    start = ncap + (north - nside) * nr;
  }
  if (north != ring) {
    theta = RSK_DPI - theta;
    start = npix - start - nr;
  }
}

// The four neighbours (ring numbering) and bilinear weights of a direction: two pixels bracketing the longitude on each
// of the two rings bracketing the latitude, linear along the ring and in colatitude between rings; inside the first
// / last ring a fictitious polar pixel = mean of the four cap pixels (HEALPix get_interpol, what
// astropy_healpix.interpolate_bilinear_lonlat evaluates).
__device__ __forceinline__ void hpx_interp_weights(const int64_t nside, const double z, const double theta,
                                                   const double phi, int64_t pix[4], double wgt[4]) {
  const double az = fabs(z);
  int64_t ir1;
  if (az <= 2.0 / 3.0) {
    ir1 = (int64_t)floor((double)nside * (2.0 - 1.5 * z));
  } else {
    const int64_t ir = (int64_t)floor((double)nside * sqrt(3.0 * (1.0 - az)));
    ir1 = z > 0.0 ? ir : 4 * nside - ir - 1;
  }
  const int64_t ir2 = ir1 + 1, npix = 12 * nside * nside;
  double theta1 = 0.0, theta2 = 0.0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t ring = h == 0 ? ir1 : ir2;
    if (ring > 0 && ring < 4 * nside) {
      int64_t start, nr;
      double th;
      bool shifted;
      hpx_ring_info(nside, ring, start, nr, th, shifted);
      const double dphi = RSK_D2PI / (double)nr;
      const double tmp = phi / dphi - (shifted ? 0.5 : 0.0);
      int64_t i1 = tmp < 0.0 ? (int64_t)(tmp - 1.0) : (int64_t)tmp;
      const double w1 = (phi - ((double)i1 + (shifted ? 0.5 : 0.0)) * dphi) / dphi;
      int64_t i2 = i1 + 1;
      if (i1 < 0) i1 += nr;
      if (i2 >= nr) i2 -= nr;
      pix[2 * h] = start + i1;
      pix[2 * h + 1] = start + i2;
      wgt[2 * h] = 1.0 - w1;
      wgt[2 * h + 1] = w1;
      if (h == 0) theta1 = th; else theta2 = th;
    }
  }
  if (ir1 == 0) {  // north of the first ring
    const double wt = theta / theta2, fac = (1.0 - wt) * 0.25;
    wgt[2] = wgt[2] * wt + fac;
    wgt[3] = wgt[3] * wt + fac;
    wgt[0] = fac;
    wgt[1] = fac;
    pix[0] = (pix[2] + 2) & 3;
    pix[1] = (pix[3] + 2) & 3;
  } else if (ir2 == 4 * nside) {  // south of the last ring
    const double wt = (theta - theta1) / (RSK_DPI - theta1), fac = wt * 0.25;
    wgt[0] = wgt[0] * (1.0 - wt) + fac;
    wgt[1] = wgt[1] * (1.0 - wt) + fac;
    wgt[2] = fac;
    wgt[3] = fac;
    pix[2] = ((pix[0] + 2) & 3) + npix - 4;
    pix[3] = ((pix[1] + 2) & 3) + npix - 4;
  } else {
    const double wt = (theta - theta1) / (theta2 - theta1);
    wgt[0] *= 1.0 - wt;
    wgt[1] *= 1.0 - wt;
    wgt[2] *= wt;
    wgt[3] *= wt;
  }
}

// One thread per pixel of the NEW frame's map: centre -> old frame -> neighbours and weights once, then every map row.
// valid[j] = 1 when all four neighbours exist in the (possibly partial) input; rows of invalid pixels are written as 0.
__global__ void __launch_bounds__(kThreads) k_healpix_interp(const int64_t nside, const int nested, const Rot9 r,
                                                             const int64_t* __restrict__ lookup, const int64_t first,
                                                             const int64_t n_in, const double* __restrict__ in,
                                                             const int64_t nrows, double* __restrict__ out,
                                                             uint8_t* __restrict__ valid) {
  const int64_t npix = 12 * nside * nside;
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (j >= npix) return;
  double z, st, phi, sp, cp;
  rsk_hpx_center(nside, nested, j, z, st, phi);
  rsk_sincos(phi, &sp, &cp);
  const double px = st * cp, py = st * sp;
  const double x = r.m[0] * px + r.m[1] * py + r.m[2] * z;
  const double y = r.m[3] * px + r.m[4] * py + r.m[5] * z;
  const double zo = r.m[6] * px + r.m[7] * py + r.m[8] * z;
  const double rxy = sqrt(x * x + y * y);
  double lon = atan2(y, x);
  if (lon < 0.0) lon += RSK_D2PI;
  const double lat = atan2(zo, rxy);
  int64_t pix[4] = {0, 0, 0, 0};
  double wgt[4] = {0.0, 0.0, 0.0, 0.0};
  hpx_interp_weights(nside, sin(lat), RSK_HALFPI - lat, lon, pix, wgt);
  int64_t pos[4];
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int64_t p = pix[k];
    if (nested) {  // ring -> nested through the pixel centre both orderings share
      double zc, sc, pc;
      rsk_hpx_center(nside, 0, p, zc, sc, pc);
      p = rsk_hpx_pixel(nside, 1, zc, sc, pc);
    }
    pos[k] = lookup ? __ldg(lookup + p) : p - first;
    ok = ok && pos[k] >= 0 && pos[k] < n_in;
  }
  if (blockIdx.y == 0) valid[j] = ok ? 1 : 0;
  for (int64_t row = blockIdx.y; row < nrows; row += gridDim.y) {
    double v = 0.0;
    if (ok) {
      const double* src = in + row * n_in;
      // products and sums rounded one by one, in neighbour order, as numpy's (values * weights).sum(axis=0)
      v = __dmul_rn(wgt[0], __ldg(src + pos[0]));
      v = __dadd_rn(v, __dmul_rn(wgt[1], __ldg(src + pos[1])));
      v = __dadd_rn(v, __dmul_rn(wgt[2], __ldg(src + pos[2])));
      v = __dadd_rn(v, __dmul_rn(wgt[3], __ldg(src + pos[3])));
    }
    out[row * npix + j] = v;
  }
}

__global__ void __launch_bounds__(kThreads) k_rise_set(const CoordSrc cs, const int64_t n, const double tan_tel,
                                                       const double buff, double* __restrict__ rise,
                                                       double* __restrict__ set) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  double lon, lat;
  frame_lonlat(cs, i, lon, lat);
  const double ha = acos(-(tan_tel * tan(lat)));  // NaN when the component never rises or never sets
  double r = lon - ha - buff, s = lon + ha + buff;
  // skymodel.py:3323-3326: one wrap each way, in this order
  if (r < 0.0) r += RSK_D2PI;
  if (s < 0.0) s += RSK_D2PI;
  if (r > RSK_D2PI) r -= RSK_D2PI;
  if (s > RSK_D2PI) s -= RSK_D2PI;
  rise[i] = r;
  set[i] = s;
}

// ---------------------------------------------------------------------------------------------------
// K1b: order-preserving compaction of the horizon mask (count -> scan of tile counts -> scatter)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t load_mask8(const uint8_t* __restrict__ m, int64_t base, int64_t n, bool aligned) {
  // bit k set <=> m[base + k] != 0
  uint32_t bits = 0;
  if (aligned && base + kCompactItems <= n) {
    const uint64_t w = __ldg(reinterpret_cast<const uint64_t*>(m + base));
#pragma unroll
    for (int k = 0; k < 8; ++k) bits |= (((w >> (8 * k)) & 0xffull) != 0 ? 1u : 0u) << k;
  } else {
#pragma unroll
    for (int k = 0; k < kCompactItems; ++k)
      if (base + k < n && m[base + k]) bits |= 1u << k;
  }
  return bits;
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  // kThreads = 256 = 8 warps
  __shared__ int s_warp[kThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int woff = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    const int c = s_warp[w];
    if (w < warp) woff += c;
    tot += c;
  }
  __syncthreads();
  *total = tot;
  return woff + incl - v;
}

__global__ void __launch_bounds__(kThreads) k_compact_count(const uint8_t* __restrict__ mask, const int64_t n,
                                                            const bool aligned, int64_t* __restrict__ tile_counts) {
  const int64_t base = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * kCompactItems;
  const int c = base < n ? __popc(load_mask8(mask, base, n, aligned)) : 0;
  int total;
  block_exclusive_scan(c, &total);
  if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

// single block: exclusive scan of the tile counts in place; grand total to tile_counts[ntiles] and *n_above
__global__ void __launch_bounds__(kThreads) k_compact_scan(int64_t* __restrict__ tile_counts, const int64_t ntiles,
                                                           int64_t* __restrict__ n_above) {
  __shared__ int64_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int64_t b0 = 0; b0 < ntiles; b0 += kThreads) {
    const int64_t j = b0 + threadIdx.x;
    const int v = j < ntiles ? (int)tile_counts[j] : 0;  // a tile holds <= 2048 items
    int total;
    const int ex = block_exclusive_scan(v, &total);
    const int64_t carry = s_carry;
    if (j < ntiles) tile_counts[j] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    tile_counts[ntiles] = s_carry;
    *n_above = s_carry;
  }
}

__global__ void __launch_bounds__(kThreads) k_compact_scatter(const uint8_t* __restrict__ mask, const int64_t n,
                                                              const bool aligned,
                                                              const int64_t* __restrict__ tile_offsets,
                                                              int64_t* __restrict__ idx) {
  const int64_t base = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * kCompactItems;
  const uint32_t bits = base < n ? load_mask8(mask, base, n, aligned) : 0u;
  int total;
  const int ex = block_exclusive_scan(__popc(bits), &total);
  int64_t o = tile_offsets[blockIdx.x] + ex;
#pragma unroll
  for (int k = 0; k < kCompactItems; ++k)
    if (bits & (1u << k)) idx[o++] = base + k;
}

// Bit j of mask[c] is set when all four Stokes parameters of component c are non-NaN at stored channel j:
// the per-component "freq_inds_use" of skymodel.py:2246-2248, for every component at once.
__global__ void __launch_bounds__(kThreads) k_valid_channels(const double* __restrict__ stokes, const int nf,
                                                             const int64_t n, uint64_t* __restrict__ mask) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  for (int w = 0; w * 64 < nf; ++w) {  // one 64-channel word at a time
    uint64_t m = 0;
    const int fend = min(nf, (w + 1) * 64);
    for (int f = w * 64; f < fend; ++f) {
      bool ok = true;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const double v = __ldg(stokes + ((int64_t)s * nf + f) * n + i);
        ok = ok && (v == v);
      }
      if (ok) m |= 1ull << (f & 63);
    }
    mask[(int64_t)w * n + i] = m;
  }
}

// Number of NaNs (KIND 0) / negative values (KIND 2) in a double array or of non-zero bytes in a byte array (KIND 1): the np.any(np.isnan(...))
// tests of skymodel.py:2157, :2196 and the size of SkyModel._polarized, without a library reduction.
template <int KIND>
__global__ void __launch_bounds__(kThreads) k_count(const void* __restrict__ x, const int64_t count,
                                                    unsigned long long* __restrict__ out) {
  int c = 0;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < count; i += (int64_t)gridDim.x * kThreads) {
    if (KIND == 0) {
      const double v = __ldg(static_cast<const double*>(x) + i);
      c += (v != v) ? 1 : 0;
    } else if (KIND == 2) {
      c += __ldg(static_cast<const double*>(x) + i) < 0.0 ? 1 : 0;
    } else {
      c += __ldg(static_cast<const uint8_t*>(x) + i) != 0 ? 1 : 0;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

// SkyModel._polarized predicate, skymodel.py:1039-1041: any_f ((Q + U) + V != 0)
__global__ void __launch_bounds__(kThreads) k_polarized(const double* __restrict__ stokes, const int64_t nf,
                                                        const int64_t n, uint8_t* __restrict__ pol) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  bool any = false;
  for (int64_t f = 0; f < nf; ++f) {
    const double q = __ldg(stokes + (1 * nf + f) * n + i);
    const double u = __ldg(stokes + (2 * nf + f) * n + i);
    const double v = __ldg(stokes + (3 * nf + f) * n + i);
    any = any || (__dadd_rn(__dadd_rn(q, u), v) != 0.0);
  }
  pol[i] = any ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------
// K2 / K3: spectral evaluation fused with (optionally) the polarisation rotation + coherency epilogue
// ---------------------------------------------------------------------------------------------------
struct EvalParams {
  // spectrum
  const double* stokes;
  int64_t nf_in, n;
  const double* alpha;
  const double* ref_freq;
  const double* breaks;
  const double* ppoly;
  int nb, degree;
  int y_in_smem;  // SUBBAND: each thread's stored spectrum is staged in dynamic shared memory
  double pivot_hz;          // SPECTRAL_INDEX: > 0 fixes the log pivot
  const int32_t* chan_map;  // dense SUBBAND on a channel subset: operator column j reads stored channel chan_map[j]
  int nf_use;               // number of operator columns (== nf_in without chan_map)
  const double* bcoef;  // banded SUBBAND: spline coefficients [4][nc][n]
  int64_t nc;
  const int32_t* first;
  const double* pband;
  const int32_t* freq_map;
  const uint8_t* polarized;
  const int32_t* group_of;  // optional filter: only components with group_of[c] == group are evaluated
  int group;
  const double2* stokes_cm;  // FULL: component-major copy [n][nf_in][4] (RskSpectrum.stokes_cm) or NULL
  // geometry (coherency epilogue only)
  CoordSrc cs;
  const double* alt_az;
  double ravg[9];
  // columns
  const int64_t* idx;
  const int64_t* cols;       // optional: thread k works on output column cols[k] (default k)
  int64_t ncols;
  const int64_t* ncols_dev;  // optional: the live column count in device memory (<= ncols)
  int64_t ntiles;
  int tile_fast;
  // frequency axis
  const double* freqs;
  int nf, fchunk, nfchunks;
  // output
  double* out;
  int64_t ld_out, col0;
  int packed;  // coherency epilogue, Hermitian-packed: 4 real planes (C00, C11, Re C01, Im C10) instead of 2x2 complex
  // apparent-coherency epilogue (EPI == 2): A = J1 C J2^H
  int beam_type;
  const double2* jones1;  // [2][2][nf][ld_jones]
  const double2* jones2;  // or NULL = jones1
  int64_t ld_jones;
};

// epilogues of k_eval
constexpr int kEpiStokes = 0;     // (4, nf, n) Stokes                      -- at_frequencies
constexpr int kEpiCoherency = 1;  // (2, 2, nf, ld) coherency, or packed   -- coherency_calc
constexpr int kEpiApparent = 2;   // (2, 2, nf, ld) J1 C J2^H               -- coherency_calc + beam Jones

// Jones matrices of one output column, [feed][basis] row-major (basis 0 = zenith-angle direction, 1 = azimuth
// direction: the basis the Az/El coherency is expressed in).
struct JonesPair {
  double2 a[4];  // antenna 1
  double2 b[4];  // antenna 2
  // Jones arrays: asynchronous global -> shared copies (cp.async) run kJonesDepth - 1 frequencies ahead of the
  // arithmetic, so the 64 / 128 bytes a thread needs per frequency are in flight without occupying registers.
  // ring[(stage * nant + ant) * 4 + q][thread]; each thread consumes only what it copied itself.
  double2* ring;
  int f0, nfc, nant, nt, depth;
  bool piped;
};
// ring depth: 96 KB per 256-thread block either way (two blocks still share an SM)
constexpr int kJonesDepth2 = 3;  // two antennas: 3 stages x 32 KB
constexpr int kJonesDepth1 = 6;  // one antenna: 6 stages x 16 KB

__device__ __forceinline__ void jones_prefetch(const EvalParams& p, const JonesPair& jp, int kk, int64_t col) {
  if (kk < jp.nfc) {
    const int stage = kk % jp.depth;
    const int64_t jplane = (int64_t)p.nf * p.ld_jones;
    const int64_t off = (int64_t)(jp.f0 + kk) * p.ld_jones + col;
    double2* dst = jp.ring + (int64_t)(stage * jp.nant) * 4 * jp.nt + threadIdx.x;
#pragma unroll
    for (int q = 0; q < 4; ++q) __pipeline_memcpy_async(dst + q * jp.nt, p.jones1 + q * jplane + off, sizeof(double2));
    if (jp.nant == 2) {
#pragma unroll
      for (int q = 0; q < 4; ++q) __pipeline_memcpy_async(dst + (4 + q) * jp.nt, p.jones2 + q * jplane + off, sizeof(double2));
    }
  }
  __pipeline_commit();
}

__device__ __forceinline__ double2 cmul(double2 x, double2 y) {
  return make_double2(x.x * y.x - x.y * y.y, x.x * y.y + x.y * y.x);
}
__device__ __forceinline__ double2 cmulc(double2 x, double2 y) {  // x * conj(y)
  return make_double2(x.x * y.x + x.y * y.y, x.y * y.x - x.x * y.y);
}
__device__ __forceinline__ double2 cadd(double2 x, double2 y) { return make_double2(x.x + y.x, x.y + y.y); }

template <int EPI>
__device__ __forceinline__ void xform(const RskRot& r, double si, double sq, double su, double sv, double* o) {
  if (EPI != kEpiStokes) {
    // R^T C R of skymodel.py:2704-2710 in Stokes terms, then utils.py:42-57 halves
    const double qr = r.c2 * sq - r.s2 * su;
    const double ur = r.s2 * sq + r.c2 * su;
    const double ir = r.n2 * si;
    o[0] = 0.5 * (ir + qr);  // C[0,0]
    o[1] = 0.5 * (ir - qr);  // C[1,1]
    o[2] = 0.5 * ur;         // Re C[0,1] = Re C[1,0]
    o[3] = 0.5 * (r.n2 * sv);  // Im C[1,0] = -Im C[0,1]
  } else {
    o[0] = si; o[1] = sq; o[2] = su; o[3] = sv;
  }
}

template <int EPI>
__device__ __forceinline__ void emit(const EvalParams& p, int f, int64_t col, const double* o, const JonesPair& jp) {
  if (EPI == kEpiApparent) {
    // A = J1 C J2^H with C = [[o0, o2 - i o3], [o2 + i o3, o1]]: the contraction
    // einsum("ab...,bc...,dc...->ad...", J, C, conj(J)) of tests/test_skymodel.py:1464-1466
    const int64_t plane = (int64_t)p.nf * p.ld_out;
    double2 j1[4], j2[4];
    if (p.beam_type == RSK_BEAM_JONES && jp.piped) {
      const int k = f - jp.f0;
      jones_prefetch(p, jp, k + jp.depth - 1, col);
      if (jp.nant == 2) __pipeline_wait_prior(kJonesDepth2 - 1);  // the copies of frequency k have landed
      else __pipeline_wait_prior(kJonesDepth1 - 1);
      const double2* src = jp.ring + (int64_t)((k % jp.depth) * jp.nant) * 4 * jp.nt + threadIdx.x;
#pragma unroll
      for (int q = 0; q < 4; ++q) j1[q] = src[q * jp.nt];
      if (jp.nant == 2) {
#pragma unroll
        for (int q = 0; q < 4; ++q) j2[q] = src[(4 + q) * jp.nt];
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) j2[q] = j1[q];
      }
    } else if (p.beam_type == RSK_BEAM_JONES) {
      // (dense subband mode, whose shared memory holds the stored spectra: direct loads)
      // __restrict__: the Jones arrays never alias the output, so the loads of the next (unrolled) frequency may
      // be issued before this one's stores -- more bytes in flight per thread for a kernel that reads 2/3 of its
      // traffic
      const int64_t jplane = (int64_t)p.nf * p.ld_jones;
      const double2* __restrict__ a = p.jones1 + (int64_t)f * p.ld_jones + col;
      const double2* __restrict__ b = (p.jones2 ? p.jones2 : p.jones1) + (int64_t)f * p.ld_jones + col;
#pragma unroll
      for (int k = 0; k < 4; ++k) j1[k] = __ldcs(a + k * jplane);
      if (p.jones2) {
#pragma unroll
        for (int k = 0; k < 4; ++k) j2[k] = __ldcs(b + k * jplane);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) j2[k] = j1[k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) { j1[k] = jp.a[k]; j2[k] = jp.b[k]; }
    }
    const double2 c00 = make_double2(o[0], 0.0), c11 = make_double2(o[1], 0.0);
    const double2 c01 = make_double2(o[2], -o[3]), c10 = make_double2(o[2], o[3]);
    const double2 t00 = cadd(cmul(j1[0], c00), cmul(j1[1], c10));
    const double2 t01 = cadd(cmul(j1[0], c01), cmul(j1[1], c11));
    const double2 t10 = cadd(cmul(j1[2], c00), cmul(j1[3], c10));
    const double2 t11 = cadd(cmul(j1[2], c01), cmul(j1[3], c11));
    double2* __restrict__ d = reinterpret_cast<double2*>(p.out) + (int64_t)f * p.ld_out + p.col0 + col;
    __stcs(d, cadd(cmulc(t00, j2[0]), cmulc(t01, j2[1])));
    __stcs(d + plane, cadd(cmulc(t00, j2[2]), cmulc(t01, j2[3])));
    __stcs(d + 2 * plane, cadd(cmulc(t10, j2[0]), cmulc(t11, j2[1])));
    __stcs(d + 3 * plane, cadd(cmulc(t10, j2[2]), cmulc(t11, j2[3])));
  } else if (EPI == kEpiCoherency && !p.packed) {
    const int64_t plane = (int64_t)p.nf * p.ld_out;
    double2* d = reinterpret_cast<double2*>(p.out) + (int64_t)f * p.ld_out + p.col0 + col;
    __stcs(d, make_double2(o[0], 0.0));
    __stcs(d + plane, make_double2(o[2], -o[3]));
    __stcs(d + 2 * plane, make_double2(o[2], o[3]));
    __stcs(d + 3 * plane, make_double2(o[1], 0.0));
  } else {
    const int64_t plane = (int64_t)p.nf * p.ld_out;
    double* d = p.out + (int64_t)f * p.ld_out + p.col0 + col;
    __stcs(d, o[0]);
    __stcs(d + plane, o[1]);
    __stcs(d + 2 * plane, o[2]);
    __stcs(d + 3 * plane, o[3]);
  }
}

// NT = threads (= columns) per block: 256, except 128 for SUBBAND so that two blocks, each holding its
// columns' stored spectra in shared memory, fit on an SM and overlap set-up with streaming.
#ifndef RSK_FULL_NT
#define RSK_FULL_NT 256  // columns per block in FULL mode (tuning: -DRSK_FULL_NT=128)
#endif
template <int MODE>
struct EvalThreads {
  static constexpr int value = (MODE == RSK_SUBBAND) ? 128 : (MODE == RSK_FULL ? RSK_FULL_NT : kThreads);
};

// Minimum resident blocks per SM asked of ptxas per mode (register cap); overridable at build time for
// tuning (-DRSK_MINB_SI=4 ...).
#ifndef RSK_MINB_SI
#define RSK_MINB_SI 3
#endif
#ifndef RSK_MINB_BANDED
#define RSK_MINB_BANDED 1
#endif
#ifndef RSK_MINB_FULL
#define RSK_MINB_FULL 1
#endif
template <int MODE, int EPI>
struct EvalMinBlocks {
  // the Jones epilogue needs ~140 registers: cap at 128 (two 256-thread blocks per SM), not at 80
  static constexpr int value = EPI == 2 ? (EvalThreads<MODE>::value == 256 ? 2 : 1)
                               : (MODE == RSK_SPECTRAL_INDEX) ? RSK_MINB_SI
                               : (MODE == 4 ? RSK_MINB_BANDED : (MODE == RSK_FULL ? RSK_MINB_FULL : 1));
};

template <int MODE, int EPI>
__global__ void __launch_bounds__(EvalThreads<MODE>::value, EvalMinBlocks<MODE, EPI>::value)
k_eval(const __grid_constant__ EvalParams p) {
  constexpr int NT = EvalThreads<MODE>::value;
  constexpr bool COH = EPI != kEpiStokes;
  constexpr int FB = EPI == kEpiApparent ? 2 : kFullBatch;  // (the Jones epilogue has no registers to spare)
  constexpr int FBP = EPI == kEpiApparent ? 2 : kFullBatchPlain;
  extern __shared__ double s_dyn[];    // SUBBAND: [4 * nf_in][NT] stored Stokes of this block's columns
  __shared__ double s_fa[kMaxFChunk];  // SPECTRAL_INDEX: ln(f / pivot); SUBBAND: local coordinate u
  __shared__ int s_iv[kMaxFChunk];     // SUBBAND: break interval; FULL: stored channel
  // Block -> (component tile, frequency chunk).  tile_fast: consecutive blocks are adjacent column
  // tiles of the same frequency chunk, so the resident blocks stream whole output rows (best DRAM page
  // locality for the stores); otherwise the chunk index runs fastest and the blocks sharing a tile run
  // back to back (inputs re-read from L2).
  int64_t tile;
  int chunk;
  if (p.tile_fast) {
    chunk = (int)(blockIdx.x / p.ntiles);
    tile = blockIdx.x - (int64_t)chunk * p.ntiles;
  } else {
    tile = blockIdx.x / p.nfchunks;
    chunk = (int)(blockIdx.x - tile * p.nfchunks);
  }
  const int64_t ncols = p.ncols_dev ? min(p.ncols, __ldg(p.ncols_dev)) : p.ncols;
  if (tile * NT >= ncols) return;  // whole block past the live columns (uniform)
  const int f0 = chunk * p.fchunk;
  const int nfc = min(p.fchunk, p.nf - f0);
  double pivot = 1.0;
  if (MODE == RSK_SPECTRAL_INDEX) pivot = p.pivot_hz > 0.0 ? p.pivot_hz : __ldg(p.freqs + p.nf / 2);
  for (int k = threadIdx.x; k < nfc; k += NT) {
    if (MODE == RSK_SPECTRAL_INDEX) {
      s_fa[k] = log(__ldg(p.freqs + f0 + k) / pivot);
    } else if (MODE == RSK_SUBBAND || MODE == kSubbandBanded) {
      const double f = __ldg(p.freqs + f0 + k);
      int lo = 0, hi = p.nb - 1;  // largest b with breaks[b] <= f, clamped to [0, nb-1]
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(p.breaks + mid) <= f) lo = mid; else hi = mid - 1;
      }
      const double b0 = __ldg(p.breaks + lo), b1 = __ldg(p.breaks + lo + 1);
      s_iv[k] = lo;
      s_fa[k] = (f - b0) / (b1 - b0);
    } else if (MODE == RSK_FULL) {
      s_iv[k] = p.freq_map ? __ldg(p.freq_map + f0 + k) : f0 + k;
    }
  }
  if (MODE != RSK_FLAT) __syncthreads();

  // FULL with a component-major copy: which layout this block reads is decided per block from how densely its
  // columns sit in the catalogue.  Measured on B200 with the batched loads below: the component-major rows win at every
  // density below 1 (0.77 / 0.91 / 0.93 / 0.94 / 0.945 of the 96 B/eval roofline at 0.05 / 0.25 / 0.5 / 0.75 / 0.95
  // against 0.21 / 0.48 / 0.69 / 0.83 / 0.90 for the gathered (4, Nf, N) reads); a block whose columns are contiguous
  // in the catalogue reads the cube's rows (0.949, the same).
  bool use_cm = false;
  if (MODE == RSK_FULL && p.stokes_cm) {
    use_cm = true;
    if (p.idx && !p.cols) {
      const int64_t k0 = tile * NT, k1 = min(k0 + NT, ncols) - 1;
      const int64_t span = __ldg(p.idx + k1) - __ldg(p.idx + k0) + 1;
      use_cm = 8 * (k1 - k0 + 1) < RSK_CM_DENSITY8 * span;
    } else if (!p.idx && !p.cols) {
      use_cm = false;  // every component is a column: plain coalesced rows
    }
  }
  const int64_t kcol = tile * NT + threadIdx.x;
  if (kcol >= ncols) return;
  const int64_t col = p.cols ? __ldg(p.cols + kcol) : kcol;
  const int64_t src = p.idx ? __ldg(p.idx + col) : col;
  // NaN-pattern launches (skymodel.py:2245-2312): this operator belongs to one group of components only.  (No
  // barrier follows, and a thread only ever consumes shared memory it filled itself, so leaving here is safe.)
  if (p.group_of && __ldg(p.group_of + src) != p.group) return;

  // per-component rotation coefficients (identity for unpolarised components, skymodel.py:2687-2688)
  RskRot rot;
  rot.n2 = 1.0; rot.c2 = 1.0; rot.s2 = 0.0;
  if (COH && p.polarized && __ldg(p.polarized + src)) {
    if (p.cs.rot) {  // coefficients already produced by K1 for this time step
      rot.n2 = __ldg(p.cs.rot + src);
      rot.c2 = __ldg(p.cs.rot + p.n + src);
      rot.s2 = __ldg(p.cs.rot + 2 * p.n + src);
    } else {
      double sl, cl, sb, cb, sa, ca, se, ce;
      frame_trig(p.cs, src, sl, cl, sb, cb);
      sincos(__ldg(p.alt_az + src), &se, &ce);
      sincos(__ldg(p.alt_az + p.n + src), &sa, &ca);
      rot = rsk_basis_rotation(p.ravg, sl, cl, sb, cb, sa, ca, se, ce);
    }
  }
  // analytic beams: one Jones matrix per column, frequency independent
  JonesPair jp;
  jp.piped = false;
  if (EPI == kEpiApparent && p.beam_type == RSK_BEAM_JONES && MODE != RSK_SUBBAND) {
    jp.ring = reinterpret_cast<double2*>(s_dyn);
    jp.f0 = f0;
    jp.nfc = nfc;
    jp.nant = p.jones2 ? 2 : 1;
    jp.depth = p.jones2 ? kJonesDepth2 : kJonesDepth1;
    jp.nt = NT;
    jp.piped = true;
    for (int kk = 0; kk < jp.depth - 1; ++kk) jones_prefetch(p, jp, kk, col);
  }
  if (EPI == kEpiApparent && p.beam_type != RSK_BEAM_JONES) {
    double ja[4] = {1.0, 0.0, 0.0, 1.0};  // RSK_BEAM_UNIFORM
    if (p.beam_type == RSK_BEAM_SHORT_DIPOLE) {
      // pyuvdata ShortDipoleBeam.efield_eval as the reference's golden test arranges it
      // (tests/test_skymodel.py:1446-1462): feeds (north, east) x basis (zenith angle, azimuth)
      double se, ce, sa, ca;
      sincos(__ldg(p.alt_az + src), &se, &ce);
      sincos(__ldg(p.alt_az + p.n + src), &sa, &ca);
      ja[0] = se * sa;  // cos(za) sin(az)
      ja[1] = ca;
      ja[2] = se * ca;  // cos(za) cos(az)
      ja[3] = -sa;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) jp.a[k] = jp.b[k] = make_double2(ja[k], 0.0);
  }
  const int64_t sstride = p.nf_in * p.n;  // stride between Stokes planes

  if (MODE == RSK_FLAT || MODE == RSK_SPECTRAL_INDEX) {
    double b[4];
    xform<EPI>(rot, __ldg(p.stokes + src), __ldg(p.stokes + sstride + src), __ldg(p.stokes + 2 * sstride + src),
               __ldg(p.stokes + 3 * sstride + src), b);
    if (MODE == RSK_FLAT) {
      // skymodel.py:2369-2372: np.repeat along the frequency axis
#pragma unroll 4
      for (int k = 0; k < nfc; ++k) emit<EPI>(p, f0 + k, col, b, jp);
    } else {
      // skymodel.py:2159-2163: stokes * (f / f_ref) ** alpha = stokes * exp(alpha (ln f/pivot - ln f_ref/pivot))
      const double alpha = __ldg(p.alpha + src);
      const double l0 = log(__ldg(p.ref_freq + src) / pivot);
#pragma unroll 4
      for (int k = 0; k < nfc; ++k) {
        const double pw = exp(alpha * (s_fa[k] - l0));
        const double o[4] = {b[0] * pw, b[1] * pw, b[2] * pw, b[3] * pw};
        emit<EPI>(p, f0 + k, col, o, jp);
      }
    }
  } else if (MODE == RSK_SUBBAND) {
    // skymodel.py:2362-2368 through the host-built piecewise-polynomial operator.  The 4 * nf_in stored
    // values of this thread's column are read from HBM once (all loads in flight together) into shared
    // memory; each change of break interval then rebuilds the 4 x 4 polynomial coefficients from there.
    // (With chan_map the operator is built on a subset of the stored channels -- the non-NaN ones of a
    // group of components, skymodel.py:2245-2312 -- and the spectra are read straight from global memory.)
    const int nfin = p.nf_use;
    double* ycol = s_dyn + threadIdx.x;
    if (p.y_in_smem) {
      const double* y = p.stokes + src;
#pragma unroll 8
      for (int q = 0; q < 4 * nfin; ++q) ycol[q * NT] = __ldg(y + (int64_t)q * p.n);
    }
    double cf[4][4];  // [power][output quantity]
    int cur = -1;
    for (int k = 0; k < nfc; ++k) {
      const int iv = s_iv[k];
      if (iv != cur) {  // uniform across the block: every thread walks the same frequencies
        cur = iv;
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int s = 0; s < 4; ++s) acc[a][s] = 0.0;
        const double* P = p.ppoly + (int64_t)iv * (p.degree + 1) * nfin;
#pragma unroll 4
        for (int j = 0; j < nfin; ++j) {
          double y0, y1, y2, y3;
          if (p.y_in_smem) {
            y0 = ycol[j * NT]; y1 = ycol[(nfin + j) * NT]; y2 = ycol[(2 * nfin + j) * NT]; y3 = ycol[(3 * nfin + j) * NT];
          } else {
            const int ch = p.chan_map ? __ldg(p.chan_map + j) : j;
            const double* y = p.stokes + (int64_t)ch * p.n + src;
            y0 = __ldg(y); y1 = __ldg(y + sstride); y2 = __ldg(y + 2 * sstride); y3 = __ldg(y + 3 * sstride);
          }
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            if (a <= p.degree) {
              const double w = __ldg(P + a * nfin + j);
              acc[a][0] = fma(w, y0, acc[a][0]);
              acc[a][1] = fma(w, y1, acc[a][1]);
              acc[a][2] = fma(w, y2, acc[a][2]);
              acc[a][3] = fma(w, y3, acc[a][3]);
            }
          }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) xform<EPI>(rot, acc[a][0], acc[a][1], acc[a][2], acc[a][3], cf[a]);
      }
      const double u = s_fa[k];
      double o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q] = fma(fma(fma(cf[3][q], u, cf[2][q]), u, cf[1][q]), u, cf[0][q]);
      emit<EPI>(p, f0 + k, col, o, jp);
    }
  } else if (MODE == kSubbandBanded) {
    // skymodel.py:2362-2368 from precomputed spline coefficients: a change of break interval costs
    // (degree+1) x 4 independent loads and a small banded product.
    const int64_t cstride = p.nc * p.n;
    const int dp1 = p.degree + 1;
    double cf[4][4];  // [power][output quantity]
    // The chunk is walked in runs of channels that share a break interval (uniform across the block): the polynomial
    // coefficients are set up once per run and the run itself is a branch-free loop the compiler can unroll, so the
    // stores of several channels are in flight together.
    for (int k0 = 0; k0 < nfc;) {
      const int iv = s_iv[k0];
      int k1 = k0 + 1;
      while (k1 < nfc && s_iv[k1] == iv) ++k1;
      {
        const double* cc = p.bcoef + (int64_t)__ldg(p.first + iv) * p.n + src;
        const double* pb = p.pband + (int64_t)iv * dp1 * dp1;
        double c[4][4];  // [m][stokes]
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
          for (int s = 0; s < 4; ++s) c[m][s] = (m < dp1) ? __ldg(cc + (int64_t)m * p.n + s * cstride) : 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          double acc[4] = {0.0, 0.0, 0.0, 0.0};
          if (a < dp1) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              if (m < dp1) {
                const double w = __ldg(pb + a * dp1 + m);
#pragma unroll
                for (int s = 0; s < 4; ++s) acc[s] = fma(w, c[m][s], acc[s]);
              }
            }
          }
          xform<EPI>(rot, acc[0], acc[1], acc[2], acc[3], cf[a]);
        }
      }
#pragma unroll 4
      for (int k = k0; k < k1; ++k) {
        const double u = s_fa[k];
        double o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) o[q] = fma(fma(fma(cf[3][q], u, cf[2][q]), u, cf[1][q]), u, cf[0][q]);
        emit<EPI>(p, f0 + k, col, o, jp);
      }
      k0 = k1;
    }
  } else if (use_cm) {
    // RSK_FULL from the component-major copy: this column's (I, Q, U, V) at one channel are one 32-byte sector, the
    // channels of a component are consecutive sectors -- exactly 32 bytes read per evaluation whatever the horizon
    // mask looks like (the gathered reads below fetch nearly every sector of a scattered mask's gaps too).  One
    // 256-bit load per evaluation.
    const double* __restrict__ row = reinterpret_cast<const double*>(p.stokes_cm) + (int64_t)src * p.nf_in * 4;
    // The output may alias nothing the kernel reads, but the compiler cannot know: left to itself it issues each
    // load after the previous channel's stores and a thread has 32 bytes in flight (1024 threads x 32 B = 32 KB per
    // SM, against ~35 KB that the HBM rate times its latency asks for).  So the loads are batched by hand.
    int k = 0;
    for (; k + FB <= nfc; k += FB) {
      double y[FB][4];
#pragma unroll
      for (int u = 0; u < FB; ++u)
        asm("ld.global.cs.v4.f64 {%0, %1, %2, %3}, [%4];"
            : "=d"(y[u][0]), "=d"(y[u][1]), "=d"(y[u][2]), "=d"(y[u][3])
            : "l"(row + 4 * (int64_t)s_iv[k + u]));
#pragma unroll
      for (int u = 0; u < FB; ++u) {
        double o[4];
        xform<EPI>(rot, y[u][0], y[u][1], y[u][2], y[u][3], o);
        emit<EPI>(p, f0 + k + u, col, o, jp);
      }
    }
    for (; k < nfc; ++k) {
      double y0, y1, y2, y3;
      asm("ld.global.cs.v4.f64 {%0, %1, %2, %3}, [%4];"
          : "=d"(y0), "=d"(y1), "=d"(y2), "=d"(y3)
          : "l"(row + 4 * (int64_t)s_iv[k]));
      double o[4];
      xform<EPI>(rot, y0, y1, y2, y3, o);
      emit<EPI>(p, f0 + k, col, o, jp);
    }
  } else {  // RSK_FULL: stored Stokes, skymodel.py:2165-2181 subset + utils.py:42-57
    // (a cp.async ring for these four 8-byte loads was measured and is slower: 0.93 -> 0.70 on the HEALPix
    // 'full' map, 0.66 -> 0.53 on the randomly gathered catalogue; 16-byte copies are what that path is good at)
    int k = 0;
    for (; k + FBP <= nfc; k += FBP) {  // loads batched by hand, as above
      double y[FBP][4];
#pragma unroll
      for (int u = 0; u < FBP; ++u) {
        const double* yp = p.stokes + (int64_t)s_iv[k + u] * p.n + src;
#pragma unroll
        for (int q = 0; q < 4; ++q) y[u][q] = __ldcs(yp + q * sstride);
      }
#pragma unroll
      for (int u = 0; u < FBP; ++u) {
        double o[4];
        xform<EPI>(rot, y[u][0], y[u][1], y[u][2], y[u][3], o);
        emit<EPI>(p, f0 + k + u, col, o, jp);
      }
    }
    for (; k < nfc; ++k) {
      const double* y = p.stokes + (int64_t)s_iv[k] * p.n + src;
      const double y0 = __ldcs(y), y1 = __ldcs(y + sstride), y2 = __ldcs(y + 2 * sstride), y3 = __ldcs(y + 3 * sstride);
      double o[4];
      xform<EPI>(rot, y0, y1, y2, y3, o);
      emit<EPI>(p, f0 + k, col, o, jp);
    }
  }
}

// stokes [4][nf][n] -> out [n][nf][4] through a shared-memory tile of kCmTile components x kCmTile channels:
// reads are coalesced along the component axis, writes along the (channel, Stokes) axis of one component.
constexpr int kCmTile = 32;
__global__ void __launch_bounds__(kThreads) k_stokes_component_major(const double* __restrict__ stokes, const int64_t nf,
                                                                     const int64_t n, double* __restrict__ out) {
  constexpr int kRow = kCmTile * 4 + 1;  // odd stride (in doubles): the transposed accesses spread over the banks
  __shared__ double tile[kCmTile * kRow];
  const int64_t c0 = (int64_t)blockIdx.x * kCmTile, f0 = (int64_t)blockIdx.y * kCmTile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // load: row r = (s, f) of the tile, lanes along components
  for (int r = warp; r < 4 * kCmTile; r += kThreads / 32) {
    const int s = r / kCmTile, f = r - s * kCmTile;
    const int64_t c = c0 + lane, ff = f0 + f;
    if (c < n && ff < nf) tile[lane * kRow + f * 4 + s] = __ldcs(stokes + (s * nf + ff) * n + c);
  }
  __syncthreads();
  // store: one component per warp trip, lanes along its kCmTile x 4 contiguous output doubles
  const int64_t nff = min((int64_t)kCmTile, nf - f0);
  for (int c = warp; c < kCmTile; c += kThreads / 32) {
    if (c0 + c >= n) break;
    double* dst = out + ((c0 + c) * nf + f0) * 4;
    for (int e = lane; e < nff * 4; e += 32) __stcs(dst + e, tile[c * kRow + e]);
  }
}

// Spline coefficients c = cmat . y per component and Stokes parameter (time independent, run once).
// blockIdx.y = Stokes parameter; cmat staged in shared memory; the spectrum of the thread's component is
// held in registers when nf_in <= NFMAX.
template <int NFMAX>
__global__ void __launch_bounds__(kThreads) k_spline_coefficients(const double* __restrict__ stokes, const int nf_in,
                                                                  const int64_t n, const double* __restrict__ cmat,
                                                                  const int nc, double* __restrict__ bcoef) {
  extern __shared__ double s_cm[];  // [nc][nf_in]
  for (int k = threadIdx.x; k < nc * nf_in; k += kThreads) s_cm[k] = __ldg(cmat + k);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const int s = blockIdx.y;
  const double* y = stokes + (int64_t)s * nf_in * n + i;
  double* out = bcoef + (int64_t)s * nc * n + i;
  if (NFMAX > 0) {
    double yr[NFMAX > 0 ? NFMAX : 1];
#pragma unroll
    for (int j = 0; j < NFMAX; ++j) yr[j] = j < nf_in ? __ldg(y + (int64_t)j * n) : 0.0;
    for (int c = 0; c < nc; ++c) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < NFMAX; ++j)
        if (j < nf_in) acc = fma(s_cm[c * nf_in + j], yr[j], acc);
      out[(int64_t)c * n] = acc;
    }
  } else {
    for (int c = 0; c < nc; ++c) {
      double acc = 0.0;
      for (int j = 0; j < nf_in; ++j) acc = fma(s_cm[c * nf_in + j], __ldg(y + (int64_t)j * n), acc);
      out[(int64_t)c * n] = acc;
    }
  }
}

// K4: Jy <-> K (sr) in place.  utils.jy_to_ksr (utils.py:105-126): astropy brightness_temperature with a
// 1 sr beam, CODATA 2018; every operation individually rounded so the factor is bit-identical to numpy's.
__device__ __forceinline__ double jy_to_ksr_factor(double nu, int to_jansky) {
  double k = __ddiv_rn(1e-26 * (299792458.0 * 299792458.0), __dmul_rn(2.0 * 1.380649e-23, __dmul_rn(nu, nu)));
  if (to_jansky) k = __ddiv_rn(1.0, k);  // skymodel.py:1559-1571
  return k;
}

constexpr int kConvChunk = 32;  // channels per block

// Block = 256 threads x 2 adjacent components x kConvChunk channels.  The per-channel factors are computed once
// per block (shared memory), the per-component ones once per thread; VEC = 128-bit accesses (n even).
template <bool VEC>
__global__ void __launch_bounds__(kThreads) k_convert_units(double* __restrict__ stokes, const int64_t nf,
                                                            const int64_t n, const double* __restrict__ fchan,
                                                            const double* __restrict__ fcomp, const int to_jansky,
                                                            const int nfchunks) {
  __shared__ double s_k[kConvChunk];
  const int64_t tile = blockIdx.x / nfchunks;
  const int chunk = (int)(blockIdx.x - tile * nfchunks);
  const int64_t f0 = (int64_t)chunk * kConvChunk;
  const int nfc = (int)min((int64_t)kConvChunk, nf - f0);
  if (!fcomp) {
    if (threadIdx.x < nfc) s_k[threadIdx.x] = jy_to_ksr_factor(__ldg(fchan + f0 + threadIdx.x), to_jansky);
    __syncthreads();
  }
  const int64_t c0 = (tile * kThreads + threadIdx.x) * 2;
  if (c0 >= n) return;
  const bool two = c0 + 1 < n;
  double k0 = 0.0, k1 = 0.0;
  if (fcomp) {
    k0 = jy_to_ksr_factor(__ldg(fcomp + c0), to_jansky);
    k1 = two ? jy_to_ksr_factor(__ldg(fcomp + c0 + 1), to_jansky) : 0.0;
  }
  const int64_t plane = nf * n;
#pragma unroll 2
  for (int k = 0; k < nfc; ++k) {
    if (!fcomp) k0 = k1 = s_k[k];
    double* row = stokes + (f0 + k) * n + c0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      double* q = row + s * plane;
      if (VEC) {
        double2 v = *reinterpret_cast<double2*>(q);
        v.x = __dmul_rn(v.x, k0);
        v.y = __dmul_rn(v.y, k1);
        *reinterpret_cast<double2*>(q) = v;
      } else {
        q[0] = __dmul_rn(q[0], k0);
        if (two) q[1] = __dmul_rn(q[1], k1);
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------------
// N4, second half: the source sum.  vis[b][a][d][f] = sum over above-horizon components s of
// (J1 C J2^H)[a][d][f][s] * exp(2 pi i f/c (uvw_b . lmn_s)) -- the apparent coherency is consumed where it is
// produced and never reaches HBM.  blockIdx.x = group of kVisF frequencies, blockIdx.y = column group,
// blockIdx.z = baseline; a thread walks its columns (stride = 256 * column groups) with kVisF x 8 accumulators
// in registers, the block reduces them once (shuffles, then shared memory) and writes one partial sum per
// (baseline, column group, frequency); k_visibility_finish adds the column groups in a fixed order, so the
// result is deterministic.
// ---------------------------------------------------------------------------------------------------
constexpr int kVisWork = 8;     // (frequencies per thread) x (baselines per thread): 64 accumulators in registers
constexpr int kVisGroups = 16;  // upper bound of column groups per (frequency group, baseline group)

struct VisParams {
  const double* stokes;
  int64_t nf_in, n;
  const double* alpha;
  const double* ref_freq;
  const int32_t* freq_map;
  const uint8_t* polarized;
  double pivot_hz;
  CoordSrc cs;
  const double* alt_az;
  double ravg[9];
  const int64_t* idx;
  int64_t ncols;
  const int64_t* ncols_dev;
  const double* freqs;
  int nf;
  int beam_type;
  const double2* jones1;
  const double2* jones2;
  int64_t ld_jones;
  const double* lmn;   // [3][n] or NULL
  const double* uvw;   // [nbl][3] metres or NULL
  int nbl;             // >= 1 (1 with uvw == NULL: the zero baseline)
  double* partial;     // [nbl][ngroups][nf][8]
  int ngroups;
};

// FPT frequencies x BPB baselines per thread (FPT * BPB = kVisWork).  The apparent coherency of a (column,
// frequency) is formed once and then multiplied by the fringe of each of the BPB baselines of this block.
template <int MODE, int FPT, int BPB>
__global__ void __launch_bounds__(kThreads) k_visibility(const __grid_constant__ VisParams p) {
  __shared__ double s_lf[FPT];    // ln(f / pivot)
  __shared__ double s_turn[FPT];  // f / c: turns of phase per metre of delay
  __shared__ double s_uvw[BPB][3];
  __shared__ double s_red[kThreads / 32][FPT * BPB * 8];
  const int f0 = blockIdx.x * FPT;
  const int nfc = min(FPT, p.nf - f0);
  const int b0 = blockIdx.z * BPB;
  const int nbc = min(BPB, p.nbl - b0);
  const double pivot = (MODE == RSK_SPECTRAL_INDEX) ? (p.pivot_hz > 0.0 ? p.pivot_hz : __ldg(p.freqs + p.nf / 2)) : 1.0;
  const bool fringe = p.uvw != nullptr && p.lmn != nullptr;
  if (threadIdx.x < FPT) {
    const double f = (p.freqs && threadIdx.x < nfc) ? __ldg(p.freqs + f0 + threadIdx.x) : 1.0;
    s_lf[threadIdx.x] = MODE == RSK_SPECTRAL_INDEX ? log(f / pivot) : 0.0;
    s_turn[threadIdx.x] = f / 299792458.0;
  }
  if (threadIdx.x < BPB * 3) {
    const int bb = threadIdx.x / 3, ax = threadIdx.x - 3 * bb;
    s_uvw[bb][ax] = (fringe && bb < nbc) ? __ldg(p.uvw + 3 * (b0 + bb) + ax) : 0.0;
  }
  __syncthreads();
  const int64_t ncols = p.ncols_dev ? min(p.ncols, __ldg(p.ncols_dev)) : p.ncols;
  double acc[FPT][BPB][8];
#pragma unroll
  for (int k = 0; k < FPT; ++k)
#pragma unroll
    for (int bb = 0; bb < BPB; ++bb)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[k][bb][c] = 0.0;
  const int64_t sstride = p.nf_in * p.n;
  for (int64_t col = (int64_t)blockIdx.y * kThreads + threadIdx.x; col < ncols; col += (int64_t)kThreads * p.ngroups) {
    const int64_t src = p.idx ? __ldg(p.idx + col) : col;
    RskRot rot;
    rot.n2 = 1.0; rot.c2 = 1.0; rot.s2 = 0.0;
    if (p.polarized && __ldg(p.polarized + src)) {
      if (p.cs.rot) {
        rot.n2 = __ldg(p.cs.rot + src);
        rot.c2 = __ldg(p.cs.rot + p.n + src);
        rot.s2 = __ldg(p.cs.rot + 2 * p.n + src);
      } else {
        double sl, cl, sb, cb, sa, ca, se, ce;
        frame_trig(p.cs, src, sl, cl, sb, cb);
        sincos(__ldg(p.alt_az + src), &se, &ce);
        sincos(__ldg(p.alt_az + p.n + src), &sa, &ca);
        rot = rsk_basis_rotation(p.ravg, sl, cl, sb, cb, sa, ca, se, ce);
      }
    }
    double ja[4] = {1.0, 0.0, 0.0, 1.0};  // analytic beams: real, frequency independent
    if (p.beam_type == RSK_BEAM_SHORT_DIPOLE) {
      double se, ce, sa, ca;
      sincos(__ldg(p.alt_az + src), &se, &ce);
      sincos(__ldg(p.alt_az + p.n + src), &sa, &ca);
      ja[0] = se * sa; ja[1] = ca; ja[2] = se * ca; ja[3] = -sa;
    }
    double delay[BPB];  // metres
#pragma unroll
    for (int bb = 0; bb < BPB; ++bb) delay[bb] = 0.0;
    if (fringe) {
      const double l = __ldg(p.lmn + src), m = __ldg(p.lmn + p.n + src), nn = __ldg(p.lmn + 2 * p.n + src);
#pragma unroll
      for (int bb = 0; bb < BPB; ++bb) delay[bb] = s_uvw[bb][0] * l + s_uvw[bb][1] * m + s_uvw[bb][2] * nn;
    }
    double b[4] = {0.0, 0.0, 0.0, 0.0}, alpha = 0.0, l0 = 0.0;
    if (MODE != RSK_FULL) {
      xform<kEpiCoherency>(rot, __ldg(p.stokes + src), __ldg(p.stokes + sstride + src), __ldg(p.stokes + 2 * sstride + src),
                           __ldg(p.stokes + 3 * sstride + src), b);
      if (MODE == RSK_SPECTRAL_INDEX) {
        alpha = __ldg(p.alpha + src);
        l0 = log(__ldg(p.ref_freq + src) / pivot);
      }
    }
#pragma unroll
    for (int k = 0; k < FPT; ++k) {
      if (k < nfc) {
        double o[4];
        if (MODE == RSK_FULL) {
          const int ch = p.freq_map ? __ldg(p.freq_map + f0 + k) : f0 + k;
          const double* y = p.stokes + (int64_t)ch * p.n + src;
          xform<kEpiCoherency>(rot, __ldcs(y), __ldcs(y + sstride), __ldcs(y + 2 * sstride), __ldcs(y + 3 * sstride), o);
        } else {
          const double pw = MODE == RSK_SPECTRAL_INDEX ? exp(alpha * (s_lf[k] - l0)) : 1.0;
#pragma unroll
          for (int q = 0; q < 4; ++q) o[q] = b[q] * pw;
        }
        double2 a4[4];
        if (p.beam_type != RSK_BEAM_JONES) {
          // real Jones matrix J = [[g0, g1], [g2, g3]] on both sides: A = J C J^T in real arithmetic
          // (C = [[o0, o2 - i o3], [o2 + i o3, o1]])
          const double g0 = ja[0], g1 = ja[1], g2 = ja[2], g3 = ja[3];
          const double r00 = g0 * g0 * o[0] + g1 * g1 * o[1] + 2.0 * g0 * g1 * o[2];
          const double r11 = g2 * g2 * o[0] + g3 * g3 * o[1] + 2.0 * g2 * g3 * o[2];
          const double r01 = g0 * g2 * o[0] + g1 * g3 * o[1] + (g0 * g3 + g1 * g2) * o[2];
          const double i01 = (g1 * g2 - g0 * g3) * o[3];  // Im A01 = -Im A10
          a4[0] = make_double2(r00, 0.0);
          a4[1] = make_double2(r01, i01);
          a4[2] = make_double2(r01, -i01);
          a4[3] = make_double2(r11, 0.0);
        } else {
          double2 j1[4], j2[4];
          const int64_t jplane = (int64_t)p.nf * p.ld_jones;
          const double2* a = p.jones1 + (int64_t)(f0 + k) * p.ld_jones + col;
#pragma unroll
          for (int q = 0; q < 4; ++q) j1[q] = __ldcs(a + q * jplane);
          if (p.jones2) {
            const double2* bb2 = p.jones2 + (int64_t)(f0 + k) * p.ld_jones + col;
#pragma unroll
            for (int q = 0; q < 4; ++q) j2[q] = __ldcs(bb2 + q * jplane);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) j2[q] = j1[q];
          }
          const double2 c00 = make_double2(o[0], 0.0), c11 = make_double2(o[1], 0.0);
          const double2 c01 = make_double2(o[2], -o[3]), c10 = make_double2(o[2], o[3]);
          const double2 t00 = cadd(cmul(j1[0], c00), cmul(j1[1], c10));
          const double2 t01 = cadd(cmul(j1[0], c01), cmul(j1[1], c11));
          const double2 t10 = cadd(cmul(j1[2], c00), cmul(j1[3], c10));
          const double2 t11 = cadd(cmul(j1[2], c01), cmul(j1[3], c11));
          a4[0] = cadd(cmulc(t00, j2[0]), cmulc(t01, j2[1]));
          a4[1] = cadd(cmulc(t00, j2[2]), cmulc(t01, j2[3]));
          a4[2] = cadd(cmulc(t10, j2[0]), cmulc(t11, j2[1]));
          a4[3] = cadd(cmulc(t10, j2[2]), cmulc(t11, j2[3]));
        }
#pragma unroll
        for (int bb = 0; bb < BPB; ++bb) {
          double2 ph = make_double2(1.0, 0.0);
          if (fringe) {
            const double turns = s_turn[k] * delay[bb];
            rsk_sincos(RSK_D2PI * (turns - rint(turns)), &ph.y, &ph.x);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double2 v = fringe ? cmul(a4[q], ph) : a4[q];
            acc[k][bb][2 * q] += v.x;
            acc[k][bb][2 * q + 1] += v.y;
          }
        }
      }
    }
  }
  // block reduction: lanes, then warps
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < FPT; ++k)
#pragma unroll
    for (int bb = 0; bb < BPB; ++bb)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        double v = acc[k][bb][c];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
        if (lane == 0) s_red[warp][(k * BPB + bb) * 8 + c] = v;
      }
  __syncthreads();
  if (threadIdx.x < FPT * BPB * 8) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) v += s_red[w][threadIdx.x];
    const int c = threadIdx.x & 7, kb = threadIdx.x >> 3;
    const int k = kb / BPB, bb = kb - k * BPB;
    if (k < nfc && bb < nbc)
      p.partial[(((int64_t)(b0 + bb) * p.ngroups + blockIdx.y) * p.nf + f0 + k) * 8 + c] = v;
  }
}

// vis[b][plane][f] (complex) = sum over column groups of partial[b][g][f][plane], groups added in order
__global__ void __launch_bounds__(kThreads) k_visibility_finish(const double* __restrict__ partial, const int ngroups,
                                                                const int nf, const int64_t total, double* __restrict__ vis) {
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;  // over nbl * nf * 8
  if (e >= total) return;
  const int c = (int)(e & 7);
  const int64_t bf = e >> 3;
  const int64_t b = bf / nf;
  const int f = (int)(bf - b * nf);
  double v = 0.0;
  for (int g = 0; g < ngroups; ++g) v += partial[((b * ngroups + g) * nf + f) * 8 + c];
  vis[((b * 4 + (c >> 1)) * nf + f) * 2 + (c & 1)] = v;
}

// ---------------------------------------------------------------------------------------------------
// host-side launch helpers
// ---------------------------------------------------------------------------------------------------
// Launch-shape overrides for tuning runs (read per call; no state is kept).
inline int tuning_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

inline int launch_status() {
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? RSK_OK : (int)e;
}

inline int check_coords(const RskCoords* c, CoordSrc* out) {
  if (!c) return RSK_ERR_NULL;
  if (c->nside < 0) return RSK_ERR_NSIDE;
  if (c->nside > 0) {
    if (c->nside > (1ll << 29)) return RSK_ERR_NSIDE;
    if (c->nested && (c->nside & (c->nside - 1)) != 0) return RSK_ERR_NSIDE;
  } else if (!c->lon || !c->lat) {
    return RSK_ERR_NULL;
  }
  out->lon = c->lon;
  out->lat = c->lat;
  out->nside = c->nside;
  out->nested = c->nested;
  out->hpx_inds = c->hpx_inds;
  out->hpx_first = c->hpx_first;
  out->rot = c->rot;
  return RSK_OK;
}

inline unsigned blocks_for(int64_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

int fill_eval_params(const RskSpectrum* spec, const double* freqs, int64_t nf, int64_t ncols, EvalParams* p) {
  if (!spec || !spec->stokes) return RSK_ERR_NULL;
  if (nf < 0 || ncols < 0 || spec->n < 0 || spec->nf_in < 1 || nf > INT32_MAX) return RSK_ERR_SIZE;
  if (spec->mode < RSK_FLAT || spec->mode > RSK_FULL) return RSK_ERR_MODE;
  if ((spec->mode == RSK_SPECTRAL_INDEX || spec->mode == RSK_SUBBAND) && !freqs) return RSK_ERR_NULL;
  if (spec->mode == RSK_SPECTRAL_INDEX && (!spec->spectral_index || !spec->ref_freq)) return RSK_ERR_NULL;
  if (spec->mode == RSK_SUBBAND) {
    if (!spec->breaks) return RSK_ERR_NULL;
    if (spec->bcoef ? (!spec->first || !spec->pband || spec->nc < 1) : !spec->ppoly) return RSK_ERR_NULL;
    if (spec->nb < 1 || spec->nb > INT32_MAX) return RSK_ERR_SIZE;
    if (spec->degree < 0 || spec->degree > 3) return RSK_ERR_UNSUPPORTED;
    if (spec->chan_map && (spec->bcoef || spec->nf_use < 1 || spec->nf_use > spec->nf_in)) return RSK_ERR_SIZE;
  }
  if (spec->mode == RSK_FULL && !spec->freq_map && nf != spec->nf_in) return RSK_ERR_SIZE;
  if ((spec->mode == RSK_FLAT || spec->mode == RSK_SPECTRAL_INDEX) && spec->nf_in != 1) return RSK_ERR_SIZE;
  p->stokes = spec->stokes;
  p->nf_in = spec->nf_in;
  p->n = spec->n;
  p->alpha = spec->spectral_index;
  p->ref_freq = spec->ref_freq;
  p->breaks = spec->breaks;
  p->ppoly = spec->ppoly;
  p->nb = (int)spec->nb;
  p->degree = spec->degree;
  p->freq_map = spec->freq_map;
  p->polarized = spec->polarized;
  p->group_of = spec->group_of;
  p->group = spec->group;
  p->stokes_cm = spec->mode == RSK_FULL ? reinterpret_cast<const double2*>(spec->stokes_cm) : nullptr;
  p->pivot_hz = spec->pivot_hz;
  p->chan_map = spec->mode == RSK_SUBBAND ? spec->chan_map : nullptr;
  p->nf_use = (int)((spec->mode == RSK_SUBBAND && spec->chan_map) ? spec->nf_use : spec->nf_in);
  p->bcoef = spec->mode == RSK_SUBBAND ? spec->bcoef : nullptr;
  p->nc = spec->nc;
  p->first = spec->first;
  p->pband = spec->pband;
  p->freqs = freqs;
  p->nf = (int)nf;
  p->ncols = ncols;
  // Frequency chunking (defaults measured on B200, see the constants at the top of this file).
  const bool dense_subband = spec->mode == RSK_SUBBAND && !spec->bcoef;
  const int nt = dense_subband ? EvalThreads<RSK_SUBBAND>::value
                               : (spec->mode == RSK_FULL ? EvalThreads<RSK_FULL>::value : kThreads);
  const int64_t tiles = (ncols + nt - 1) / nt;
  // SUBBAND: stage each column's 4 * nf_in stored values in shared memory when they fit
  p->y_in_smem = (dense_subband && !spec->chan_map && 4 * spec->nf_in * nt * (int64_t)sizeof(double) <= kMaxDynSmem) ? 1 : 0;
  // (stored Stokes read from the component-major copy: 32-frequency chunks measured best, 0.86 against 0.84 with 64
  // and 0.81 with 128 of the 96 B/eval roofline, profiles/r02_full_path_fchunk_sweep.txt)
  int fchunk = tuning_int("RSK_K2_FCHUNK", dense_subband ? kSubbandFChunk : (p->stokes_cm ? kFullCmFChunk : kDefaultFChunk));
  if (fchunk < 1) fchunk = 1;
  if (fchunk > kMaxFChunk) fchunk = kMaxFChunk;
  if (fchunk > nf) fchunk = (int)(nf > 0 ? nf : 1);
  p->fchunk = fchunk;
  p->nfchunks = (int)((nf + fchunk - 1) / fchunk);
  if (p->nfchunks < 1) p->nfchunks = 1;
  p->ntiles = tiles > 0 ? tiles : 1;
  p->tile_fast = tuning_int("RSK_K2_TILEFAST", kDefaultTileFast);
  return RSK_OK;
}

template <int MODE, int EPI>
int launch_one(const EvalParams& p, dim3 grid, int smem, cudaStream_t st) {
  constexpr int NT = EvalThreads<MODE>::value;
  {
    int need = 0;
    if (EPI == kEpiApparent && p.beam_type == RSK_BEAM_JONES && MODE != RSK_SUBBAND)
      need += (p.jones2 ? 2 * kJonesDepth2 : kJonesDepth1) * 4 * NT * (int)sizeof(double2);  // the Jones cp.async ring
    if (need > smem) smem = need;
  }
  if (MODE == RSK_SUBBAND && p.y_in_smem) {
    const int need = (int)(4 * p.nf_in * NT * sizeof(double));
    if (need > smem) smem = need;
  }
  if (smem > 40 * 1024) {  // dynamic + the kernel's ~3 KB of static shared memory must stay under the 48 KB default
    const cudaError_t e = cudaFuncSetAttribute(k_eval<MODE, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
  }
  k_eval<MODE, EPI><<<grid, NT, smem, st>>>(p);
  return launch_status();
}

template <int EPI>
int launch_eval(const EvalParams& p, int mode, cudaStream_t st) {
  if (p.ncols == 0 || p.nf == 0) return RSK_OK;
  const int64_t nblocks = p.ntiles * p.nfchunks;
  if (nblocks > 0x7fffffffll) return RSK_ERR_SIZE;
  const dim3 grid((unsigned)nblocks);
  // The kernels use no dynamic shared memory; requesting some only caps how many blocks share an SM.
  int smem = tuning_int("RSK_K2_SMEM", kDefaultOccSmem);
  if (smem < 0) smem = 0;
  if (smem > kMaxDynSmem) smem = kMaxDynSmem;
  switch (mode) {
    case RSK_FLAT: return launch_one<RSK_FLAT, EPI>(p, grid, smem, st);
    case RSK_SPECTRAL_INDEX: return launch_one<RSK_SPECTRAL_INDEX, EPI>(p, grid, smem, st);
    case RSK_SUBBAND:
      if (p.bcoef) return launch_one<kSubbandBanded, EPI>(p, grid, smem, st);
      return launch_one<RSK_SUBBAND, EPI>(p, grid, smem, st);
    case RSK_FULL: return launch_one<RSK_FULL, EPI>(p, grid, smem, st);
    default: return RSK_ERR_MODE;
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" {

int rsk_positions(const RskCoords* coords, int64_t n, const double* frame2icrs, const double* astrom,
                  const double* ravg, double* alt_az, double* pos_lmn, uint8_t* above, void* stream) {
  CoordSrc cs;
  const int rc = check_coords(coords, &cs);
  if (rc != RSK_OK) return rc;
  if (!astrom || !alt_az) return RSK_ERR_NULL;
  if (n < 0) return RSK_ERR_SIZE;
  if (n == 0) return RSK_OK;
  RskCtx ctx;
  rsk_pack_context(astrom, frame2icrs, reinterpret_cast<double*>(&ctx));
  Ravg rv = {};
  if (ravg && cs.rot) {
    for (int i = 0; i < 9; ++i) rv.m[i] = ravg[i];
    rv.valid = 1;
  }
  k_positions<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(cs, n, ctx, rv, alt_az, pos_lmn, above);
  return launch_status();
}

int rsk_positions_step(const RskCoords* coords, int64_t n, const double* step, double* alt_az, double* pos_lmn,
                       uint8_t* above, void* stream) {
  CoordSrc cs;
  const int rc = check_coords(coords, &cs);
  if (rc != RSK_OK) return rc;
  if (!step || !alt_az) return RSK_ERR_NULL;
  if (n < 0) return RSK_ERR_SIZE;
  if (n == 0) return RSK_OK;
  k_positions_ctxdev<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(cs, n, step, alt_az, pos_lmn, above);
  return launch_status();
}

int rsk_positions_sweep(const RskCoords* coords, int64_t n, const double* ctx, int64_t T, double* alt_az,
                        double* pos_lmn, uint8_t* above, void* stream) {
  CoordSrc cs;
  const int rc = check_coords(coords, &cs);
  if (rc != RSK_OK) return rc;
  if (!ctx || !alt_az) return RSK_ERR_NULL;
  if (n < 0 || T < 0) return RSK_ERR_SIZE;
  if (n == 0 || T == 0) return RSK_OK;
  // components per thread (independent chains per thread): 1 measured best once the chain was trimmed to ~90 fp64
  // instructions (profiles/r01_sweep_tuning.log); RSK_SWEEP_CPT = 1, 2 or 4 overrides it for tuning runs
  int cpt = tuning_int("RSK_SWEEP_CPT", 1);
  if (cpt != 1 && cpt != 2 && cpt != 4) cpt = 1;
  const int64_t ntchunks = (T + kSweepTimes - 1) / kSweepTimes;
  const int64_t ntiles = (n + cpt * kThreads - 1) / (cpt * kThreads);
  const int64_t nblocks = ntiles * ntchunks;
  if (nblocks > 0x7fffffffll) return RSK_ERR_SIZE;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)nblocks;
  const int ntch = (int)ntchunks;
#define RSK_SWEEP(C, L) k_positions_sweep<C, L><<<grid, kThreads, 0, st>>>(cs, n, ctx, T, ntch, alt_az, pos_lmn, above)
  if (pos_lmn) {
    if (cpt == 1) RSK_SWEEP(1, true); else if (cpt == 2) RSK_SWEEP(2, true); else RSK_SWEEP(4, true);
  } else {
    if (cpt == 1) RSK_SWEEP(1, false); else if (cpt == 2) RSK_SWEEP(2, false); else RSK_SWEEP(4, false);
  }
#undef RSK_SWEEP
  return launch_status();
}

int rsk_healpix_lonlat(int64_t nside, int nested, const int64_t* hpx_inds, int64_t hpx_first, int64_t n, double* lon,
                       double* lat, void* stream) {
  RskCoords c = {nullptr, nullptr, nside, nested, 0, hpx_inds, hpx_first, nullptr};
  CoordSrc cs;
  if (nside <= 0) return RSK_ERR_NSIDE;
  const int rc = check_coords(&c, &cs);
  if (rc != RSK_OK) return rc;
  if (!lon || !lat) return RSK_ERR_NULL;
  if (n < 0) return RSK_ERR_SIZE;
  if (n == 0) return RSK_OK;
  k_healpix_lonlat<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(nside, nested, hpx_inds, hpx_first, n, lon, lat);
  return launch_status();
}

int rsk_lonlat_to_healpix(int64_t nside, int nested, const double* lon, const double* lat, int64_t n, int64_t* pix,
                          void* stream) {
  if (nside <= 0 || nside > (1ll << 29)) return RSK_ERR_NSIDE;
  if (nested && (nside & (nside - 1)) != 0) return RSK_ERR_NSIDE;
  if (n < 0) return RSK_ERR_SIZE;
  if (n == 0) return RSK_OK;
  if (!lon || !lat || !pix) return RSK_ERR_NULL;
  k_lonlat_to_healpix<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(nside, nested, lon, lat, n, pix);
  return launch_status();
}

int rsk_scale(double* x, int64_t count, double factor, int divide, void* stream) {
  if (count < 0) return RSK_ERR_SIZE;
  if (count == 0) return RSK_OK;
  if (!x) return RSK_ERR_NULL;
  if ((count + kThreads - 1) / kThreads > 0x7fffffffll) return RSK_ERR_SIZE;
  k_scale<<<blocks_for(count), kThreads, 0, (cudaStream_t)stream>>>(x, count, factor, divide);
  return launch_status();
}

int rsk_select_mask(const RskCoords* coords, const double* stokes, int64_t nf, int64_t n, const int64_t* comp_inds,
                    int64_t m, const RskSelect* sel, uint8_t* keep, void* stream) {
  if (!sel) return RSK_ERR_NULL;
  if (nf < 0 || n < 0 || m < 0) return RSK_ERR_SIZE;
  if (sel->non_nan < 0 || sel->non_nan > 2) return RSK_ERR_MODE;
  if (m == 0) return RSK_OK;
  if (!keep) return RSK_ERR_NULL;
  const bool need_coords = sel->use_lat || sel->use_lon || sel->cut_nonrising;
  const bool need_stokes = sel->use_min || sel->use_max || sel->non_negative || sel->non_nan;
  CoordSrc cs = {};
  if (need_coords) {
    const int rc = check_coords(coords, &cs);
    if (rc != RSK_OK) return rc;
  }
  if (need_stokes && !stokes) return RSK_ERR_NULL;
  if (sel->use_lat && !(sel->lat_lo <= sel->lat_hi)) return RSK_ERR_SIZE;
  k_select_mask<<<blocks_for(m), kThreads, 0, (cudaStream_t)stream>>>(cs, stokes, nf, n, comp_inds, m, *sel, keep);
  return launch_status();
}

int rsk_gather_rows(const void* in, int64_t nrows, int64_t n_in, const int64_t* idx, int64_t n_out, int elem_bytes,
                    void* out, void* stream) {
  if (nrows < 0 || n_in < 0 || n_out < 0) return RSK_ERR_SIZE;
  if (elem_bytes != 1 && elem_bytes != 8) return RSK_ERR_MODE;
  if (nrows == 0 || n_out == 0) return RSK_OK;
  if (!in || !idx || !out) return RSK_ERR_NULL;
  if ((n_out + kThreads - 1) / kThreads > 0x7fffffffll) return RSK_ERR_SIZE;
  constexpr int kRpt = 4;
  const int64_t row_groups = (nrows + kRpt - 1) / kRpt;
  const dim3 grid(blocks_for(n_out), (unsigned)std::min<int64_t>(row_groups, 65535));
  cudaStream_t st = (cudaStream_t)stream;
  if (elem_bytes == 8)
    k_gather_rows<double, kRpt><<<grid, kThreads, 0, st>>>((const double*)in, nrows, n_in, idx, n_out, (double*)out);
  else
    k_gather_rows<uint8_t, kRpt><<<grid, kThreads, 0, st>>>((const uint8_t*)in, nrows, n_in, idx, n_out, (uint8_t*)out);
  return launch_status();
}

int rsk_concat_rows(const void* a, int64_t n_a, const void* b, int64_t n_b, int64_t nrows, int elem_bytes, void* out,
                    void* stream) {
  if (nrows < 0 || n_a < 0 || n_b < 0) return RSK_ERR_SIZE;
  if (elem_bytes != 1 && elem_bytes != 8) return RSK_ERR_MODE;
  if (nrows == 0 || n_a + n_b == 0) return RSK_OK;
  if (!out) return RSK_ERR_NULL;
  if ((n_a + n_b + kThreads - 1) / kThreads > 0x7fffffffll) return RSK_ERR_SIZE;
  const dim3 grid(blocks_for(n_a + n_b), (unsigned)std::min<int64_t>(nrows, 65535));
  cudaStream_t st = (cudaStream_t)stream;
  if (elem_bytes == 8)
    k_concat_rows<double><<<grid, kThreads, 0, st>>>((const double*)a, n_a, (const double*)b, n_b, nrows, (double)NAN,
                                                     (double*)out);
  else
    k_concat_rows<uint8_t><<<grid, kThreads, 0, st>>>((const uint8_t*)a, n_a, (const uint8_t*)b, n_b, nrows, (uint8_t)0,
                                                      (uint8_t*)out);
  return launch_status();
}

int rsk_rotate_lonlat(const double* lon, const double* lat, int64_t n, const double* rot, double* lon_out,
                      double* lat_out, void* stream) {
  if (n < 0) return RSK_ERR_SIZE;
  if (n == 0) return RSK_OK;
  if (!lon || !lat || !rot || !lon_out || !lat_out) return RSK_ERR_NULL;
  Rot9 r;
  for (int i = 0; i < 9; ++i) r.m[i] = rot[i];
  k_rotate_lonlat<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(lon, lat, n, r, lon_out, lat_out);
  return launch_status();
}

int rsk_healpix_interp(int64_t nside, int nested, const double* rot_new_to_old, const int64_t* lookup, int64_t hpx_first,
                       int64_t n_in, const double* in, int64_t nrows, double* out, uint8_t* valid, void* stream) {
  if (nside <= 0 || nside > (1ll << 29)) return RSK_ERR_NSIDE;
  if (nested && (nside & (nside - 1)) != 0) return RSK_ERR_NSIDE;
  if (n_in < 0 || nrows < 0) return RSK_ERR_SIZE;
  if (!rot_new_to_old || !valid || (nrows > 0 && (!in || !out))) return RSK_ERR_NULL;
  const int64_t npix = 12 * nside * nside;
  if ((npix + kThreads - 1) / kThreads > 0x7fffffffll) return RSK_ERR_SIZE;
  Rot9 r;
  for (int i = 0; i < 9; ++i) r.m[i] = rot_new_to_old[i];
  const dim3 grid(blocks_for(npix), (unsigned)std::max<int64_t>(1, std::min<int64_t>(nrows, 64)));
  k_healpix_interp<<<grid, kThreads, 0, (cudaStream_t)stream>>>(nside, nested, r, lookup, hpx_first, n_in, in, nrows, out,
                                                                valid);
  return launch_status();
}

int rsk_rise_set_lsts(const RskCoords* coords, int64_t n, double tel_lat, double horizon_buffer, double* rise_lst,
                      double* set_lst, void* stream) {
  CoordSrc cs;
  const int rc = check_coords(coords, &cs);
  if (rc != RSK_OK) return rc;
  if (n < 0) return RSK_ERR_SIZE;
  if (n == 0) return RSK_OK;
  if (!rise_lst || !set_lst) return RSK_ERR_NULL;
  k_rise_set<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(cs, n, tan(tel_lat), horizon_buffer, rise_lst, set_lst);
  return launch_status();
}

int rsk_segment_sum(const double* in, int64_t nrows, int64_t n_in, const int64_t* order, const int64_t* start,
                    int64_t nseg, int quadrature, double* out, void* stream) {
  if (nrows < 0 || n_in < 0 || nseg < 0 || nrows > 65535) return RSK_ERR_SIZE;
  if (nrows == 0 || nseg == 0) return RSK_OK;
  if (!in || !order || !start || !out) return RSK_ERR_NULL;
  const dim3 grid(blocks_for(nseg), (unsigned)nrows);
  k_segment_sum<<<grid, kThreads, 0, (cudaStream_t)stream>>>(in, nrows, n_in, order, start, nseg, quadrature, out);
  return launch_status();
}

int rsk_compact_index(const uint8_t* above, int64_t n, int64_t* idx, int64_t* n_above, void* work, size_t work_bytes,
                      void* stream) {
  if (!n_above) return RSK_ERR_NULL;
  if (n < 0) return RSK_ERR_SIZE;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    const cudaError_t e = cudaMemsetAsync(n_above, 0, sizeof(int64_t), st);
    return e == cudaSuccess ? RSK_OK : (int)e;
  }
  if (!above || !idx) return RSK_ERR_NULL;
  if (!work || work_bytes < rsk_workspace_bytes(RSK_OP_COMPACT, n, 0)) return RSK_ERR_WORKSPACE;
  const int64_t ntiles = (n + kCompactTile - 1) / kCompactTile;
  if (ntiles > 0x7fffffffll) return RSK_ERR_SIZE;
  const bool aligned = (reinterpret_cast<uintptr_t>(above) & 7u) == 0;
  int64_t* tiles = static_cast<int64_t*>(work);
  k_compact_count<<<(unsigned)ntiles, kThreads, 0, st>>>(above, n, aligned, tiles);
  k_compact_scan<<<1, kThreads, 0, st>>>(tiles, ntiles, n_above);
  k_compact_scatter<<<(unsigned)ntiles, kThreads, 0, st>>>(above, n, aligned, tiles, idx);
  return launch_status();
}

int rsk_valid_channels(const double* stokes, int64_t nf_in, int64_t n, uint64_t* mask, void* stream) {
  if (n < 0 || nf_in < 0 || nf_in > INT32_MAX) return RSK_ERR_SIZE;
  if (n == 0 || nf_in == 0) return RSK_OK;
  if (!stokes || !mask) return RSK_ERR_NULL;
  k_valid_channels<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(stokes, (int)nf_in, n, mask);
  return launch_status();
}

int rsk_count(const void* x, int64_t count, int kind, int64_t* out, void* stream) {
  if (!out) return RSK_ERR_NULL;
  if (count < 0) return RSK_ERR_SIZE;
  if (kind != RSK_COUNT_NAN_F64 && kind != RSK_COUNT_NONZERO_U8 && kind != RSK_COUNT_NEGATIVE_F64) return RSK_ERR_MODE;
  cudaStream_t st = (cudaStream_t)stream;
  const cudaError_t e = cudaMemsetAsync(out, 0, sizeof(int64_t), st);
  if (e != cudaSuccess) return (int)e;
  if (count == 0) return RSK_OK;
  if (!x) return RSK_ERR_NULL;
  const unsigned grid = (unsigned)std::min<int64_t>((count + kThreads - 1) / kThreads, 148 * 16);
  if (kind == RSK_COUNT_NAN_F64)
    k_count<0><<<grid, kThreads, 0, st>>>(x, count, reinterpret_cast<unsigned long long*>(out));
  else if (kind == RSK_COUNT_NEGATIVE_F64)
    k_count<2><<<grid, kThreads, 0, st>>>(x, count, reinterpret_cast<unsigned long long*>(out));
  else
    k_count<1><<<grid, kThreads, 0, st>>>(x, count, reinterpret_cast<unsigned long long*>(out));
  return launch_status();
}

int rsk_polarized_mask(const double* stokes, int64_t nf_in, int64_t n, uint8_t* pol, void* stream) {
  if (n < 0 || nf_in < 0) return RSK_ERR_SIZE;
  if (n == 0) return RSK_OK;
  if (!stokes || !pol) return RSK_ERR_NULL;
  k_polarized<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(stokes, nf_in, n, pol);
  return launch_status();
}

static int coherency_impl(const RskSpectrum* spec, const RskCoords* coords, const double* alt_az, const double* ravg,
                          const int64_t* idx, const int64_t* cols, int64_t ncols, const int64_t* ncols_dev,
                          const double* freqs, int64_t nf, double* out, int64_t ld_out, int64_t col0, int packed,
                          void* stream, const RskBeam* beam = nullptr) {
  EvalParams p = {};
  p.packed = packed;
  if (beam) {
    if (beam->type < RSK_BEAM_JONES || beam->type > RSK_BEAM_UNIFORM) return RSK_ERR_MODE;
    p.beam_type = beam->type;
    if (beam->type == RSK_BEAM_JONES) {
      if (ncols > 0 && nf > 0 && !beam->jones1) return RSK_ERR_NULL;
      if (cols ? beam->ld_jones < 1 : beam->ld_jones < ncols) return RSK_ERR_LD;
      p.jones1 = reinterpret_cast<const double2*>(beam->jones1);
      p.jones2 = reinterpret_cast<const double2*>(beam->jones2);
      p.ld_jones = beam->ld_jones;
    } else if (beam->type == RSK_BEAM_SHORT_DIPOLE && !alt_az) {
      return RSK_ERR_NULL;
    }
  }
  int rc = fill_eval_params(spec, freqs, nf, ncols, &p);
  if (rc != RSK_OK) return rc;
  if (ncols > 0 && nf > 0 && !out) return RSK_ERR_NULL;
  if (col0 < 0 || (!cols && ld_out < col0 + ncols)) return RSK_ERR_LD;
  if (!idx && !cols && ncols > spec->n) return RSK_ERR_SIZE;
  if (spec->polarized) {  // rotation needed: geometry must be present
    if (!coords) return RSK_ERR_NULL;
    if (coords->rot) {
      p.cs.rot = coords->rot;  // precomputed by rsk_positions: nothing else is read
    } else {
      rc = check_coords(coords, &p.cs);
      if (rc != RSK_OK) return rc;
      if (!alt_az || !ravg) return RSK_ERR_NULL;
      for (int i = 0; i < 9; ++i) p.ravg[i] = ravg[i];
    }
  }
  p.alt_az = alt_az;
  p.idx = idx;
  p.cols = cols;
  p.ncols_dev = ncols_dev;
  p.out = out;
  p.ld_out = ld_out;
  p.col0 = col0;
  if (beam) return launch_eval<kEpiApparent>(p, spec->mode, (cudaStream_t)stream);
  return launch_eval<kEpiCoherency>(p, spec->mode, (cudaStream_t)stream);
}

int rsk_spline_coefficients(const double* stokes, int64_t nf_in, int64_t n, const double* cmat, int64_t nc,
                            double* bcoef, void* stream) {
  if (nf_in < 1 || n < 0 || nc < 1) return RSK_ERR_SIZE;
  if (n == 0) return RSK_OK;
  if (!stokes || !cmat || !bcoef) return RSK_ERR_NULL;
  const size_t smem = (size_t)nc * nf_in * sizeof(double);
  if (smem > (size_t)kMaxDynSmem) return RSK_ERR_UNSUPPORTED;
  const dim3 grid(blocks_for(n), 4);
  cudaStream_t st = (cudaStream_t)stream;
  if (nf_in <= 32) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_spline_coefficients<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_spline_coefficients<32><<<grid, kThreads, smem, st>>>(stokes, (int)nf_in, n, cmat, (int)nc, bcoef);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_spline_coefficients<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_spline_coefficients<0><<<grid, kThreads, smem, st>>>(stokes, (int)nf_in, n, cmat, (int)nc, bcoef);
  }
  return launch_status();
}

int rsk_coherency(const RskSpectrum* spec, const RskCoords* coords, const double* alt_az, const double* ravg,
                  const int64_t* idx, const int64_t* cols, int64_t ncols, const int64_t* ncols_dev, const double* freqs,
                  int64_t nf, double* out, int64_t ld_out, int64_t col0, void* stream) {
  return coherency_impl(spec, coords, alt_az, ravg, idx, cols, ncols, ncols_dev, freqs, nf, out, ld_out, col0, 0, stream);
}

int rsk_coherency_packed(const RskSpectrum* spec, const RskCoords* coords, const double* alt_az, const double* ravg,
                         const int64_t* idx, const int64_t* cols, int64_t ncols, const int64_t* ncols_dev,
                         const double* freqs, int64_t nf, double* out, int64_t ld_out, int64_t col0, void* stream) {
  return coherency_impl(spec, coords, alt_az, ravg, idx, cols, ncols, ncols_dev, freqs, nf, out, ld_out, col0, 1, stream);
}

int rsk_apparent_coherency(const RskSpectrum* spec, const RskCoords* coords, const double* alt_az, const double* ravg,
                            const int64_t* idx, const int64_t* cols, int64_t ncols, const int64_t* ncols_dev,
                            const double* freqs, int64_t nf, const RskBeam* beam, double* out, int64_t ld_out,
                            int64_t col0, void* stream) {
  if (!beam) return RSK_ERR_NULL;
  return coherency_impl(spec, coords, alt_az, ravg, idx, cols, ncols, ncols_dev, freqs, nf, out, ld_out, col0, 0, stream,
                        beam);
}

static int visibility_groups(int64_t ncols) {
  int g = (int)((ncols + 16 * kThreads - 1) / (16 * kThreads));  // at least ~16 columns per thread
  return g < 1 ? 1 : (g > kVisGroups ? kVisGroups : g);
}

int rsk_visibility(const RskSpectrum* spec, const RskCoords* coords, const double* alt_az, const double* ravg,
                   const int64_t* idx, int64_t ncols, const int64_t* ncols_dev, const double* freqs, int64_t nf,
                   const RskBeam* beam, const double* pos_lmn, const double* uvw_m, int64_t nbl, double* vis, void* work,
                   size_t work_bytes, void* stream) {
  if (!spec || !spec->stokes || !beam) return RSK_ERR_NULL;
  if (nf < 0 || ncols < 0 || spec->n < 0 || spec->nf_in < 1 || nf > INT32_MAX || nbl < 0 || nbl > 65535) return RSK_ERR_SIZE;
  if (spec->mode != RSK_FLAT && spec->mode != RSK_SPECTRAL_INDEX && spec->mode != RSK_FULL) return RSK_ERR_UNSUPPORTED;
  if (beam->type < RSK_BEAM_JONES || beam->type > RSK_BEAM_UNIFORM) return RSK_ERR_MODE;
  if (spec->mode == RSK_SPECTRAL_INDEX && (!spec->spectral_index || !spec->ref_freq || !freqs)) return RSK_ERR_NULL;
  if (spec->mode != RSK_FULL && spec->nf_in != 1) return RSK_ERR_SIZE;
  if (spec->mode == RSK_FULL && !spec->freq_map && nf != spec->nf_in) return RSK_ERR_SIZE;
  const bool fringe = uvw_m != nullptr && nbl > 0;
  if (fringe && (!pos_lmn || !freqs)) return RSK_ERR_NULL;
  const int64_t nb = fringe ? nbl : 1;
  if (nf == 0) return RSK_OK;
  if (!vis) return RSK_ERR_NULL;
  if (!idx && ncols > spec->n) return RSK_ERR_SIZE;
  VisParams p = {};
  p.stokes = spec->stokes;
  p.nf_in = spec->nf_in;
  p.n = spec->n;
  p.alpha = spec->spectral_index;
  p.ref_freq = spec->ref_freq;
  p.freq_map = spec->freq_map;
  p.polarized = spec->polarized;
  p.pivot_hz = spec->pivot_hz;
  if (spec->polarized) {
    if (!coords) return RSK_ERR_NULL;
    if (coords->rot) {
      p.cs.rot = coords->rot;
    } else {
      const int rc = check_coords(coords, &p.cs);
      if (rc != RSK_OK) return rc;
      if (!alt_az || !ravg) return RSK_ERR_NULL;
      for (int i = 0; i < 9; ++i) p.ravg[i] = ravg[i];
    }
  }
  if (beam->type == RSK_BEAM_SHORT_DIPOLE && !alt_az) return RSK_ERR_NULL;
  if (beam->type == RSK_BEAM_JONES) {
    if (ncols > 0 && !beam->jones1) return RSK_ERR_NULL;
    if (beam->ld_jones < ncols) return RSK_ERR_LD;
    p.jones1 = reinterpret_cast<const double2*>(beam->jones1);
    p.jones2 = reinterpret_cast<const double2*>(beam->jones2);
    p.ld_jones = beam->ld_jones;
  }
  p.beam_type = beam->type;
  p.alt_az = alt_az;
  p.idx = idx;
  p.ncols = ncols;
  p.ncols_dev = ncols_dev;
  p.freqs = freqs;
  p.nf = (int)nf;
  p.lmn = fringe ? pos_lmn : nullptr;
  p.uvw = fringe ? uvw_m : nullptr;
  p.nbl = (int)nb;
  p.ngroups = visibility_groups(ncols);
  const size_t need = (size_t)nb * p.ngroups * nf * 8 * sizeof(double);
  if (!work || work_bytes < need) return RSK_ERR_WORKSPACE;
  p.partial = static_cast<double*>(work);
  cudaStream_t st = (cudaStream_t)stream;
  // work per thread: 8 frequencies x 1 baseline, 4 x 2 or 2 x 4 -- the apparent coherency is shared by the
  // baselines of a block, the per-column set-up by its frequencies
#define RSK_VIS(M, F, B)                                                                                 \
  do {                                                                                                   \
    const dim3 grid((unsigned)((nf + (F) - 1) / (F)), (unsigned)p.ngroups, (unsigned)((nb + (B) - 1) / (B))); \
    k_visibility<M, F, B><<<grid, kThreads, 0, st>>>(p);                                                 \
  } while (0)
#define RSK_VIS_MODE(F, B)                                                        \
  switch (spec->mode) {                                                           \
    case RSK_FLAT: RSK_VIS(RSK_FLAT, F, B); break;                                \
    case RSK_SPECTRAL_INDEX: RSK_VIS(RSK_SPECTRAL_INDEX, F, B); break;            \
    default: RSK_VIS(RSK_FULL, F, B); break;                                      \
  }
  if (nb >= 4) { RSK_VIS_MODE(2, 4) }
  else if (nb >= 2) { RSK_VIS_MODE(4, 2) }
  else { RSK_VIS_MODE(8, 1) }
#undef RSK_VIS_MODE
#undef RSK_VIS
  int rc = launch_status();
  if (rc != RSK_OK) return rc;
  const int64_t total = nb * nf * 8;
  k_visibility_finish<<<blocks_for(total), kThreads, 0, st>>>(p.partial, p.ngroups, (int)nf, total, vis);
  return launch_status();
}

size_t rsk_visibility_workspace_bytes(int64_t ncols, int64_t nf, int64_t nbl) {
  if (ncols < 0 || nf < 0 || nbl < 0) return 0;
  return (size_t)(nbl > 0 ? nbl : 1) * visibility_groups(ncols) * (size_t)nf * 8 * sizeof(double);
}

int rsk_at_frequencies(const RskSpectrum* spec, const double* freqs, int64_t nf, const int64_t* comps, int64_t ncomps,
                       double* stokes_out, void* stream) {
  EvalParams p = {};
  if (!spec) return RSK_ERR_NULL;
  if (comps && (ncomps < 0 || ncomps > spec->n)) return RSK_ERR_SIZE;
  const int64_t ncols = comps ? ncomps : spec->n;
  const int rc = fill_eval_params(spec, freqs, nf, ncols, &p);
  if (rc != RSK_OK) return rc;
  if (ncols > 0 && nf > 0 && !stokes_out) return RSK_ERR_NULL;
  p.idx = nullptr;
  p.cols = comps;
  p.out = stokes_out;
  p.ld_out = spec->n;
  p.col0 = 0;
  return launch_eval<kEpiStokes>(p, spec->mode, (cudaStream_t)stream);
}

int rsk_copy_2d(void* dst, size_t dst_pitch_bytes, const void* src, size_t src_pitch_bytes, size_t width_bytes,
                size_t height, void* stream) {
  if (width_bytes == 0 || height == 0) return RSK_OK;
  if (!dst || !src) return RSK_ERR_NULL;
  if (dst_pitch_bytes < width_bytes || src_pitch_bytes < width_bytes) return RSK_ERR_LD;
  const cudaError_t e = cudaMemcpy2DAsync(dst, dst_pitch_bytes, src, src_pitch_bytes, width_bytes, height,
                                          cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
  return e == cudaSuccess ? RSK_OK : (int)e;
}

int rsk_stokes_component_major(const double* stokes, int64_t nf, int64_t n, double* out, void* stream) {
  if (nf < 0 || n < 0) return RSK_ERR_SIZE;
  if (nf == 0 || n == 0) return RSK_OK;
  if (!stokes || !out) return RSK_ERR_NULL;
  const int64_t gx = (n + kCmTile - 1) / kCmTile, gy = (nf + kCmTile - 1) / kCmTile;
  if (gx > 0x7fffffffll || gy > 65535) return RSK_ERR_SIZE;
  k_stokes_component_major<<<dim3((unsigned)gx, (unsigned)gy), kThreads, 0, (cudaStream_t)stream>>>(stokes, nf, n, out);
  return launch_status();
}

int rsk_stokes_to_coherency(const double* stokes, int64_t nf, int64_t n, double* out, void* stream) {
  RskSpectrum spec = {};
  spec.mode = RSK_FULL;
  spec.stokes = stokes;
  spec.nf_in = nf;
  spec.n = n;
  if (nf < 0 || n < 0) return RSK_ERR_SIZE;
  if (nf == 0 || n == 0) return RSK_OK;
  RskCoords none = {};
  return rsk_coherency(&spec, &none, nullptr, nullptr, nullptr, nullptr, n, nullptr, nullptr, nf, out, n, 0, stream);
}

int rsk_coherency_to_stokes(const double* coherency, int64_t count, double* stokes, void* stream) {
  if (count < 0) return RSK_ERR_SIZE;
  if (count == 0) return RSK_OK;
  if (!coherency || !stokes) return RSK_ERR_NULL;
  if ((count + kThreads - 1) / kThreads > 0x7fffffffll) return RSK_ERR_SIZE;
  k_coherency_to_stokes<<<blocks_for(count), kThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const double2*>(coherency), count, stokes);
  return launch_status();
}

int rsk_convert_units(double* stokes, int64_t nf, int64_t n, const double* freq_per_chan, const double* freq_per_comp,
                      int to_jansky, void* stream) {
  if (nf < 0 || n < 0) return RSK_ERR_SIZE;
  if (nf == 0 || n == 0) return RSK_OK;
  if (!stokes || (!freq_per_chan && !freq_per_comp)) return RSK_ERR_NULL;
  const int64_t ntiles = (n + 2 * kThreads - 1) / (2 * kThreads);
  const int64_t nfchunks = (nf + kConvChunk - 1) / kConvChunk;
  if (ntiles * nfchunks > 0x7fffffffll) return RSK_ERR_SIZE;
  const bool vec = (n % 2 == 0) && (reinterpret_cast<uintptr_t>(stokes) & 15u) == 0;
  const unsigned grid = (unsigned)(ntiles * nfchunks);
  if (vec)
    k_convert_units<true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(stokes, nf, n, freq_per_chan, freq_per_comp,
                                                                        to_jansky, (int)nfchunks);
  else
    k_convert_units<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(stokes, nf, n, freq_per_chan, freq_per_comp,
                                                                         to_jansky, (int)nfchunks);
  return launch_status();
}

}  // extern "C"
