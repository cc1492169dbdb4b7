// rsk_device.cuh -- per-component fp64 math shared by the kernels (and, through RSK_HD, by the host
// helper rsk_ravg that pushes the three ICRS axes through the very same chain).
//
// What each piece replaces in the reference (RadioAstronomySoftwareGroup/pyradiosky):
//   rsk_chain()          astropy ICRS->AltAz as called at src/pyradiosky/skymodel.py:2451-2454
//                        (builtin_frames.utils.atciqz = eraLd + eraAb + eraRxp, then erfa.atioq) and
//                        the lmn / horizon lines :2464-2479
//   rsk_hpx_center()     astropy_healpix pixel centres, skymodel.py:1656-1669
//   rsk_basis_rotation() skymodel.py:2557-2646 + spherical_coords_transforms.py:140-186, :259-310
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define RSK_HD __host__ __device__ __forceinline__
#else
#define RSK_HD inline
#endif

#define RSK_D2PI 6.283185307179586476925287
#define RSK_DPI 3.141592653589793238462643
#define RSK_HALFPI 1.5707963267948966192313216916398
#define RSK_SRS 1.97412574336e-8 /* ERFA_SRS: Schwarzschild radius of the Sun, au */
#define RSK_CELMIN 1e-6          /* eraAtioq */
#define RSK_SELMIN 0.05          /* eraAtioq */

// Device-ready astrometry context (RSK_CTX_LEN doubles): eraASTROM with the orthogonal steps between
// aberration and the latitude rotation pre-multiplied on the host.
// The first 18 doubles are everything a folded context (see rsk_ctx_folded) needs per evaluation: nine aligned
// 16-byte shared-memory loads in the sweep kernel.
struct RskCtx {
  double eh[3];   // Sun -> observer unit vector
  double w2;      // SRS / em
  double v[3];    // observer barycentric velocity / c
  double bm1;     // sqrt(1 - |v|^2)
  double a1[9];   // polar-motion(xpl, ypl) . Rz(eral) . bpn : GCRS proper direction -> (-HA, Dec) frame
  double inv1pbm1; // 1 / (1 + bm1)
  double sphi, cphi;
  double diurab;
  double refa, refb;
  double em;      // Sun-observer distance, au
  double m_ident; // != 0: frame matrix is the identity, skip it
  double m[9];    // catalogue frame -> ICRS
};
static_assert(sizeof(RskCtx) == 34 * sizeof(double), "RskCtx layout");


// ---- lean fp64 primitives ---------------------------------------------------------------------------
// The position chain is bound by the fp64 pipe, not by HBM, once many time steps are evaluated per
// source (ncu, profiles/r01_summary.md), so the two atan2, the divisions and the square roots are
// written out branch-free here instead of calling libdevice (which spends about as many issue slots on
// special-case handling and on re-materialising 64-bit constants as on arithmetic).  Accuracy: ~1 ulp
// (2e-16 rad), five orders of magnitude inside the 1 micro-arcsecond budget.  On the host the libm
// versions are used.

// atan2 by table-driven argument reduction.  With t = min/max in [0, 1] and t_k = k/8 the nearest eighth,
//   atan(t) = atan(t_k) + atan((min - t_k max) / (max + t_k min)),
// so the reduction costs two fused multiply-adds *before* the one division, the interval index k comes from a
// single-precision quotient (fp32 / SFU pipes, not the fp64 pipe), and the remaining argument is below
// 1/16 + 2^-20: atan(u) = u + u s Q(s), s = u^2, with a degree-4 Q (mpmath.chebyfit on s in [0, 0.0064] as in
// tools/gen_atan_poly.py: max error of the double-precision Horner evaluation 1.4e-17).  The previous form, a degree-19 polynomial on
// [0, 1], spent twice the fp64 issue slots; the LST sweep is bound by that pipe (profiles/r01_summary.md).
#if defined(__CUDACC__)
__device__ const double2 kRskAtanTab[16] = {  // (atan(k/8), k/8)
    {0x0.0p+0, 0x0.0p+0},
    {0x1.fd5ba9aac2f6ep-4, 0x1.0000000000000p-3},
    {0x1.f5b75f92c80ddp-3, 0x1.0000000000000p-2},
    {0x1.6f61941e4def1p-2, 0x1.8000000000000p-2},
    {0x1.dac670561bb4fp-2, 0x1.0000000000000p-1},
    {0x1.1e00babdefeb4p-1, 0x1.4000000000000p-1},
    {0x1.4978fa3269ee1p-1, 0x1.8000000000000p-1},
    {0x1.700a7c5784634p-1, 0x1.c000000000000p-1},
    {0x1.921fb54442d18p-1, 0x1.0000000000000p+0},
    {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}};  // (padding: the index is not clamped)
// Azimuth table: entry [octant][k] = (B, +-k/8) with the octant = 4 (y < 0) + 2 (x < 0) + (|y| > |x|), so that
// atan2(y, x) mod 2 pi = B + sign * atan(u): every reflection of the arctangent (swap, x < 0, y < 0, the wrap of
// negative angles by eraAnp) is in B and in the sign, which rides on the second member (-0.0 for k = 0).
__device__ const double2 kRskAzTab[80] = {
    {0x0.0p+0, 0x0.0p+0}, {0x1.fd5ba9aac2f6ep-4, 0x1.0000000000000p-3}, {0x1.f5b75f92c80ddp-3, 0x1.0000000000000p-2}, {0x1.6f61941e4def1p-2, 0x1.8000000000000p-2}, {0x1.dac670561bb4fp-2, 0x1.0000000000000p-1},
    {0x1.1e00babdefeb4p-1, 0x1.4000000000000p-1}, {0x1.4978fa3269ee1p-1, 0x1.8000000000000p-1}, {0x1.700a7c5784634p-1, 0x1.c000000000000p-1}, {0x1.921fb54442d18p-1, 0x1.0000000000000p+0},  // octant 0
    {0x1.921fb54442d18p+0, -0x0.0p+0}, {0x1.7249faa996a21p+0, -0x1.0000000000000p-3}, {0x1.5368c951e9cfdp+0, -0x1.0000000000000p-2}, {0x1.3647503caf55cp+0, -0x1.8000000000000p-2}, {0x1.1b6e192ebbe44p+0, -0x1.0000000000000p-1},
    {0x1.031f57e54adbep+0, -0x1.4000000000000p-1}, {0x1.dac670561bb4fp-1, -0x1.8000000000000p-1}, {0x1.b434ee31013fdp-1, -0x1.c000000000000p-1}, {0x1.921fb54442d18p-1, -0x1.0000000000000p+0},  // octant 1
    {0x1.921fb54442d18p+1, -0x0.0p+0}, {0x1.8234d7f6ecb9dp+1, -0x1.0000000000000p-3}, {0x1.72c43f4b1650ap+1, -0x1.0000000000000p-2}, {0x1.643382c07913ap+1, -0x1.8000000000000p-2}, {0x1.56c6e7397f5aep+1, -0x1.0000000000000p-1},
    {0x1.4a9f8694c6d6bp+1, -0x1.4000000000000p-1}, {0x1.3fc176b7a8560p+1, -0x1.8000000000000p-1}, {0x1.361d162e61b8bp+1, -0x1.c000000000000p-1}, {0x1.2d97c7f3321d2p+1, -0x1.0000000000000p+0},  // octant 2
    {0x1.921fb54442d18p+0, 0x0.0p+0}, {0x1.b1f56fdeef00fp+0, 0x1.0000000000000p-3}, {0x1.d0d6a1369bd34p+0, 0x1.0000000000000p-2}, {0x1.edf81a4bd64d4p+0, 0x1.8000000000000p-2}, {0x1.0468a8ace4df6p+1, 0x1.0000000000000p-1},
    {0x1.109009519d639p+1, 0x1.4000000000000p-1}, {0x1.1b6e192ebbe44p+1, 0x1.8000000000000p-1}, {0x1.251279b802819p+1, 0x1.c000000000000p-1}, {0x1.2d97c7f3321d2p+1, 0x1.0000000000000p+0},  // octant 3
    {0x1.921fb54442d18p+2, -0x0.0p+0}, {0x1.8a2a469d97c5bp+2, -0x1.0000000000000p-3}, {0x1.8271fa47ac911p+2, -0x1.0000000000000p-2}, {0x1.7b299c025df29p+2, -0x1.8000000000000p-2}, {0x1.74734e3ee1163p+2, -0x1.0000000000000p-1},
    {0x1.6e5f9dec84d42p+2, -0x1.4000000000000p-1}, {0x1.68f095fdf593cp+2, -0x1.8000000000000p-1}, {0x1.641e65b952452p+2, -0x1.c000000000000p-1}, {0x1.5fdbbe9bba775p+2, -0x1.0000000000000p+0},  // octant 4
    {0x1.2d97c7f3321d2p+2, 0x0.0p+0}, {0x1.358d3699dd290p+2, 0x1.0000000000000p-3}, {0x1.3d4582efc85d9p+2, 0x1.0000000000000p-2}, {0x1.448de13516fc1p+2, 0x1.8000000000000p-2}, {0x1.4b442ef893d87p+2, 0x1.0000000000000p-1},
    {0x1.5157df4af01a9p+2, 0x1.4000000000000p-1}, {0x1.56c6e7397f5aep+2, 0x1.8000000000000p-1}, {0x1.5b99177e22a99p+2, 0x1.c000000000000p-1}, {0x1.5fdbbe9bba775p+2, 0x1.0000000000000p+0},  // octant 5
    {0x1.921fb54442d18p+1, 0x0.0p+0}, {0x1.a20a929198e94p+1, 0x1.0000000000000p-3}, {0x1.b17b2b3d6f526p+1, 0x1.0000000000000p-2}, {0x1.c00be7c80c8f6p+1, 0x1.8000000000000p-2}, {0x1.cd78834f06482p+1, 0x1.0000000000000p-1},
    {0x1.d99fe3f3becc5p+1, 0x1.4000000000000p-1}, {0x1.e47df3d0dd4d1p+1, 0x1.8000000000000p-1}, {0x1.ee22545a23ea5p+1, 0x1.c000000000000p-1}, {0x1.f6a7a2955385ep+1, 0x1.0000000000000p+0},  // octant 6
    {0x1.2d97c7f3321d2p+2, -0x0.0p+0}, {0x1.25a2594c87114p+2, -0x1.0000000000000p-3}, {0x1.1dea0cf69bdcbp+2, -0x1.0000000000000p-2}, {0x1.16a1aeb14d3e3p+2, -0x1.8000000000000p-2}, {0x1.0feb60edd061dp+2, -0x1.0000000000000p-1},
    {0x1.09d7b09b741fcp+2, -0x1.4000000000000p-1}, {0x1.0468a8ace4df6p+2, -0x1.8000000000000p-1}, {0x1.ff2cf0d083217p+1, -0x1.c000000000000p-1}, {0x1.f6a7a2955385ep+1, -0x1.0000000000000p+0},  // octant 7
    {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}};
// Sweep-grade tables (rsk_sweep_angles below): 1/16 spacing, so the reduced argument is below 1/32 and a degree-2
// polynomial in u^2 reaches 3e-15 rad; generated by tools/gen_sweep_tables.py.
__device__ const double2 kRskAtanTab16[32] = {
    {0x0.0p+0, 0x0.0p+0}, {0x1.ff55bb72cfdeap-5, 0x1.0000000000000p-4}, {0x1.fd5ba9aac2f6ep-4, 0x1.0000000000000p-3}, {0x1.7b97b4bce5b02p-3, 0x1.8000000000000p-3}, {0x1.f5b75f92c80ddp-3, 0x1.0000000000000p-2}, {0x1.362773707ebccp-2, 0x1.4000000000000p-2}, {0x1.6f61941e4def1p-2, 0x1.8000000000000p-2}, {0x1.a64eec3cc23fdp-2, 0x1.c000000000000p-2}, {0x1.dac670561bb4fp-2, 0x1.0000000000000p-1}, {0x1.0657e94db30d0p-1, 0x1.2000000000000p-1}, {0x1.1e00babdefeb4p-1, 0x1.4000000000000p-1}, {0x1.345f01cce37bbp-1, 0x1.6000000000000p-1}, {0x1.4978fa3269ee1p-1, 0x1.8000000000000p-1}, {0x1.5d58987169b18p-1, 0x1.a000000000000p-1}, {0x1.700a7c5784634p-1, 0x1.c000000000000p-1}, {0x1.819d0b7158a4dp-1, 0x1.e000000000000p-1}, {0x1.921fb54442d18p-1, 0x1.0000000000000p+0},
    {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}};
__device__ const double kRskCk16[32] = {
    0x1.0000000000000p+0, 0x1.ff00bf608b827p-1, 0x1.fc0bd88a0f1d9p-1, 0x1.f73b05f60fd39p-1, 0x1.f0b6848d2af1cp-1, 0x1.e8b19984b1caep-1, 0x1.df6689b7e6350p-1, 0x1.d51292feca74cp-1, 0x1.c9f25c5bfedd9p-1, 0x1.be3f2617e31c0p-1, 0x1.b22cd56d87d7ep-1, 0x1.a5e8d2b5f03a5p-1, 0x1.999999999999ap-1, 0x1.8d5ed070ef27dp-1, 0x1.8151bb86fee1dp-1, 0x1.7585e681fe6f2p-1, 0x1.6a09e667f3bcdp-1,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
__device__ const double2 kRskAzTab16[144] = {
    {0x0.0p+0, 0x0.0p+0}, {0x1.ff55bb72cfdeap-5, 0x1.0000000000000p-4}, {0x1.fd5ba9aac2f6ep-4, 0x1.0000000000000p-3}, {0x1.7b97b4bce5b02p-3, 0x1.8000000000000p-3}, {0x1.f5b75f92c80ddp-3, 0x1.0000000000000p-2}, {0x1.362773707ebccp-2, 0x1.4000000000000p-2}, {0x1.6f61941e4def1p-2, 0x1.8000000000000p-2}, {0x1.a64eec3cc23fdp-2, 0x1.c000000000000p-2}, {0x1.dac670561bb4fp-2, 0x1.0000000000000p-1}, {0x1.0657e94db30d0p-1, 0x1.2000000000000p-1}, {0x1.1e00babdefeb4p-1, 0x1.4000000000000p-1}, {0x1.345f01cce37bbp-1, 0x1.6000000000000p-1}, {0x1.4978fa3269ee1p-1, 0x1.8000000000000p-1}, {0x1.5d58987169b18p-1, 0x1.a000000000000p-1}, {0x1.700a7c5784634p-1, 0x1.c000000000000p-1}, {0x1.819d0b7158a4dp-1, 0x1.e000000000000p-1}, {0x1.921fb54442d18p-1, 0x1.0000000000000p+0},  // octant 0
    {0x1.921fb54442d18p+0, -0x0.0p+0}, {0x1.82250768ac529p+0, -0x1.0000000000000p-4}, {0x1.7249faa996a21p+0, -0x1.0000000000000p-3}, {0x1.62acbeaca61b8p+0, -0x1.8000000000000p-3}, {0x1.5368c951e9cfdp+0, -0x1.0000000000000p-2}, {0x1.4495d86823225p+0, -0x1.4000000000000p-2}, {0x1.3647503caf55cp+0, -0x1.8000000000000p-2}, {0x1.288bfa3512419p+0, -0x1.c000000000000p-2}, {0x1.1b6e192ebbe44p+0, -0x1.0000000000000p-1}, {0x1.0ef3c09d694b0p+0, -0x1.2000000000000p-1}, {0x1.031f57e54adbep+0, -0x1.4000000000000p-1}, {0x1.efe068bba2275p-1, -0x1.6000000000000p-1}, {0x1.dac670561bb4fp-1, -0x1.8000000000000p-1}, {0x1.c6e6d2171bf18p-1, -0x1.a000000000000p-1}, {0x1.b434ee31013fdp-1, -0x1.c000000000000p-1}, {0x1.a2a25f172cfe4p-1, -0x1.e000000000000p-1}, {0x1.921fb54442d18p-1, -0x1.0000000000000p+0},  // octant 1
    {0x1.921fb54442d18p+1, -0x0.0p+0}, {0x1.8a225e5677921p+1, -0x1.0000000000000p-4}, {0x1.8234d7f6ecb9dp+1, -0x1.0000000000000p-3}, {0x1.7a6639f874768p+1, -0x1.8000000000000p-3}, {0x1.72c43f4b1650ap+1, -0x1.0000000000000p-2}, {0x1.6b5ac6d632f9fp+1, -0x1.4000000000000p-2}, {0x1.643382c07913ap+1, -0x1.8000000000000p-2}, {0x1.5d55d7bcaa899p+1, -0x1.c000000000000p-2}, {0x1.56c6e7397f5aep+1, -0x1.0000000000000p-1}, {0x1.5089baf0d60e4p+1, -0x1.2000000000000p-1}, {0x1.4a9f8694c6d6bp+1, -0x1.4000000000000p-1}, {0x1.4507f4d109f29p+1, -0x1.6000000000000p-1}, {0x1.3fc176b7a8560p+1, -0x1.8000000000000p-1}, {0x1.3ac98f27e8652p+1, -0x1.a000000000000p-1}, {0x1.361d162e61b8bp+1, -0x1.c000000000000p-1}, {0x1.31b87267eca85p+1, -0x1.e000000000000p-1}, {0x1.2d97c7f3321d2p+1, -0x1.0000000000000p+0},  // octant 2
    {0x1.921fb54442d18p+0, 0x0.0p+0}, {0x1.a21a631fd9508p+0, 0x1.0000000000000p-4}, {0x1.b1f56fdeef00fp+0, 0x1.0000000000000p-3}, {0x1.c192abdbdf879p+0, 0x1.8000000000000p-3}, {0x1.d0d6a1369bd34p+0, 0x1.0000000000000p-2}, {0x1.dfa992206280bp+0, 0x1.4000000000000p-2}, {0x1.edf81a4bd64d4p+0, 0x1.8000000000000p-2}, {0x1.fbb3705373617p+0, 0x1.c000000000000p-2}, {0x1.0468a8ace4df6p+1, 0x1.0000000000000p-1}, {0x1.0aa5d4f58e2c0p+1, 0x1.2000000000000p-1}, {0x1.109009519d639p+1, 0x1.4000000000000p-1}, {0x1.16279b155a47bp+1, 0x1.6000000000000p-1}, {0x1.1b6e192ebbe44p+1, 0x1.8000000000000p-1}, {0x1.206600be7bd52p+1, 0x1.a000000000000p-1}, {0x1.251279b802819p+1, 0x1.c000000000000p-1}, {0x1.29771d7e7791fp+1, 0x1.e000000000000p-1}, {0x1.2d97c7f3321d2p+1, 0x1.0000000000000p+0},  // octant 3
    {0x1.921fb54442d18p+2, -0x0.0p+0}, {0x1.8e2109cd5d31cp+2, -0x1.0000000000000p-4}, {0x1.8a2a469d97c5bp+2, -0x1.0000000000000p-3}, {0x1.8642f79e5ba40p+2, -0x1.8000000000000p-3}, {0x1.8271fa47ac911p+2, -0x1.0000000000000p-2}, {0x1.7ebd3e0d3ae5cp+2, -0x1.4000000000000p-2}, {0x1.7b299c025df29p+2, -0x1.8000000000000p-2}, {0x1.77bac68076ad8p+2, -0x1.c000000000000p-2}, {0x1.74734e3ee1163p+2, -0x1.0000000000000p-1}, {0x1.7154b81a8c6fep+2, -0x1.2000000000000p-1}, {0x1.6e5f9dec84d42p+2, -0x1.4000000000000p-1}, {0x1.6b93d50aa6621p+2, -0x1.6000000000000p-1}, {0x1.68f095fdf593cp+2, -0x1.8000000000000p-1}, {0x1.6674a236159b5p+2, -0x1.a000000000000p-1}, {0x1.641e65b952452p+2, -0x1.c000000000000p-1}, {0x1.61ec13d617bcfp+2, -0x1.e000000000000p-1}, {0x1.5fdbbe9bba775p+2, -0x1.0000000000000p+0},  // octant 4
    {0x1.2d97c7f3321d2p+2, 0x0.0p+0}, {0x1.3196736a17bcep+2, 0x1.0000000000000p-4}, {0x1.358d3699dd290p+2, 0x1.0000000000000p-3}, {0x1.39748599194aap+2, 0x1.8000000000000p-3}, {0x1.3d4582efc85d9p+2, 0x1.0000000000000p-2}, {0x1.40fa3f2a3a08fp+2, 0x1.4000000000000p-2}, {0x1.448de13516fc1p+2, 0x1.8000000000000p-2}, {0x1.47fcb6b6fe412p+2, 0x1.c000000000000p-2}, {0x1.4b442ef893d87p+2, 0x1.0000000000000p-1}, {0x1.4e62c51ce87ecp+2, 0x1.2000000000000p-1}, {0x1.5157df4af01a9p+2, 0x1.4000000000000p-1}, {0x1.5423a82cce8cap+2, 0x1.6000000000000p-1}, {0x1.56c6e7397f5aep+2, 0x1.8000000000000p-1}, {0x1.5942db015f535p+2, 0x1.a000000000000p-1}, {0x1.5b99177e22a99p+2, 0x1.c000000000000p-1}, {0x1.5dcb69615d31cp+2, 0x1.e000000000000p-1}, {0x1.5fdbbe9bba775p+2, 0x1.0000000000000p+0},  // octant 5
    {0x1.921fb54442d18p+1, 0x0.0p+0}, {0x1.9a1d0c320e110p+1, 0x1.0000000000000p-4}, {0x1.a20a929198e94p+1, 0x1.0000000000000p-3}, {0x1.a9d93090112c8p+1, 0x1.8000000000000p-3}, {0x1.b17b2b3d6f526p+1, 0x1.0000000000000p-2}, {0x1.b8e4a3b252a92p+1, 0x1.4000000000000p-2}, {0x1.c00be7c80c8f6p+1, 0x1.8000000000000p-2}, {0x1.c6e992cbdb198p+1, 0x1.c000000000000p-2}, {0x1.cd78834f06482p+1, 0x1.0000000000000p-1}, {0x1.d3b5af97af94cp+1, 0x1.2000000000000p-1}, {0x1.d99fe3f3becc5p+1, 0x1.4000000000000p-1}, {0x1.df3775b77bb07p+1, 0x1.6000000000000p-1}, {0x1.e47df3d0dd4d1p+1, 0x1.8000000000000p-1}, {0x1.e975db609d3dep+1, 0x1.a000000000000p-1}, {0x1.ee22545a23ea5p+1, 0x1.c000000000000p-1}, {0x1.f286f82098fabp+1, 0x1.e000000000000p-1}, {0x1.f6a7a2955385ep+1, 0x1.0000000000000p+0},  // octant 6
    {0x1.2d97c7f3321d2p+2, -0x0.0p+0}, {0x1.29991c7c4c7d6p+2, -0x1.0000000000000p-4}, {0x1.25a2594c87114p+2, -0x1.0000000000000p-3}, {0x1.21bb0a4d4aefap+2, -0x1.8000000000000p-3}, {0x1.1dea0cf69bdcbp+2, -0x1.0000000000000p-2}, {0x1.1a3550bc2a315p+2, -0x1.4000000000000p-2}, {0x1.16a1aeb14d3e3p+2, -0x1.8000000000000p-2}, {0x1.1332d92f65f92p+2, -0x1.c000000000000p-2}, {0x1.0feb60edd061dp+2, -0x1.0000000000000p-1}, {0x1.0ccccac97bbb8p+2, -0x1.2000000000000p-1}, {0x1.09d7b09b741fcp+2, -0x1.4000000000000p-1}, {0x1.070be7b995adbp+2, -0x1.6000000000000p-1}, {0x1.0468a8ace4df6p+2, -0x1.8000000000000p-1}, {0x1.01ecb4e504e6fp+2, -0x1.a000000000000p-1}, {0x1.ff2cf0d083217p+1, -0x1.c000000000000p-1}, {0x1.fac84d0a0e111p+1, -0x1.e000000000000p-1}, {0x1.f6a7a2955385ep+1, -0x1.0000000000000p+0},  // octant 7
    {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}};
// atan(u) = u + u s (q0 + q1 s + q2 s^2), |u| <= 1/32 + 2^-20: Taylor coefficients, truncation u^9 / 9 = 3.2e-15 rad;
// sqrt(1 + s) = 1 + s (1/2 + s (-1/8 + s / 16)), s <= 9.8e-4: truncation 5 s^4 / 128 = 3.6e-14 relative
__constant__ double kRskSweepQ[6] = {-1.0 / 3.0, 1.0 / 5.0, -1.0 / 7.0, 0.5, -0.125, 0.0625};
// 64-bit literals the chain needs: operands straight from the constant bank (a literal in the code costs two
// UMOV per use inside the sweep loop)
__constant__ double kRskC[4] = {1e-6, 1e-300, RSK_HALFPI, RSK_DPI};
// atan(u) = u + u s Q(s), s = u^2, |u| <= 0.08 (degree 4: 1.4e-17 maximum error)
__constant__ double kRskAtanQ[5] = {-0x1.5555555555539p-2, 0x1.999999992c40dp-3, -0x1.2492470e5748ap-3,
                                    0x1.c7154d717ee6ap-4, -0x1.6f5e1489511f6p-4};
#endif

// a / b for finite a and normal b > 0
RSK_HD double rsk_div(double a, double b) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));  // ~20 good bits
  const double e = fma(-b, r, 1.0);
  r = fma(r, fma(e, e, e), r);                             // cubic step: ~2^-60
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);                         // one residual correction
#else
  return a / b;
#endif
}

// 1 / sqrt(x) for normal x > 0
RSK_HD double rsk_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x * r, r, 1.0);
  r = fma(r, fma(0.375 * e, e, 0.5 * e), r);               // cubic step
  const double e2 = fma(-x * r, r, 1.0);
  return fma(0.5 * r, e2, r);
#else
  return 1.0 / sqrt(x);
#endif
}

// sqrt(x) for x >= 0 (0 -> 0): reciprocal-square-root seed, one coupled (Goldschmidt) iteration on
// g ~ sqrt(x), h ~ 1 / (2 sqrt(x)), one residual correction: 7 fp64 instructions, < 1 ulp.
RSK_HD double rsk_sqrt(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x + kRskC[1]));  // x = 0 -> 0 (the bias matters below 1e-284)
  double g = x * r, h = 0.5 * r;
  const double e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  return fma(fma(-g, g, x), h, g);
#else
  return sqrt(x);
#endif
}

// atan2(y, x) for finite arguments; result in (-pi, pi], 0 for y = x = 0.  Written for the issue slots it
// takes, not only the fp64 ones (the sweep kernel is issue-bound): explicit compare + select instead of
// fmin/fmax (whose NaN handling costs ~8 instructions each), coefficients in the constant bank (operands of
// DFMA, no register moves), no branch.
// XPOS: the caller guarantees x >= 0 and (x, y) != (0, 0) (result in [-pi/2, pi/2]; two fix-ups fewer).
template <bool XPOS>
RSK_HD double rsk_atan2_t(double y, double x) {
#if defined(__CUDA_ARCH__)
  const bool sw = fabs(y) > fabs(x);
  const double mn = fabs(sw ? x : y), mx = fabs(sw ? y : x);
  // interval index from a single-precision quotient: fp32 / SFU pipes.  (float)mx == 0 needs mx < 1e-45, i.e.
  // a direction closer than that to the pole of the angle, where k = 0 still gives a finite, if inexact, value.
  float fr;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(fr) : "f"((float)mx));
  // not clamped: mn <= mx survives the rounding to float and the reciprocal is good to an ulp, so the quotient
  // is at most 1 + 2^-22 and k <= 8; NaN (0 * inf) converts to 0
  const int k = __float2int_rn(8.0f * ((float)mn * fr));
  const double2 tab = __ldg(&kRskAtanTab[k]);
  const double num = fma(-tab.y, mx, mn);
  const double den = fma(tab.y, mn, mx);
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
  const double e = fma(-den, r, 1.0);
  r = fma(r, fma(e, e, e), r);  // 1 / den to ~2^-60
  const double u = num * r;     // |u| <= 0.07: an ulp of u is below 1e-17
  const double s = u * u;
  double p = kRskAtanQ[4];
  p = fma(p, s, kRskAtanQ[3]);
  p = fma(p, s, kRskAtanQ[2]);
  p = fma(p, s, kRskAtanQ[1]);
  p = fma(p, s, kRskAtanQ[0]);
  double a = tab.x + fma(u * s, p, u);
  a = sw ? kRskC[2] - a : a;
  if (!XPOS) {
    a = x < 0.0 ? kRskC[3] - a : a;
    a = mx > 0.0 ? a : 0.0;  // atan2(0, 0) = 0 (eraAtioq's own convention for the azimuth at the zenith)
  }
  return copysign(a, y);
#else
  return (y != 0.0 || x != 0.0) ? atan2(y, x) : 0.0;
#endif
}

RSK_HD double rsk_atan2(double y, double x) { return rsk_atan2_t<false>(y, x); }

// eraAnp(atan2(y, x)), 0 for y = x = 0: the azimuth of eraAtioq in [0, 2 pi] in one go.  Same reduction as
// rsk_atan2_t, but the octant picks (B, sign) from kRskAzTab, so the reflections and the wrap cost one table load, one
// sign flip and the final addition instead of three compares, three subtractions and four 64-bit selects.
// (Like eraAnp, an angle within an ulp below zero comes back as 2 pi itself.)
RSK_HD double rsk_azimuth(double y, double x) {
#if defined(__CUDA_ARCH__)
  const bool sw = fabs(y) > fabs(x);
  const double mn = fabs(sw ? x : y);
  double mx = fabs(sw ? y : x);
  const float fmx = (float)mx;
  float fr;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(fr) : "f"(fmx));
  const int k = __float2int_rn(8.0f * ((float)mn * fr));  // <= 8 and 0 for NaN, see rsk_atan2_t
  int oct = (int)((unsigned)__double2hiint(y) >> 31) * 4 + (int)((unsigned)__double2hiint(x) >> 31) * 2 + (sw ? 1 : 0);
  oct = fmx == 0.0f ? 0 : oct;  // the zenith itself: azimuth 0 whatever the signs of the zeros
  const double2 tab = __ldg(&kRskAzTab[oct * 9 + k]);
  // keeps the quotient below finite (0) at the zenith: the smallest normal number instead of 0
  mx = __hiloint2double(max(__double2hiint(mx), 0x00100000), __double2loint(mx));
  const double t = fabs(tab.y);
  const double num = fma(-t, mx, mn);
  const double den = fma(t, mn, mx);
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
  const double e = fma(-den, r, 1.0);
  r = fma(r, fma(e, e, e), r);
  const double u = num * r;
  const double s = u * u;
  double p = kRskAtanQ[4];
  p = fma(p, s, kRskAtanQ[3]);
  p = fma(p, s, kRskAtanQ[2]);
  p = fma(p, s, kRskAtanQ[1]);
  p = fma(p, s, kRskAtanQ[0]);
  const double w = fma(u * s, p, u);
  const double ws = __hiloint2double(__double2hiint(w) ^ (__double2hiint(tab.y) & 0x80000000), __double2loint(w));
  return tab.x + ws;
#else
  const double a = (y != 0.0 || x != 0.0) ? atan2(y, x) : 0.0;
  return a < 0.0 ? a + RSK_D2PI : a;
#endif
}

// Azimuth (eraAnp(atan2(y, x)), 0 at the zenith) and altitude (atan2(z, sqrt(x^2 + y^2))) of a direction in one go,
// for the LST-sweep kernel, which is bound by fp64 issue (profiles/r02_sweep_floor.md): 30 fp64 instructions instead of
// the 43 of rsk_azimuth + rsk_sqrt + rsk_atan2_t, paid for with accuracy margin -- 1/16-spaced tables and a degree-2
// polynomial (3e-15 rad), one quadratic Newton step on the reciprocal seeds (relative 2^-40 on |u| <= 1/32: 3e-14
// rad), and sqrt(x^2 + y^2) taken from the azimuth's own reduction, rho = den cos(phi_k) sqrt(1 + u^2), instead of a
// square root (4e-14 relative).  Total below 1e-13 rad = 0.02 micro-arcseconds against the 1 uas criterion.
#if defined(__CUDACC__)
__device__ __forceinline__ void rsk_sweep_angles(double x, double y, double z, double& az_out, double& alt_out) {
  const bool sw = fabs(y) > fabs(x);
  const double mn = fabs(sw ? x : y);
  double mx = fabs(sw ? y : x);
  const float fmx = (float)mx;
  float fr;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(fr) : "f"(fmx));
  const int k = __float2int_rn(16.0f * ((float)mn * fr));  // <= 16 and 0 for NaN, as in rsk_atan2_t
  int oct = (int)((unsigned)__double2hiint(y) >> 31) * 4 + (int)((unsigned)__double2hiint(x) >> 31) * 2 + (sw ? 1 : 0);
  oct = fmx == 0.0f ? 0 : oct;  // the zenith itself: azimuth 0 whatever the signs of the zeros
  const double2 tab = __ldg(&kRskAzTab16[oct * 17 + k]);
  const double ck = __ldg(&kRskCk16[k]);
  mx = __hiloint2double(max(__double2hiint(mx), 0x00100000), __double2loint(mx));  // finite quotient at the zenith
  const double t = fabs(tab.y);
  const double num = fma(-t, mx, mn);
  const double den = fma(t, mn, mx);
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
  r = fma(r, fma(-den, r, 1.0), r);
  const double u = num * r;
  const double s = u * u;
  const double p = fma(fma(kRskSweepQ[2], s, kRskSweepQ[1]), s, kRskSweepQ[0]);
  const double w = fma(u * s, p, u);
  const double ws = __hiloint2double(__double2hiint(w) ^ (__double2hiint(tab.y) & 0x80000000), __double2loint(w));
  az_out = tab.x + ws;
  // rho = |(x, y)| = den cos(phi_k) sqrt(1 + u^2)
  const double rho = (den * ck) * fma(s, fma(fma(kRskSweepQ[5], s, kRskSweepQ[4]), s, kRskSweepQ[3]), 1.0);
  // alt = atan2(z, rho), rho >= 0
  const double az_ = fabs(z);
  const bool sw2 = az_ > rho;
  const double mn2 = sw2 ? rho : az_, mx2 = sw2 ? az_ : rho;
  float fr2;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(fr2) : "f"((float)mx2));
  const int k2 = __float2int_rn(16.0f * ((float)mn2 * fr2));
  const double2 tab2 = __ldg(&kRskAtanTab16[k2]);
  const double num2 = fma(-tab2.y, mx2, mn2);
  const double den2 = fma(tab2.y, mn2, mx2);
  double r2;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r2) : "d"(den2));
  r2 = fma(r2, fma(-den2, r2, 1.0), r2);
  const double u2 = num2 * r2;
  const double s2 = u2 * u2;
  const double p2 = fma(fma(kRskSweepQ[2], s2, kRskSweepQ[1]), s2, kRskSweepQ[0]);
  double a = tab2.x + fma(u2 * s2, p2, u2);
  a = sw2 ? kRskC[2] - a : a;
  alt_out = copysign(a, z);
}
#endif

// sin and cos of an angle of catalogue size (|x| <= 16: longitudes, latitudes, HEALPix azimuths): two-term
// Cody-Waite reduction by pi/2 (exact for |k| <= 10) and the fdlibm kernel polynomials on [-pi/4, pi/4]
// (< 1 ulp, checked against long double libm).  About a third of the issue slots of libdevice's sincos,
// which carries a Payne-Hanek path for huge arguments; larger |x| take that path here too.
RSK_HD void rsk_sincos(double x, double* sn, double* cs) {
#if defined(__CUDA_ARCH__)
  if (fabs(x) > 16.0) {
    sincos(x, sn, cs);
    return;
  }
  const double k = rint(x * 0.63661977236758134308);  // 2 / pi
  double r = fma(-k, 1.57079632679489655800e+00, x);
  r = fma(-k, 6.12323399573676603587e-17, r);
  const int q = (int)k;
  const double z = r * r;
  double ps = 1.58969099521155010221e-10;
  ps = fma(ps, z, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  const double ks = fma(r * z, ps, r);
  double pc = -1.13596475577881948265e-11;
  pc = fma(pc, z, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  const double kc = 1.0 - fma(-z * z, pc, 0.5 * z);
  const double s0 = (q & 1) ? kc : ks;
  const double c0 = (q & 1) ? ks : kc;
  *sn = (q & 2) ? -s0 : s0;
  *cs = ((q + 1) & 2) ? -c0 : c0;
#else
  *sn = sin(x);
  *cs = cos(x);
#endif
}

struct RskAltAz {
  double alt, az;   // radians
  double l, m, n;   // direction cosines: east, north, up
};

// A context is "folded" when it has no diurnal-aberration term and no refraction (every context astropy
// builds for the reference's AltAz(obstime, location), skymodel.py:2453): rsk_pack_context then multiplies
// the latitude rotation into a1 and stores sphi = 1, cphi = 0, and the chain below skips the steps that
// have become identities.
RSK_HD bool rsk_ctx_folded(const RskCtx& c) {
  return c.cphi == 0.0 && c.sphi == 1.0 && c.diurab == 0.0 && c.refa == 0.0 && c.refb == 0.0;
}

// (px,py,pz): unit vector of the component in the catalogue frame (in ICRS already when FRAME_DONE).
// WANT_LMN = false skips the direction cosines (the LST-sweep kernel does not always need them).
// FOLD: 0 = look at the context, 1 = the caller has checked rsk_ctx_folded(c) already (the sweep kernel does it
// once per block, outside its time loop).
// UNIT: (with FRAME_DONE) the caller's vector is of unit length to rounding, so p.p is taken as 1 in the light
// deflection (a relative change of 1e-16 in a term below 3e-5 rad).
// SWEEP: (device, folded contexts, no direction cosines) alt / az from rsk_sweep_angles instead of the full-margin
// routines -- the LST-sweep kernel only.
template <bool WANT_LMN, bool FRAME_DONE = false, int FOLD = 0, bool UNIT = false, bool SWEEP = false>
RSK_HD RskAltAz rsk_chain_t(const RskCtx& c, double px, double py, double pz, double pp_in = -1.0) {
  if (!FRAME_DONE && c.m_ident == 0.0) {
    const double x = c.m[0] * px + c.m[1] * py + c.m[2] * pz;
    const double y = c.m[3] * px + c.m[4] * py + c.m[5] * pz;
    const double z = c.m[6] * px + c.m[7] * py + c.m[8] * pz;
    px = x; py = y; pz = z;
  }
  // eraLd(bm=1, p, q=p, e=eh, em, dlim=1e-6): light deflection by the Sun (astropy atciqz sets q = p
  // for coordinates without distance, which is every SkyModel component).  p x (e x p) is expanded as
  // e (p.p) - p (p.e), and q.(q+e) = p.p + p.e.  The deflection is below 1e-5 rad, so a quotient good to
  // 2^-40 (reciprocal seed + one Newton step) leaves an error below 1e-17 rad.
  // p.p: time independent, so callers that loop over contexts pass it in (pp_in >= 0)
  const double pp = (FRAME_DONE && UNIT) ? 1.0 : (FRAME_DONE && pp_in >= 0.0) ? pp_in : px * px + py * py + pz * pz;
  const double pe = px * c.eh[0] + py * c.eh[1] + pz * c.eh[2];
#if defined(__CUDA_ARCH__)
  double den = pp + pe;
  den = den < kRskC[0] ? kRskC[0] : den;
  double rd;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rd) : "d"(den));
  rd = fma(fma(-den, rd, 1.0), rd, rd);
  const double w = c.w2 * rd;  // SRS / em / max(q.(q+e), dlim)
#else
  const double w = c.w2 / fmax(pp + pe, 1e-6);
#endif
  double nx, ny, nz;
  if (FRAME_DONE && UNIT) {
    nx = fma(w, fma(-px, pe, c.eh[0]), px);
    ny = fma(w, fma(-py, pe, c.eh[1]), py);
    nz = fma(w, fma(-pz, pe, c.eh[2]), pz);
  } else {
    nx = px + w * (c.eh[0] * pp - px * pe);
    ny = py + w * (c.eh[1] * pp - py * pe);
    nz = pz + w * (c.eh[2] * pp - pz * pe);
  }
  // eraAb(pnat, v, s=em, bm1): annual (+ observer) aberration
  // p' = pnat bm1 + w1 v + w2 (v - (pnat.v) pnat), grouped as pnat (bm1 - w2 pnat.v) + v (w1 + w2)
  const double pdv = nx * c.v[0] + ny * c.v[1] + nz * c.v[2];
  const double g1 = fma(-c.w2, pdv, c.bm1);
  const double g2 = fma(pdv, c.inv1pbm1, 1.0) + c.w2;
  double ax = fma(nx, g1, c.v[0] * g2);
  double ay = fma(ny, g1, c.v[1] * g2);
  double az_ = fma(nz, g1, c.v[2] * g2);
  RskAltAz o;
  if (FOLD == 1 || rsk_ctx_folded(c)) {
    // No diurnal aberration, no refraction: everything after the aberration is one rotation (a1 holds
    // latitude . polar motion . Earth rotation . bpn) and both angles are ratios of its components, so the
    // normalisation of eraAb is not needed for them.
    // (the north component: -x of the (-HA, Dec) -> Az/El step, negated term by term so that no fp64 instruction
    // is spent on the sign)
    const double xn = (-c.a1[0]) * ax - c.a1[1] * ay - c.a1[2] * az_;
    const double yaeo = c.a1[3] * ax + c.a1[4] * ay + c.a1[5] * az_;
    const double zaeo = c.a1[6] * ax + c.a1[7] * ay + c.a1[8] * az_;
#if defined(__CUDA_ARCH__)
    if (SWEEP && !WANT_LMN) {
      rsk_sweep_angles(xn, yaeo, zaeo, o.az, o.alt);
      o.l = o.m = o.n = 0.0;
      return o;
    }
#endif
    o.az = rsk_azimuth(yaeo, xn);  // eraAtioq + eraAnp: [0, 2 pi], 0 at the zenith
    const double r2 = xn * xn + yaeo * yaeo;
    const double rxy = rsk_sqrt(r2);
    // astropy forms pi/2 - zd with zd = atan2(rxy, z); atan2(z, rxy) is the same angle (to an ulp of pi/2)
    // without the two reflections
    o.alt = rsk_atan2_t<true>(zaeo, rxy);
    if (WANT_LMN) {
      const double ninv = rsk_rsqrt(r2 + zaeo * zaeo);
      o.l = yaeo * ninv;
      o.m = xn * ninv;
      o.n = zaeo * ninv;
    } else {
      o.l = o.m = o.n = 0.0;
    }
    return o;
  }
  const double rinv = rsk_rsqrt(ax * ax + ay * ay + az_ * az_);
  ax *= rinv; ay *= rinv; az_ *= rinv;
  // eraRxp(bpn) + eraAtioq up to the polar-motion step, as one host-built rotation.  (ERFA goes
  // through c2s -> anp -> s2c(ri - eral, di); rotating the vector directly is the same to ~1e-16.)
  const double xhd = c.a1[0] * ax + c.a1[1] * ay + c.a1[2] * az_;
  const double yhd = c.a1[3] * ax + c.a1[4] * ay + c.a1[5] * az_;
  const double zhd = c.a1[6] * ax + c.a1[7] * ay + c.a1[8] * az_;
  // diurnal aberration (0 in contexts made by eraApco, kept for generality)
  const double f = 1.0 - c.diurab * yhd;
  const double xhdt = f * xhd;
  const double yhdt = f * (yhd + c.diurab);
  const double zhdt = f * zhd;
  // (-HA, Dec) -> Az/El cartesian (S=0, E=90)
  const double xaet = c.sphi * xhdt - c.cphi * zhdt;
  const double yaet = yhdt;
  const double zaet = c.cphi * xhdt + c.sphi * zhdt;
  const double azobs = rsk_atan2(yaet, -xaet);  // 0 at the zenith, as eraAtioq
  // A tan z + B tan^3 z refraction (identity for refa = refb = 0, which is what AltAz(obstime,
  // location) with its default zero pressure gives at skymodel.py:2453)
  double xaeo = xaet, yaeo = yaet, zaeo = zaet;
  const bool refracted = (c.refa != 0.0 || c.refb != 0.0);
  if (refracted) {
    const double r = fmax(sqrt(xaet * xaet + yaet * yaet), RSK_CELMIN);
    const double z = fmax(zaet, RSK_SELMIN);
    const double tz = r / z;
    const double wr = c.refb * tz * tz;
    const double del = (c.refa + wr) * tz / (1.0 + (c.refa + 3.0 * wr) / (z * z));
    const double cosdel = 1.0 - del * del / 2.0;
    const double fr = cosdel - del * z / r;
    xaeo = xaet * fr;
    yaeo = yaet * fr;
    zaeo = cosdel * zaet + del * r;
  }
  const double rxy = rsk_sqrt(xaeo * xaeo + yaeo * yaeo);
  const double zdobs = rsk_atan2(rxy, zaeo);
  o.alt = RSK_HALFPI - zdobs;                       // astropy: lat = PIOVER2 - zen
  o.az = azobs < 0.0 ? azobs + RSK_D2PI : azobs;    // eraAnp
  if (WANT_LMN) {
    // skymodel.py:2468-2470: l = sin(az) cos(alt), m = cos(az) cos(alt), n = sin(alt); taken from the
    // cartesian vector instead of four more trig calls.  The vector is already of unit length unless
    // refraction or diurnal aberration rescaled it.
    double ninv = 1.0;
    if (refracted || c.diurab != 0.0) ninv = rsk_rsqrt(rxy * rxy + zaeo * zaeo);
    o.l = yaeo * ninv;
    o.m = -xaeo * ninv;
    o.n = zaeo * ninv;
  } else {
    o.l = o.m = o.n = 0.0;
  }
  return o;
}

RSK_HD RskAltAz rsk_chain(const RskCtx& c, double px, double py, double pz) {
  return rsk_chain_t<true>(c, px, py, pz);
}

// ---- HEALPix pixel centres -----------------------------------------------------------------------

RSK_HD int64_t rsk_isqrt(int64_t v) {
  int64_t r = (int64_t)sqrt((double)v);
  if (r * r > v) --r;
  else if ((r + 1) * (r + 1) <= v) ++r;
  return r;
}

RSK_HD int64_t rsk_compact_bits(uint64_t v) {
  v &= 0x5555555555555555ull;
  v = (v | (v >> 1)) & 0x3333333333333333ull;
  v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0Full;
  v = (v | (v >> 4)) & 0x00FF00FF00FF00FFull;
  v = (v | (v >> 8)) & 0x0000FFFF0000FFFFull;
  v = (v | (v >> 16)) & 0x00000000FFFFFFFFull;
  return (int64_t)v;
}

// Pixel centre as (z = sin lat, st = cos lat, phi = lon).  st is formed from the exact 1-|z| in the
// polar caps so there is no cancellation near the poles.
RSK_HD void rsk_hpx_center(int64_t nside, int nested, int64_t pix, double& z, double& st, double& phi) {
  const int64_t nl4 = 4 * nside;
  const int64_t npix = 12 * nside * nside;
  const int64_t ncap = 2 * nside * (nside - 1);
  int64_t nr, jp, kshift, jr_eq;
  int region;  // +1 north cap, 0 belt, -1 south cap
  if (nested) {
    // nside is a power of two here (checked by the entry points): shifts and masks instead of 64-bit divisions
#if defined(__CUDA_ARCH__)
    const int sh = 2 * (63 - __clzll((long long)nside));
#else
    const int sh = 2 * (63 - __builtin_clzll((unsigned long long)nside));
#endif
    const int64_t face = pix >> sh;
    const uint64_t ipf = (uint64_t)pix & (((uint64_t)1 << sh) - 1);
    const int64_t ix = rsk_compact_bits(ipf);
    const int64_t iy = rsk_compact_bits(ipf >> 1);
    const int64_t jrll = 2 + (face >> 2);
    const int64_t fm = face & 3;
    const int64_t jpll = (face >> 2) == 1 ? 2 * fm : 2 * fm + 1;
    const int64_t jr = jrll * nside - ix - iy - 1;
    if (jr < nside) { region = 1; nr = jr; kshift = 0; }
    else if (jr > 3 * nside) { region = -1; nr = nl4 - jr; kshift = 0; }
    else { region = 0; nr = nside; kshift = (jr - nside) & 1; }
    jp = (jpll * nr + ix - iy + 1 + kshift) / 2;
    if (jp > nl4) jp -= nl4;
    if (jp < 1) jp += nl4;
    jr_eq = jr;
  } else if (pix < ncap) {
    region = 1;
    nr = (1 + rsk_isqrt(1 + 2 * pix)) >> 1;
    jp = pix + 1 - 2 * nr * (nr - 1);
    kshift = 0; jr_eq = 0;
  } else if (pix < npix - ncap) {
    region = 0;
    const int64_t ip = pix - ncap;
    // 32-bit division when the map is small enough (nside <= 8192 always is): a 64-bit one costs ~100 instructions
    const int64_t q = ((uint64_t)ip | (uint64_t)nl4) >> 32 ? ip / nl4 : (int64_t)((uint32_t)ip / (uint32_t)nl4);
    jr_eq = q + nside;
    jp = ip - q * nl4 + 1;
    nr = nside;
    kshift = (jr_eq - nside) & 1;
  } else {
    region = -1;
    const int64_t ip = npix - pix;
    nr = (1 + rsk_isqrt(2 * ip - 1)) >> 1;
    jp = 4 * nr + 1 - (ip - 2 * nr * (nr - 1));
    kshift = 0; jr_eq = 0;
  }
  const double nrf = (double)nr;
  double omz;
  if (region != 0) {
    omz = nrf * nrf / (3.0 * (double)nside * (double)nside);
    z = region > 0 ? 1.0 - omz : omz - 1.0;
  } else {
    z = (double)(2 * nside - jr_eq) * (2.0 / (3.0 * (double)nside));
    omz = 1.0 - fabs(z);
  }
  st = rsk_sqrt(omz * (2.0 - omz));
  phi = ((double)jp - (double)(kshift + 1) * 0.5) * (RSK_HALFPI / nrf);
}

RSK_HD int64_t rsk_spread_bits(int64_t v) {
  uint64_t x = (uint64_t)v & 0xFFFFFFFFull;
  x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
  x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
  x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
  x = (x | (x << 2)) & 0x3333333333333333ull;
  x = (x | (x << 1)) & 0x5555555555555555ull;
  return (int64_t)x;
}

// Pixel containing the direction (z = sin lat, st = cos lat >= 0, phi = lon): the inverse of rsk_hpx_center
// (standard HEALPix ang2pix; astropy_healpix.lonlat_to_healpix at skymodel.py:1916).  st carries the polar
// caps without the cancellation of 1 - |z|.
RSK_HD int64_t rsk_hpx_pixel(int64_t nside, int nested, double z, double st, double phi) {
  const double za = fabs(z);
  double tt = fmod(phi, RSK_D2PI);
  if (tt < 0.0) tt += RSK_D2PI;
  tt *= 0.63661977236758134308;  // in [0, 4]
  if (tt >= 4.0) tt = 0.0;
  const int64_t nl4 = 4 * nside;
  if (za <= 2.0 / 3.0) {  // equatorial belt
    const double temp1 = (double)nside * (0.5 + tt);
    const double temp2 = (double)nside * (z * 0.75);
    const int64_t jp = (int64_t)(temp1 - temp2);  // ascending edge line
    const int64_t jm = (int64_t)(temp1 + temp2);  // descending edge line
    if (nested) {
      const int64_t ifp = jp / nside, ifm = jm / nside;
      const int64_t face = ifp == ifm ? (ifp | 4) : (ifp < ifm ? ifp : ifm + 8);
      const int64_t ix = jm & (nside - 1);
      const int64_t iy = nside - (jp & (nside - 1)) - 1;
      return (face & 15) * nside * nside + rsk_spread_bits(ix) + 2 * rsk_spread_bits(iy);
    }
    const int64_t ir = nside + 1 + jp - jm;  // ring counted from z = 2/3, in [1, 2 nside + 1]
    const int64_t kshift = 1 - (ir & 1);
    int64_t ip = (jp + jm - nside + kshift + 1 + 2 * nl4) >> 1;
    ip %= nl4;
    return 2 * nside * (nside - 1) + (ir - 1) * nl4 + ip;
  }
  // polar caps
  int64_t ntt = (int64_t)tt;
  if (ntt > 3) ntt = 3;
  const double tp = tt - (double)ntt;
  const double tmp = za > 0.99 ? (double)nside * st / sqrt((1.0 + za) / 3.0) : (double)nside * sqrt(3.0 * (1.0 - za));
  int64_t jp = (int64_t)(tp * tmp);
  int64_t jm = (int64_t)((1.0 - tp) * tmp);
  if (nested) {
    if (jp > nside - 1) jp = nside - 1;
    if (jm > nside - 1) jm = nside - 1;
    int64_t face, ix, iy;
    if (z >= 0.0) { face = ntt; ix = nside - jm - 1; iy = nside - jp - 1; }
    else { face = ntt + 8; ix = jp; iy = jm; }
    return face * nside * nside + rsk_spread_bits(ix) + 2 * rsk_spread_bits(iy);
  }
  const int64_t ir = jp + jm + 1;  // ring counted from the closest pole
  int64_t ip = (int64_t)(tt * (double)ir);
  ip %= 4 * ir;
  return z > 0.0 ? 2 * ir * (ir - 1) + ip : 12 * nside * nside - 2 * ir * (ir + 1) + ip;
}

// ---- polarisation basis rotation -------------------------------------------------------------------

// Per-component coefficients of the 2x2 rotation R^T C R of skymodel.py:2704-2710 with
// R = [[cosX, sinX], [-sinX, cosX]]: in Stokes terms
//   I' = n2 I,  V' = n2 V,  Q' = c2 Q - s2 U,  U' = s2 Q + c2 U
// with n2 = cosX^2 + sinX^2 (1 to rounding, kept because the reference's einsum carries it),
// c2 = cosX^2 - sinX^2, s2 = 2 sinX cosX.  The angle is NOT small (it is near pi at the HERA site).
struct RskRot {
  double n2, c2, s2;
};

// (sl, cl, sb, cb): sin/cos of catalogue-frame lon and lat.  (sa, ca, se, ce): sin/cos of az and alt.
// ravg: row-major ICRS->AltAz average rotation (rsk_ravg).
RSK_HD RskRot rsk_basis_rotation(const double* ravg, double sl, double cl, double sb, double cb,
                                 double sa, double ca, double se, double ce) {
  // r_hat(pi/2 - lat, lon) and r_hat(pi/2 - alt, az)   [sct.r_hat, spherical_coords_transforms.py:8-31]
  const double fx = cl * cb, fy = sl * cb, fz = sb;
  const double tx = ca * ce, ty = sa * ce, tz = se;
  // intermediate_vec = R_avg . frame_vec   [skymodel.py:2595]
  const double ix = ravg[0] * fx + ravg[1] * fy + ravg[2] * fz;
  const double iy = ravg[3] * fx + ravg[4] * fy + ravg[5] * fz;
  const double iz = ravg[6] * fx + ravg[7] * fy + ravg[8] * fz;
  // vecs2rot(i, t): Rodrigues about n = i x t.  With K = [n_hat]x, sin(psi) K = [n]x and
  // (1 - cos psi) K^2 = [n]x^2 / (1 + cos psi): no normalisation, no atan2, finite at psi = 0
  // (where upstream divides by zero and raises, spherical_coords_transforms.py:300-303).
  const double nx = iy * tz - iz * ty;
  const double ny = iz * tx - ix * tz;
  const double nz = ix * ty - iy * tx;
  const double k = 1.0 / (1.0 + (ix * tx + iy * ty + iz * tz));
  // theta_hat(beta, alpha) at the AltAz point, beta = pi/2 - alt: (cos az sin alt, sin az sin alt, -cos alt)
  const double wx = ca * se, wy = sa * se, wz = -ce;
  // R_perturb^T w = w - n x w + n x (n x w) / (1 + c)
  const double c1x = ny * wz - nz * wy;
  const double c1y = nz * wx - nx * wz;
  const double c1z = nx * wy - ny * wx;
  const double c2x = ny * c1z - nz * c1y;
  const double c2y = nz * c1x - nx * c1z;
  const double c2z = nx * c1y - ny * c1x;
  const double ux = wx - c1x + k * c2x;
  const double uy = wy - c1y + k * c2y;
  const double uz = wz - c1z + k * c2z;
  // bh = R_exact^T w = R_avg^T (R_perturb^T w)   [sct :180]
  const double bx = ravg[0] * ux + ravg[3] * uy + ravg[6] * uz;
  const double by = ravg[1] * ux + ravg[4] * uy + ravg[7] * uz;
  const double bz = ravg[2] * ux + ravg[5] * uy + ravg[8] * uz;
  // cosX = bh . theta_hat(theta, phi), sinX = bh . phi_hat(theta, phi)   [sct :182-183]
  const double cosx = bx * (cl * sb) + by * (sl * sb) - bz * cb;
  const double sinx = by * cl - bx * sl;
  RskRot r;
  r.n2 = cosx * cosx + sinx * sinx;
  r.c2 = cosx * cosx - sinx * sinx;
  r.s2 = 2.0 * sinx * cosx;
  return r;
}
