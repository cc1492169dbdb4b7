// rsk_host.cpp -- host-side (CPU, once per time step) pieces of librsk: the astrometry context
// builder, frame matrices, R_avg, context packing.  No GPU work here.
//
// Reference call sites replaced (RadioAstronomySoftwareGroup/pyradiosky):
//   rsk_astrom_context   astropy ErfaAstrom.apco(AltAz(obstime, location)) -> erfa.apco, reached from
//                        src/pyradiosky/skymodel.py:2451-2454 and :2518-2527
//   rsk_frame_to_icrs    astropy frame graph legs FK5/Galactic -> ICRS used by the same transform
//   rsk_ravg             SkyModel._calc_average_rotation_matrix, skymodel.py:2499-2555
// The ERFA routines named in comments (eraApco, eraApcs, eraPvtob, eraC2ixys, eraEra00, eraSp00,
// eraGd2gce, eraPom00, eraPfw06/eraFw2m) are restated from their published algorithms; liberfa is a
// third-party dependency of the reference (via astropy) that is not available in this image.
#include <math.h>
#include <string.h>

#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <stdint.h>
#include <stdlib.h>

#include "../../include/rsk.h"
#include "rsk_device.cuh"

namespace {

const double DAS2R = 4.848136811095359935899141e-6;
const double DJ00 = 2451545.0;
const double DJY = 365.25;
const double DJC = 36525.0;
const double DAYSEC = 86400.0;
const double DAU = 149597870.7e3;
const double AULT = DAU / 299792458.0;  // erfam.h: ERFA_AULT = ERFA_DAU / ERFA_CMPS
const double WGS84_A = 6378137.0;
const double WGS84_F = 1.0 / 298.257223563;
const double DEG = RSK_DPI / 180.0;

// astrom[31] slots (eraASTROM order)
enum {
  A_PMT = 0, A_EB = 1, A_EH = 4, A_EM = 7, A_V = 8, A_BM1 = 11, A_BPN = 12, A_ALONG = 21, A_PHI = 22,
  A_XPL = 23, A_YPL = 24, A_SPHI = 25, A_CPHI = 26, A_DIURAB = 27, A_ERAL = 28, A_REFA = 29, A_REFB = 30
};

struct M3 {
  double a[9];
};

M3 ident() {
  M3 r;
  memset(r.a, 0, sizeof(r.a));
  r.a[0] = r.a[4] = r.a[8] = 1.0;
  return r;
}
M3 mul(const M3& x, const M3& y) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.a[3 * i + j] = x.a[3 * i] * y.a[j] + x.a[3 * i + 1] * y.a[3 + j] + x.a[3 * i + 2] * y.a[6 + j];
  return r;
}
M3 transpose(const M3& x) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.a[3 * i + j] = x.a[3 * j + i];
  return r;
}
void mulv(const M3& x, const double* v, double* o) {
  for (int i = 0; i < 3; ++i) o[i] = x.a[3 * i] * v[0] + x.a[3 * i + 1] * v[1] + x.a[3 * i + 2] * v[2];
}
// ERFA-style passive rotations
M3 rotx(double p) {
  M3 r = ident();
  const double s = sin(p), c = cos(p);
  r.a[4] = c; r.a[5] = s; r.a[7] = -s; r.a[8] = c;
  return r;
}
M3 roty(double p) {
  M3 r = ident();
  const double s = sin(p), c = cos(p);
  r.a[0] = c; r.a[2] = -s; r.a[6] = s; r.a[8] = c;
  return r;
}
M3 rotz(double p) {
  M3 r = ident();
  const double s = sin(p), c = cos(p);
  r.a[0] = c; r.a[1] = s; r.a[3] = -s; r.a[4] = c;
  return r;
}

double anp(double a) {
  double w = fmod(a, RSK_D2PI);
  if (w < 0) w += RSK_D2PI;
  return w;
}
double anpm(double a) {
  double w = fmod(a, RSK_D2PI);
  if (fabs(w) >= RSK_DPI) w -= copysign(RSK_D2PI, a);
  return w;
}

// TAI-UTC (s) since 1972, IERS Bulletin C (eraDat restated as a table)
double tai_minus_utc(double jd_utc) {
  static const struct { int mjd; int dat; } leap[] = {
      {41317, 10}, {41499, 11}, {41683, 12}, {42048, 13}, {42413, 14}, {42778, 15}, {43144, 16},
      {43509, 17}, {43874, 18}, {44239, 19}, {44786, 20}, {45151, 21}, {45516, 22}, {46247, 23},
      {47161, 24}, {47892, 25}, {48257, 26}, {48804, 27}, {49169, 28}, {49534, 29}, {50083, 30},
      {50630, 31}, {51179, 32}, {53736, 33}, {54832, 34}, {56109, 35}, {57204, 36}, {57754, 37}};
  const double mjd = jd_utc - 2400000.5;
  int dat = 10;
  for (size_t i = 0; i < sizeof(leap) / sizeof(leap[0]); ++i)
    if (mjd >= leap[i].mjd) dat = leap[i].dat;
  return (double)dat;
}

// eraEra00
double era00(double dj1, double dj2) {
  double d1 = dj1, d2 = dj2;
  if (dj1 >= dj2) { d1 = dj2; d2 = dj1; }
  const double t = d1 + (d2 - DJ00);
  const double f = fmod(d1, 1.0) + fmod(d2, 1.0);
  return anp(RSK_D2PI * (f + 0.7790572732640 + 0.00273781191135448 * t));
}

// CIP X, Y: leading terms of the IAU 2006/2000A series; s ~ -XY/2.  t = TT Julian centuries.
void cip_xys(double t, double* x, double* y, double* s) {
  const double om = (450160.398036 - 6962890.5431 * t) * DAS2R;
  const double lp = (1287104.79305 + 129596581.0481 * t) * DAS2R;
  const double f = (335779.526232 + 1739527262.8478 * t) * DAS2R;
  const double d = (1072260.70369 + 1602961601.2090 * t) * DAS2R;
  double xx = 2004.191898 * t - 0.4297829 * t * t - 6.844318 * sin(om) - 0.523908 * sin(2 * (f - d + om)) -
              0.090552 * sin(2 * (f + om)) + 0.082169 * sin(2 * om) + 0.058707 * sin(lp);
  double yy = -0.006951 - 0.025896 * t - 22.4072747 * t * t + 9.205236 * cos(om) +
              0.573033 * cos(2 * (f - d + om)) + 0.097847 * cos(2 * (f + om)) - 0.089618 * cos(2 * om);
  *x = xx * DAS2R;
  *y = yy * DAS2R;
  *s = -(*x) * (*y) / 2.0;
}

// Geocentric Sun, low-precision theory of the Astronomical Almanac, referred to the J2000 equinox (au)
void sun_j2000(double jd_tt, double* p) {
  const double n = jd_tt - DJ00;
  const double L = (280.460 + 0.9856474 * n) * DEG;
  const double g = (357.528 + 0.9856003 * n) * DEG;
  double lam = L + (1.915 * sin(g) + 0.020 * sin(2 * g)) * DEG;
  lam -= (1.396971 * (n / DJC)) * DEG;
  const double R = 1.00014 - 0.01671 * cos(g) - 0.00014 * cos(2 * g);
  const double eps = 23.4392911 * DEG;
  p[0] = R * cos(lam);
  p[1] = R * cos(eps) * sin(lam);
  p[2] = R * sin(eps) * sin(lam);
}

// eraC2ixys
M3 c2ixys(double x, double y, double s) {
  const double r2 = x * x + y * y;
  const double e = (r2 > 0.0) ? atan2(y, x) : 0.0;
  const double d = atan(sqrt(r2 / (1.0 - r2)));
  return mul(rotz(-(e + s)), mul(roty(d), rotz(e)));
}

void gd2gc(double elong, double phi, double height, double* xyz) {
  const double sp = sin(phi), cp = cos(phi);
  double w = 1.0 - WGS84_F;
  w *= w;
  const double d = cp * cp + w * sp * sp;
  const double ac = WGS84_A / sqrt(d);
  const double as = w * ac;
  const double r = (ac + height) * cp;
  xyz[0] = r * cos(elong);
  xyz[1] = r * sin(elong);
  xyz[2] = (as + height) * sp;
}

// IAU 2006 precession (Fukushima-Williams angles, eraPfw06 + eraFw2m): GCRS -> mean of date at TT
// Julian centuries t; dividing out the t = 0 frame bias leaves the precession-only matrix eraBp06's rp.
M3 fw_matrix(double t) {
  const double gamb = (-0.052928 + (10.556378 + (0.4932044 + (-0.00031238 + (-0.000002788 + 0.0000000260 * t) * t) * t) * t) * t) * DAS2R;
  const double phib = (84381.412819 + (-46.811016 + (0.0511268 + (0.00053289 + (-0.000000440 - 0.0000000176 * t) * t) * t) * t) * t) * DAS2R;
  const double psib = (-0.041775 + (5038.481484 + (1.5584175 + (-0.00018522 + (-0.000026452 - 0.0000000148 * t) * t) * t) * t) * t) * DAS2R;
  const double epsa = (84381.406 + (-46.836769 + (-0.0001831 + (0.00200340 + (-0.000000576 - 0.0000000434 * t) * t) * t) * t) * t) * DAS2R;
  return mul(rotx(-epsa), mul(rotz(-psib), mul(rotx(phib), rotz(gamb))));
}
M3 precession_from_j2000(double t) { return mul(fw_matrix(t), transpose(fw_matrix(0.0))); }

// astropy builtin_frames.icrs_fk5_transforms._icrs_to_fk5_matrix
M3 icrs_to_fk5_j2000() {
  const double eta0 = -19.9 / 3600000.0 * DEG, xi0 = 9.1 / 3600000.0 * DEG, da0 = -22.9 / 3600000.0 * DEG;
  return mul(rotx(-eta0), mul(roty(xi0), rotz(da0)));
}

// orthogonal polar factor by Newton iteration X <- (X + X^-T) / 2 (quadratic; input is orthogonal to ~1e-7)
M3 polar_factor(M3 x) {
  for (int it = 0; it < 8; ++it) {
    const double* a = x.a;
    const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    M3 cof;  // cofactor matrix = det * X^-T
    cof.a[0] = a[4] * a[8] - a[5] * a[7];
    cof.a[1] = a[5] * a[6] - a[3] * a[8];
    cof.a[2] = a[3] * a[7] - a[4] * a[6];
    cof.a[3] = a[2] * a[7] - a[1] * a[8];
    cof.a[4] = a[0] * a[8] - a[2] * a[6];
    cof.a[5] = a[1] * a[6] - a[0] * a[7];
    cof.a[6] = a[1] * a[5] - a[2] * a[4];
    cof.a[7] = a[2] * a[3] - a[0] * a[5];
    cof.a[8] = a[0] * a[4] - a[1] * a[3];
    M3 nx;
    double delta = 0.0;
    for (int i = 0; i < 9; ++i) {
      nx.a[i] = 0.5 * (a[i] + cof.a[i] / det);
      delta = fmax(delta, fabs(nx.a[i] - a[i]));
    }
    x = nx;
    if (delta < 1e-17) break;
  }
  return x;
}

}  // namespace

extern "C" {

int rsk_version(void) { return RSK_VERSION; }

const char* rsk_strerror(int code) {
  switch (code) {
    case RSK_OK: return "success";
    case RSK_ERR_NULL: return "a required pointer argument is NULL";
    case RSK_ERR_SIZE: return "negative or inconsistent size argument";
    case RSK_ERR_MODE: return "unknown spectral mode or enum value";
    case RSK_ERR_LD: return "ld_out is smaller than col0 + ncols";
    case RSK_ERR_WORKSPACE: return "workspace missing or too small";
    case RSK_ERR_UNSUPPORTED: return "unsupported configuration (e.g. spline degree > 3)";
    case RSK_ERR_NSIDE: return "invalid HEALPix nside";
    default: return code > 0 ? "CUDA runtime error (code is a cudaError_t)" : "unknown error";
  }
}

int rsk_geodetic_to_geocentric(double elong, double phi, double height, double* xyz) {
  if (!xyz) return RSK_ERR_NULL;
  gd2gc(elong, phi, height, xyz);
  return RSK_OK;
}

int rsk_astrom_context(double jd_utc1, double jd_utc2, double dut1, double elong, double phi, double height,
                       double xp, double yp, double* astrom) {
  if (!astrom) return RSK_ERR_NULL;
  memset(astrom, 0, RSK_ASTROM_LEN * sizeof(double));
  const double dtai = tai_minus_utc(jd_utc1 + jd_utc2);
  const double tt1 = jd_utc1, tt2 = jd_utc2 + (dtai + 32.184) / DAYSEC;
  const double ut11 = jd_utc1, ut12 = jd_utc2 + dut1 / DAYSEC;
  const double t = ((tt1 - DJ00) + tt2) / DJC;
  double x, y, s;
  cip_xys(t, &x, &y, &s);
  const double theta = era00(ut11, ut12);
  const double sp = -47e-6 * t * DAS2R;  // eraSp00
  // Earth: heliocentric = barycentric at this accuracy; velocity by central difference (au/day)
  double sun0[3], sunp[3], sunm[3], epos[3], evel[3];
  const double h = 0.01, jd_tt = tt1 + tt2;
  sun_j2000(jd_tt, sun0);
  sun_j2000(jd_tt + h, sunp);
  sun_j2000(jd_tt - h, sunm);
  for (int i = 0; i < 3; ++i) {
    epos[i] = -sun0[i];
    evel[i] = -(sunp[i] - sunm[i]) / (2 * h);
  }
  // eraApco: CIRS -> apparent [HA, Dec] rotation and what is solved from it
  const M3 r = mul(rotz(elong), mul(rotx(-yp), mul(roty(-xp), rotz(theta + sp))));
  double a = r.a[0], b = r.a[1];
  const double eral = (a != 0.0 || b != 0.0) ? atan2(b, a) : 0.0;
  const double c = r.a[2];
  const double xpl = atan2(c, sqrt(a * a + b * b));
  a = r.a[5]; b = r.a[8];
  const double ypl = (a != 0.0 || b != 0.0) ? -atan2(a, b) : 0.0;
  const M3 rc2i = c2ixys(x, y, s);
  // eraPvtob: observer position / velocity in CIRS (m, m/s)
  const double om = 1.00273781191135448 * RSK_D2PI / DAYSEC;
  double xyzm[3], xyz[3];
  gd2gc(elong, phi, height, xyzm);
  const M3 rpm = mul(rotx(-yp), mul(roty(-xp), rotz(sp)));  // eraPom00
  mulv(transpose(rpm), xyzm, xyz);
  const double st = sin(theta), ct = cos(theta);
  const double pc[3] = {ct * xyz[0] - st * xyz[1], st * xyz[0] + ct * xyz[1], xyz[2]};
  const double vc[3] = {om * (-st * xyz[0] - ct * xyz[1]), om * (ct * xyz[0] - st * xyz[1]), 0.0};
  double pg[3], vg[3];  // eraTrxpv: into GCRS
  const M3 rt = transpose(rc2i);
  mulv(rt, pc, pg);
  mulv(rt, vc, vg);
  // eraApcs
  const double audms = DAU / DAYSEC, cr = AULT / DAYSEC;
  astrom[A_PMT] = ((tt1 - DJ00) + tt2) / DJY;
  double ph[3], v2 = 0.0, em2 = 0.0;
  for (int i = 0; i < 3; ++i) {
    const double dp = pg[i] / DAU, dv = vg[i] / audms;
    astrom[A_EB + i] = epos[i] + dp;
    ph[i] = epos[i] + dp;
    em2 += ph[i] * ph[i];
    const double w = (evel[i] + dv) * cr;
    astrom[A_V + i] = w;
    v2 += w * w;
  }
  const double em = sqrt(em2);
  astrom[A_EM] = em;
  for (int i = 0; i < 3; ++i) astrom[A_EH + i] = ph[i] / em;
  astrom[A_BM1] = sqrt(1.0 - v2);
  for (int i = 0; i < 9; ++i) astrom[A_BPN + i] = rc2i.a[i];
  astrom[A_ALONG] = anpm(eral - theta);
  astrom[A_PHI] = phi;
  astrom[A_XPL] = xpl;
  astrom[A_YPL] = ypl;
  astrom[A_SPHI] = sin(phi);
  astrom[A_CPHI] = cos(phi);
  astrom[A_DIURAB] = 0.0;  // eraApco disables the (redundant) diurnal-aberration step
  astrom[A_ERAL] = eral;
  astrom[A_REFA] = 0.0;    // AltAz default pressure 0 -> no refraction (skymodel.py:2453)
  astrom[A_REFB] = 0.0;
  return RSK_OK;
}

int rsk_astrom_contexts(const double* jd_utc1, const double* jd_utc2, int64_t T, double dut1, double elong, double phi,
                        double height, double xp, double yp, double* astroms) {
  if (T < 0) return RSK_ERR_SIZE;
  if (T > 0 && (!jd_utc1 || !jd_utc2 || !astroms)) return RSK_ERR_NULL;
  for (int64_t t = 0; t < T; ++t) {
    const int rc = rsk_astrom_context(jd_utc1[t], jd_utc2[t], dut1, elong, phi, height, xp, yp, astroms + t * RSK_ASTROM_LEN);
    if (rc != RSK_OK) return rc;
  }
  return RSK_OK;
}

int rsk_frame_to_icrs(int frame, double equinox_jyear, double* m) {
  if (!m) return RSK_ERR_NULL;
  M3 to_frame;  // ICRS -> frame
  if (frame == RSK_FRAME_ICRS) {
    to_frame = ident();
  } else if (frame == RSK_FRAME_FK5) {
    to_frame = icrs_to_fk5_j2000();
    if (equinox_jyear != 2000.0 && equinox_jyear != 0.0)
      to_frame = mul(precession_from_j2000((equinox_jyear - 2000.0) / 100.0), to_frame);
  } else if (frame == RSK_FRAME_GALACTIC) {
    // astropy Galactic: FK5(J2000) -> Galactic, NGP (192.8594812065348, 27.12825118085622) deg, lon0 122.9319185680026 deg
    const M3 g = mul(rotz((180.0 - 122.9319185680026) * DEG), mul(roty((90.0 - 27.12825118085622) * DEG), rotz(192.8594812065348 * DEG)));
    to_frame = mul(g, icrs_to_fk5_j2000());
  } else {
    return RSK_ERR_MODE;
  }
  const M3 t = transpose(to_frame);
  memcpy(m, t.a, sizeof(t.a));
  return RSK_OK;
}

int rsk_pack_context(const double* astrom, const double* frame2icrs, double* ctx_out) {
  if (!astrom || !ctx_out) return RSK_ERR_NULL;
  RskCtx c;
  for (int i = 0; i < 3; ++i) {
    c.eh[i] = astrom[A_EH + i];
    c.v[i] = astrom[A_V + i];
  }
  c.em = astrom[A_EM];
  c.bm1 = astrom[A_BM1];
  M3 bpn;
  memcpy(bpn.a, astrom + A_BPN, sizeof(bpn.a));
  // eraAtioq: v = s2c(ri - eral, di) == Rz(eral) . pi ; then the polar-motion rotation
  const double sx = sin(astrom[A_XPL]), cx = cos(astrom[A_XPL]);
  const double sy = sin(astrom[A_YPL]), cy = cos(astrom[A_YPL]);
  M3 pm;
  pm.a[0] = cx;       pm.a[1] = 0.0; pm.a[2] = sx;
  pm.a[3] = sx * sy;  pm.a[4] = cy;  pm.a[5] = -cx * sy;
  pm.a[6] = -sx * cy; pm.a[7] = sy;  pm.a[8] = cx * cy;
  M3 a1 = mul(pm, mul(rotz(astrom[A_ERAL]), bpn));
  c.sphi = astrom[A_SPHI];
  c.cphi = astrom[A_CPHI];
  c.diurab = astrom[A_DIURAB];
  c.refa = astrom[A_REFA];
  c.refb = astrom[A_REFB];
  if (c.diurab == 0.0 && c.refa == 0.0 && c.refb == 0.0) {
    // Nothing non-linear sits between the polar-motion step and the latitude rotation (eraApco sets diurab
    // = 0; AltAz without pressure gives refa = refb = 0): fold the latitude rotation into a1 and mark the
    // context as folded (rsk_ctx_folded) so the kernels skip the identity steps.
    M3 rl;
    rl.a[0] = c.sphi; rl.a[1] = 0.0; rl.a[2] = -c.cphi;
    rl.a[3] = 0.0;    rl.a[4] = 1.0; rl.a[5] = 0.0;
    rl.a[6] = c.cphi; rl.a[7] = 0.0; rl.a[8] = c.sphi;
    a1 = mul(rl, a1);
    c.sphi = 1.0;
    c.cphi = 0.0;
  }
  memcpy(c.a1, a1.a, sizeof(a1.a));
  c.w2 = RSK_SRS / astrom[A_EM];
  bool is_ident = true;
  const M3 id = ident();
  if (frame2icrs)
    for (int i = 0; i < 9; ++i) is_ident = is_ident && (frame2icrs[i] == id.a[i]);
  c.m_ident = is_ident ? 1.0 : 0.0;
  c.inv1pbm1 = 1.0 / (1.0 + astrom[A_BM1]);
  memcpy(c.m, is_ident ? id.a : frame2icrs, sizeof(c.m));
  memcpy(ctx_out, &c, sizeof(c));
  return RSK_OK;
}

int rsk_pack_step(const double* astrom, const double* frame2icrs, int with_ravg, double* step) {
  if (!astrom || !step) return RSK_ERR_NULL;
  int rc = rsk_pack_context(astrom, frame2icrs, step);
  if (rc != RSK_OK) return rc;
  if (with_ravg) {
    rc = rsk_ravg(astrom, step + RSK_CTX_LEN);
    if (rc != RSK_OK) return rc;
  } else {
    for (int i = 0; i < 9; ++i) step[RSK_CTX_LEN + i] = 0.0;
  }
  step[RSK_CTX_LEN + 9] = with_ravg ? 1.0 : 0.0;
  return RSK_OK;
}

int rsk_pack_contexts(const double* astroms, int64_t T, const double* frame2icrs, double* ctx) {
  if (T < 0) return RSK_ERR_SIZE;
  if (T > 0 && (!astroms || !ctx)) return RSK_ERR_NULL;
  for (int64_t t = 0; t < T; ++t) {
    const int rc = rsk_pack_context(astroms + t * RSK_ASTROM_LEN, frame2icrs, ctx + t * RSK_CTX_LEN);
    if (rc != RSK_OK) return rc;
  }
  return RSK_OK;
}

int rsk_ravg(const double* astrom, double* ravg) {
  if (!astrom || !ravg) return RSK_ERR_NULL;
  RskCtx c;
  rsk_pack_context(astrom, nullptr, reinterpret_cast<double*>(&c));
  // skymodel.py:2513-2527: the three ICRS unit axes through the source transform; column k of
  // R_screwy is the AltAz cartesian (x north, y east, z up) of axis k.
  M3 screwy;
  for (int k = 0; k < 3; ++k) {
    const RskAltAz o = rsk_chain(c, k == 0 ? 1.0 : 0.0, k == 1 ? 1.0 : 0.0, k == 2 ? 1.0 : 0.0);
    screwy.a[0 + k] = o.m;  // cos(alt) cos(az)
    screwy.a[3 + k] = o.l;  // cos(alt) sin(az)
    screwy.a[6 + k] = o.n;  // sin(alt)
  }
  // scipy.linalg.orthogonal_procrustes(R_screwy, I)[0].T == orthogonal polar factor of R_screwy
  const M3 r = polar_factor(screwy);
  memcpy(ravg, r.a, sizeof(r.a));
  return RSK_OK;
}

// Host-side unpack of the Hermitian-packed transfer format (see rsk.h): pure data movement, split over
// threads by frequency row, non-temporal stores so the output does not pass through the caches.  With AVX-512
// (chosen at run time) a whole 64-byte line leaves per store: measured 1.2x the 16-byte SSE2 form on one core,
// and no change with 8+ threads, where the expansion is bound by the host's memory system, not the store width
// (profiles/r01_summary.md).
}  // extern "C"

namespace {

// one output row: out[c] = (a[c], sign * b[c]) as complex128, b == nullptr -> imaginary part 0
void expand_row_sse2(const double* a, const double* b, double sign, double* out, int64_t ncols) {
#if defined(__SSE2__)
  if ((reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
    if (b) {
      for (int64_t c = 0; c < ncols; ++c) _mm_stream_pd(out + 2 * c, _mm_set_pd(sign * b[c], a[c]));
    } else {
      for (int64_t c = 0; c < ncols; ++c) _mm_stream_pd(out + 2 * c, _mm_set_pd(0.0, a[c]));
    }
    return;
  }
#endif
  for (int64_t c = 0; c < ncols; ++c) {
    out[2 * c] = a[c];
    out[2 * c + 1] = b ? sign * b[c] : 0.0;
  }
}

#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx512f"))) void expand_row_avx512(const double* a, const double* b, double sign, double* out,
                                                           int64_t ncols) {
  int64_t c = 0;
  // peel until the destination is 64-byte aligned (it is 16-byte aligned: complex128 elements)
  while (c < ncols && (reinterpret_cast<uintptr_t>(out + 2 * c) & 63u) != 0) {
    out[2 * c] = a[c];
    out[2 * c + 1] = b ? sign * b[c] : 0.0;
    ++c;
  }
  const __m512d sgn = _mm512_set1_pd(sign);
  for (; c + 4 <= ncols; c += 4) {
    const __m512d va = _mm512_castpd256_pd512(_mm256_loadu_pd(a + c));
    __m512d v = _mm512_maskz_expand_pd(0x55, va);  // a0 0 a1 0 a2 0 a3 0
    if (b) {
      const __m512d vb = _mm512_mul_pd(_mm512_castpd256_pd512(_mm256_loadu_pd(b + c)), sgn);
      v = _mm512_mask_expand_pd(v, 0xAA, vb);      // a0 b0 a1 b1 ...
    }
    _mm512_stream_pd(out + 2 * c, v);
  }
  for (; c < ncols; ++c) {
    out[2 * c] = a[c];
    out[2 * c + 1] = b ? sign * b[c] : 0.0;
  }
}
#endif

typedef void (*ExpandRow)(const double*, const double*, double, double*, int64_t);

ExpandRow pick_expand_row() {
#if defined(__x86_64__) && defined(__GNUC__)
  if (__builtin_cpu_supports("avx512f") && !getenv("RSK_NO_AVX512")) return expand_row_avx512;
#endif
  return expand_row_sse2;
}

}  // namespace

extern "C" {

int rsk_unpack_coherency_host(const double* packed, int64_t nf, int64_t ncols, int64_t ld_packed, double* out,
                              int64_t nf_out, int64_t f0_out, int64_t ld_out, int nthreads) {
  if (nf < 0 || ncols < 0 || ld_packed < ncols || ld_out < ncols || f0_out < 0 || nf_out < f0_out + nf) return RSK_ERR_SIZE;
  if (nf == 0 || ncols == 0) return RSK_OK;
  if (!packed || !out) return RSK_ERR_NULL;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if ((int64_t)nthreads > nf) nthreads = (int)nf;
  const ExpandRow expand = pick_expand_row();
  auto work = [=](int t) {
    const int64_t r0 = nf * t / nthreads, r1 = nf * (t + 1) / nthreads;
    constexpr int64_t kBlock = 2048;  // columns per pass: the four input pieces stay in L1 while four rows are written
    for (int64_t r = r0; r < r1; ++r) {
      const double* h00 = packed + (0 * nf + r) * ld_packed;
      const double* h11 = packed + (1 * nf + r) * ld_packed;
      const double* re = packed + (2 * nf + r) * ld_packed;
      const double* im = packed + (3 * nf + r) * ld_packed;
      double* o00 = out + 2 * ((0 * nf_out + f0_out + r) * ld_out);
      double* o01 = out + 2 * ((1 * nf_out + f0_out + r) * ld_out);
      double* o10 = out + 2 * ((2 * nf_out + f0_out + r) * ld_out);
      double* o11 = out + 2 * ((3 * nf_out + f0_out + r) * ld_out);
      for (int64_t c0 = 0; c0 < ncols; c0 += kBlock) {
        const int64_t nb = ncols - c0 < kBlock ? ncols - c0 : kBlock;
        expand(h00 + c0, nullptr, 0.0, o00 + 2 * c0, nb);
        expand(re + c0, im + c0, -1.0, o01 + 2 * c0, nb);
        expand(re + c0, im + c0, 1.0, o10 + 2 * c0, nb);
        expand(h11 + c0, nullptr, 0.0, o11 + 2 * c0, nb);
      }
    }
#if defined(__SSE2__)
    _mm_sfence();
#endif
  };
  if (nthreads == 1) {
    work(0);
    return RSK_OK;
  }
  std::vector<std::thread> pool;
  pool.reserve(nthreads - 1);
  for (int t = 1; t < nthreads; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  return RSK_OK;
}

size_t rsk_workspace_bytes(int op, int64_t n, int64_t nf) {
  (void)nf;
  if (op == RSK_OP_COMPACT) {
    if (n < 0) return 0;
    const int64_t tiles = (n + 2047) / 2048;
    return (size_t)(tiles + 1) * sizeof(int64_t);
  }
  return 0;
}

}  // extern "C"
