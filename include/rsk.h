/*
 * rsk.h -- C ABI of librsk, the B200 (sm_100a) implementation of pyradiosky's per-component
 * evaluation hot path.
 *
 * The reference (RadioAstronomySoftwareGroup/pyradiosky) is pure Python and has no FFI; its boundary
 * for this path is the method surface of ``SkyModel`` (src/pyradiosky/skymodel.py) and two functions
 * of src/pyradiosky/utils.py.  Each entry point below names the reference code it replaces; the
 * ctypes stub a pyradiosky maintainer would add is shown in INTEGRATION.md and shipped as
 * pyradiosky_b200/_lib.py.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes only.  Unless marked HOST, every pointer is a DEVICE pointer
 *    on the current CUDA device (the caller selects the device, e.g. with torch.cuda.device).
 *  - The caller owns every buffer.  The library never allocates or frees device memory and keeps no
 *    global state.  Workspaces are caller-provided and sized by rsk_workspace_bytes().
 *  - All functions that launch work are asynchronous on ``stream`` (a cudaStream_t passed as void*;
 *    NULL = the legacy default stream) and never synchronise.
 *  - Return value: 0 = success; negative = argument error (RSK_ERR_*); positive = the cudaError_t of
 *    the failed launch.  No exceptions, no abort().
 *  - Angles are radians, frequencies Hz, all floating point is IEEE double.  Complex outputs are
 *    interleaved (re, im) doubles, i.e. numpy complex128.
 *  - Array layouts are the reference's: Stokes (4, Nfreqs, Ncomponents) and coherency
 *    (2, 2, Nfreqs, Ncomponents), C order, component axis fastest (skymodel.py:614-619, :2664-2667).
 */
#ifndef RSK_H
#define RSK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSK_VERSION 200 /* 0.2.0: RskSpectrum gained group_of / group / stokes_cm; rsk_count, rsk_copy_2d, rsk_stokes_component_major; rsk_valid_channels writes multi-word masks */

/* Length of the astrometry context: exactly ERFA's eraASTROM / erfa.dt_eraASTROM, as doubles:
 * pmt, eb[3], eh[3], em, v[3], bm1, bpn[3][3] (row-major), along, phi, xpl, ypl, sphi, cphi, diurab,
 * eral, refa, refb  ->  1+3+3+1+3+1+9+10 = 31.  A host with astropy can pass
 * ``erfa_astrom.get().apco(AltAz(obstime=..., location=...))`` bytes straight through. */
#define RSK_ASTROM_LEN 31
/* Length of the device-ready context produced by rsk_pack_context(). */
#define RSK_CTX_LEN 34
/* Length of the per-time-step device record of rsk_pack_step() / rsk_positions_step(): packed context, R_avg[9],
 * one flag (R_avg valid). */
#define RSK_STEP_LEN 44

/* error codes */
#define RSK_OK 0
#define RSK_ERR_NULL (-1)        /* a required pointer is NULL */
#define RSK_ERR_SIZE (-2)        /* negative or inconsistent size */
#define RSK_ERR_MODE (-3)        /* unknown spectral mode / bad enum value */
#define RSK_ERR_LD (-4)          /* ld_out < col0 + ncols */
#define RSK_ERR_WORKSPACE (-5)   /* workspace missing or too small */
#define RSK_ERR_UNSUPPORTED (-6) /* e.g. spline degree > 3 */
#define RSK_ERR_NSIDE (-7)       /* nside not a positive power of two (nested) / not positive (ring) */

/* spectral modes = SkyModel.spectral_type (skymodel.py:560-577) */
#define RSK_FLAT 0
#define RSK_SPECTRAL_INDEX 1
#define RSK_SUBBAND 2
#define RSK_FULL 3

/* catalogue frames understood by rsk_frame_to_icrs() */
#define RSK_FRAME_ICRS 0
#define RSK_FRAME_FK5 1      /* equinox given as a Julian epoch, e.g. 2000.0 */
#define RSK_FRAME_GALACTIC 2

/* workspace selectors for rsk_workspace_bytes() */
#define RSK_OP_COMPACT 1

/* Where component coordinates come from.
 * Point catalogues: lon/lat arrays in the catalogue frame (SkyModel.skycoord, skymodel.py:2449).
 * HEALPix maps (nside > 0): pixel centres are generated on device, replacing
 * astropy_healpix.HEALPix(nside, order, frame).healpix_to_skycoord(hpx_inds) (skymodel.py:1656-1669);
 * hpx_inds == NULL means component i is pixel hpx_first + i (full map or a contiguous shard). */
typedef struct RskCoords {
  const double* lon;       /* [n] or NULL when nside > 0 */
  const double* lat;       /* [n] or NULL when nside > 0 */
  int64_t nside;           /* 0 = point catalogue */
  int32_t nested;          /* hpx_order: 0 = ring, 1 = nested (skymodel.py:545-552) */
  int32_t reserved;
  const int64_t* hpx_inds; /* [n] or NULL */
  int64_t hpx_first;
  /* Optional per-component polarisation-rotation coefficients [3][n] = (cos^2 X + sin^2 X,
   * cos^2 X - sin^2 X, 2 sin X cos X) of skymodel.py:2603-2646.  rsk_positions writes them (it already has
   * every sine and cosine they need) when this pointer and its ravg argument are non-NULL; rsk_coherency
   * reads them instead of recomputing them when the pointer is non-NULL. */
  double* rot;
} RskCoords;

/* The spectral model of the catalogue = what SkyModel.at_frequencies (skymodel.py:2084-2389) reads. */
typedef struct RskSpectrum {
  int32_t mode;                 /* RSK_FLAT | RSK_SPECTRAL_INDEX | RSK_SUBBAND | RSK_FULL */
  int32_t degree;               /* SUBBAND: degree (0..3) of the piecewise-polynomial operator */
  const double* stokes;         /* [4][nf_in][n] */
  int64_t nf_in;                /* stored channels; 1 for FLAT / SPECTRAL_INDEX */
  int64_t n;                    /* components in the model (leading dimension of stokes) */
  const double* spectral_index; /* SPECTRAL_INDEX: [n] */
  const double* ref_freq;       /* SPECTRAL_INDEX: [n] Hz (SkyModel.reference_frequency) */
  /* SUBBAND: scipy.interpolate.interp1d(freq_array, stokes, axis=1, kind) (skymodel.py:2362-2368) is
   * linear in the data with knots shared by all components, so the host converts it once into a
   * piecewise-polynomial operator: on break interval b, with u = (f - breaks[b]) / (breaks[b+1] -
   * breaks[b]), value = sum_p u^p * sum_j ppoly[b][p][j] * stokes[s][j][comp]. */
  const double* breaks;         /* [nb + 1] ascending Hz */
  const double* ppoly;          /* [nb][degree + 1][nf_in] */
  int64_t nb;
  /* FULL: channel f of the request reads stored channel freq_map[f] (subset selection of
   * skymodel.py:2165-2181); NULL = identity (requires nf == nf_in). */
  const int32_t* freq_map;      /* [nf] or NULL */
  /* polarised flag per component = membership of SkyModel._polarized (skymodel.py:1037-1042, frozen
   * at construction upstream); NULL = nothing is rotated.  Build it with rsk_polarized_mask(). */
  const uint8_t* polarized;     /* [n] or NULL */
  /* SUBBAND, optional banded form (faster): spline coefficients bcoef = cmat . stokes precomputed once
   * per catalogue by rsk_spline_coefficients(); on interval b only degree+1 of them contribute:
   * value = sum_p u^p * sum_m pband[b][p][m] * bcoef[s][first[b] + m][comp].  When bcoef is non-NULL it
   * is used instead of stokes/ppoly (breaks, nb, degree are shared by both forms). */
  const double* bcoef;          /* [4][nc][n] or NULL */
  int64_t nc;                   /* number of spline coefficients per spectrum */
  const int32_t* first;         /* [nb] */
  const double* pband;          /* [nb][degree + 1][degree + 1] */
  /* SUBBAND, dense form on a subset of the stored channels (NaN handling, skymodel.py:2245-2312: a group
   * of components whose non-NaN channels are chan_map[0..nf_use)): ppoly is [nb][degree+1][nf_use] and its
   * column j multiplies stored channel chan_map[j].  NULL = all nf_in channels. Not with bcoef. */
  const int32_t* chan_map;      /* [nf_use] or NULL */
  int64_t nf_use;
  /* SPECTRAL_INDEX: the power law is evaluated as exp(alpha * (ln(f / pivot) - ln(f_ref / pivot))); any
   * pivot inside the band works (it only keeps the logarithms small).  0 = the middle requested channel.
   * Callers that evaluate the frequency axis in slabs pass one pivot so every slab rounds identically. */
  double pivot_hz;
  /* Optional component filter (NaN handling, skymodel.py:2245-2312, without a host round trip): when group_of is
   * non-NULL a column is evaluated only if group_of[component] == group; every other column of the output is left
   * untouched.  The caller first evaluates all columns with the catalogue's operator and then launches once per
   * NaN-pattern group with that group's operator (chan_map form) and this filter -- nothing depends on which
   * components are above the horizon, so the whole step can be queued (or captured in a CUDA graph) as is. */
  const int32_t* group_of;      /* [n] or NULL */
  int32_t group;
  int32_t reserved;
  /* FULL, optional: a component-major copy of the stored Stokes, [n][nf_in][4] (the four Stokes parameters of one
   * (component, channel) in one 32-byte sector), made once per catalogue by rsk_stokes_component_major().  The
   * coherency kernels read the rows of the above-horizon components from it: with the (4, nf, n) layout a horizon
   * mask that is scattered over the catalogue order drags in nearly every sector of the below-horizon components
   * as well (~1.9x the bytes at 50 % density); this layout reads exactly 32 bytes per evaluation for any mask. */
  const double* stokes_cm;      /* [n][nf_in][4] or NULL */
} RskSpectrum;

/* Antenna response for rsk_apparent_coherency.  Jones matrices are [feed][basis] with
 * basis 0 = the zenith-angle direction and basis 1 = the azimuth direction, i.e. the (theta_hat, phi_hat)
 * basis the output of coherency_calc is expressed in. */
#define RSK_BEAM_JONES 0        /* caller-evaluated matrices (any beam model), one per column and frequency */
#define RSK_BEAM_SHORT_DIPOLE 1 /* analytic short dipole, feeds (north, east), as the reference's golden test
                                 * arranges pyuvdata's ShortDipoleBeam (tests/test_skymodel.py:1446-1462):
                                 * J = [[cos za sin az, cos az], [cos za cos az, -sin az]], all frequencies */
#define RSK_BEAM_UNIFORM 2      /* J = identity: the output equals rsk_coherency's */
typedef struct RskBeam {
  int32_t type;
  int32_t reserved;
  const double* jones1; /* RSK_BEAM_JONES: complex [2][2][nf][ld_jones], column k = output column k */
  const double* jones2; /* second antenna of the baseline, or NULL = jones1 */
  int64_t ld_jones;
} RskBeam;

/* ---- host-side helpers (no GPU work) -------------------------------------------------------- */

int rsk_version(void);
const char* rsk_strerror(int code);

/* Astrometry context for a terrestrial observer = what astropy's ErfaAstrom.apco(AltAz(obstime,
 * location)) feeds erfa.apco at the reference call site skymodel.py:2451-2454.  Medium-accuracy
 * inputs (truncated CIP series, low-precision solar ephemeris); callers with astropy
 * should pass the real eraASTROM instead.  All pointers HOST.
 * jd_utc1 + jd_utc2 = UTC Julian date; dut1 = UT1-UTC (s); elong/phi = WGS84 geodetic east longitude
 * and latitude (rad); height (m); xp, yp = polar motion (rad). */
int rsk_astrom_context(double jd_utc1, double jd_utc2, double dut1, double elong, double phi,
                       double height, double xp, double yp, double* astrom /* HOST [31] */);

/* The same for T times at once (an LST sweep's contexts in one call; Earth-orientation parameters shared):
 * jd_utc1/jd_utc2 HOST [T] -> astroms HOST [T][31]. */
int rsk_astrom_contexts(const double* jd_utc1, const double* jd_utc2, int64_t T, double dut1, double elong,
                        double phi, double height, double xp, double yp, double* astroms);

/* Geodetic WGS84 -> geocentric metres (astropy EarthLocation.geocentric; tutorial.rst:1231-1233). */
int rsk_geodetic_to_geocentric(double elong, double phi, double height, double* xyz /* HOST [3] */);

/* 3x3 row-major matrix taking catalogue-frame unit vectors to ICRS (the first leg of astropy's
 * frame -> ICRS -> AltAz graph for non-ICRS catalogues, e.g. GLEAM's FK5, skymodel.py:4182). HOST. */
int rsk_frame_to_icrs(int frame, double equinox_jyear, double* m /* HOST [9] */);

/* SkyModel._calc_average_rotation_matrix (skymodel.py:2499-2555): the orthogonal polar factor of the
 * AltAz images of the three ICRS axes, row-major, ICRS -> AltAz.  HOST pointers; runs the same
 * per-source chain as the kernels on the CPU for the three axes. */
int rsk_ravg(const double* astrom /* HOST [31] */, double* ravg /* HOST [9] */);

/* Pre-digest an eraASTROM into the RSK_CTX_LEN doubles the sweep kernel reads from device memory
 * (the single-time entry points do this internally and pass it by value).  HOST pointers. */
int rsk_pack_context(const double* astrom /* HOST [31] */, const double* frame2icrs /* HOST [9] or NULL */,
                     double* ctx /* HOST [RSK_CTX_LEN] */);

/* The same for T contexts at once: astroms HOST [T][31] -> ctx HOST [T][RSK_CTX_LEN]. */
int rsk_pack_contexts(const double* astroms, int64_t T, const double* frame2icrs /* HOST [9] or NULL */,
                      double* ctx);

size_t rsk_workspace_bytes(int op, int64_t n, int64_t nf);

/* ---- K1: positions --------------------------------------------------------------------------- */

/* SkyModel.update_positions (skymodel.py:2413-2479) for one time: catalogue lon/lat (or HEALPix pixel
 * centres) -> alt_az [2][n] = (alt; az), pos_lmn [3][n] (optional) and above_horizon [n] = alt > 0.
 * The per-source arithmetic is astropy's atciqz (light deflection eraLd, aberration eraAb, BPN eraRxp)
 * followed by erfa.atioq (Earth rotation, polar motion, diurnal aberration, latitude rotation,
 * A/B refraction); alt = pi/2 - zd as astropy does. */
int rsk_positions(const RskCoords* coords, int64_t n,
                  const double* frame2icrs /* HOST [9] row-major, NULL = identity */,
                  const double* astrom /* HOST [31] */,
                  const double* ravg /* HOST [9] or NULL: with coords->rot, also emit the rotation coefficients */,
                  double* alt_az /* [2][n] */, double* pos_lmn /* [3][n] or NULL */,
                  uint8_t* above /* [n] or NULL */, void* stream);

/* rsk_positions with everything that depends on the time read from DEVICE memory: ``step`` holds the RSK_STEP_LEN
 * doubles of rsk_pack_step() (packed context, R_avg, flag).  No launch argument changes from one time step to the
 * next, so positions -> compaction -> coherency can be captured once in a CUDA graph (cudaStreamBeginCapture on the
 * caller's stream; all entry points only enqueue work) and replayed after copying a new ``step`` record into place.
 * With coords->rot non-NULL and the flag set, the rotation coefficients are written as by rsk_positions. */
int rsk_positions_step(const RskCoords* coords, int64_t n, const double* step /* DEVICE [RSK_STEP_LEN] */,
                       double* alt_az, double* pos_lmn, uint8_t* above, void* stream);

/* HOST: the record above from an eraASTROM (rsk_pack_context + rsk_ravg). */
int rsk_pack_step(const double* astrom /* HOST [31] */, const double* frame2icrs /* HOST [9] or NULL */,
                  int with_ravg, double* step /* HOST [RSK_STEP_LEN] */);

/* LST sweep: the same for T contexts in one launch (BASELINE config 5; the loop of
 * tests/test_skymodel.py:2661-2740).  ctx = T packed contexts from rsk_pack_context(), uploaded by
 * the caller.  Outputs are time-major: alt_az [T][2][n], pos_lmn [T][3][n] or NULL, above [T][n] or NULL. */
int rsk_positions_sweep(const RskCoords* coords, int64_t n, const double* ctx /* [T][RSK_CTX_LEN] */,
                        int64_t T, double* alt_az, double* pos_lmn, uint8_t* above, void* stream);

/* SkyModel.get_lon_lat, healpix branch (skymodel.py:1656-1669): pixel-centre lon/lat arrays. */
int rsk_healpix_lonlat(int64_t nside, int nested, const int64_t* hpx_inds /* [n] or NULL */,
                       int64_t hpx_first, int64_t n, double* lon, double* lat, void* stream);

/* The inverse map: pix[i] = the HEALPix pixel that contains (lon[i], lat[i]) -- what
 * astropy_healpix.HEALPix(nside, order, frame).skycoord_to_healpix does for SkyModel.assign_to_healpix
 * (skymodel.py:1916).  rsk_lonlat_to_healpix(rsk_healpix_lonlat(p)) == p for every pixel. */
int rsk_lonlat_to_healpix(int64_t nside, int nested, const double* lon, const double* lat, int64_t n,
                          int64_t* pix /* [n] */, void* stream);

/* In-place x[i] *= factor (divide == 0) or x[i] /= factor (divide != 0), correctly rounded: the pixel-area
 * scaling of healpix_to_point (stokes * area, skymodel.py:1722) and _point_to_healpix / assign_to_healpix
 * (stokes / area, :1787, :1990-2003) with area = 4 pi / (12 nside^2) sr. */
int rsk_scale(double* x, int64_t count, double factor, int divide, void* stream);

/* Flux of the components that map to one pixel (skymodel.py:1967-1979): out[r][j] = sum over
 * k in [start[j], start[j+1]) of in[r][order[k]], accumulated in that order (quadrature != 0: the square root
 * of the sum of squares, used for stokes_error); r runs over the nrows = 4 * Nfreqs rows of length n_in. */
int rsk_segment_sum(const double* in /* [nrows][n_in] */, int64_t nrows, int64_t n_in,
                    const int64_t* order /* [n_in] */, const int64_t* start /* [nseg + 1] */, int64_t nseg,
                    int quadrature, double* out /* [nrows][nseg] */, void* stream);

/* ---- component down-selection: SkyModel.select / cut_nonrising / calculate_rise_set_lsts ---------- */

/* The per-component predicates of SkyModel.select (skymodel.py:3127-3283 with _select_lat :3012-3026, _select_lon
 * :3028-3045, _select_brightness :3047-3124) and cut_nonrising (:3331-3393), evaluated together.  A component is
 * kept when it passes every enabled test; the tests are the reference's comparisons (>=, <= bounds inclusive; a NaN
 * Stokes I among the tested channels fails a brightness cut, as np.min / np.max propagate it). */
typedef struct RskSelect {
  int32_t use_lat;        /* lat_lo <= lat <= lat_hi */
  int32_t use_lon;        /* lon_lo <= lon <= lon_hi, or lon >= lon_lo || lon <= lon_hi when lon_hi < lon_lo (wrap) */
  int32_t use_min;        /* min over the tested channels of Stokes I >= min_brightness */
  int32_t use_max;        /* max over the tested channels of Stokes I <= max_brightness */
  int32_t non_nan;        /* 0 = off, 1 = "any": no NaN in any Stokes at any channel, 2 = "all": not NaN at all channels */
  int32_t non_negative;   /* no Stokes I < 0 at any channel */
  int32_t cut_nonrising;  /* drop tan(tel_lat) tan(lat) < -1 (never above the horizon) */
  int32_t reserved;
  double lat_lo, lat_hi, lon_lo, lon_hi; /* radians, in the catalogue frame */
  double min_brightness, max_brightness; /* in the unit of stokes */
  double tan_tel_lat;
  const uint8_t* freq_use; /* DEVICE [nf]: channels the brightness cuts look at, NULL = all (brightness_freq_range) */
} RskSelect;

/* keep[j] = 1 if component comp_inds[j] (j itself when comp_inds is NULL) passes, else 0.  Feed keep to
 * rsk_compact_index for the order-preserving index list.  For HEALPix coordinates the pixel centres are used, as
 * SkyModel.get_lon_lat gives them. */
int rsk_select_mask(const RskCoords* coords, const double* stokes /* [4][nf][n] */, int64_t nf, int64_t n,
                    const int64_t* comp_inds /* DEVICE [m] in [0, n), or NULL */, int64_t m, const RskSelect* sel,
                    uint8_t* keep /* [m] */, void* stream);

/* out[r][j] = in[r][idx[j]] for nrows rows: the fancy index along the component axis that
 * SkyModel._select_along_param_axis applies to every Ncomponents-shaped parameter (stokes, coordinates, spectral
 * parameters, masks).  elem_bytes = 8 (doubles, int64 indices) or 1 (masks). */
int rsk_gather_rows(const void* in /* [nrows][n_in] */, int64_t nrows, int64_t n_in, const int64_t* idx /* [n_out] */,
                    int64_t n_out, int elem_bytes, void* out /* [nrows][n_out] */, void* stream);

/* out[r] = a[r] followed by b[r] for nrows rows: np.concatenate along the component axis, as SkyModel.concat applies
 * it to every Ncomponents-shaped parameter (skymodel.py:2905-2999).  a or b may be NULL: that side is then filled
 * with NaN (elem_bytes = 8; the reference's fill for a numeric parameter only one object has, :2960-2999) or 0
 * (elem_bytes = 1). */
int rsk_concat_rows(const void* a /* [nrows][n_a] or NULL */, int64_t n_a, const void* b /* [nrows][n_b] or NULL */,
                    int64_t n_b, int64_t nrows, int elem_bytes, void* out /* [nrows][n_a + n_b] */, void* stream);

/* SkyModel.transform_to (skymodel.py:1419-1446, astropy SkyCoord.transform_to) between frames related by a fixed
 * rotation (icrs, fk5(equinox), galactic): unit vector of (lon, lat), times rot (HOST [9], row-major, = the transpose
 * of the new frame's rsk_frame_to_icrs matrix times the old frame's), back to angles; lon in [0, 2 pi].  In place
 * (lon_out == lon, lat_out == lat) is allowed. */
int rsk_rotate_lonlat(const double* lon, const double* lat, int64_t n, const double* rot /* HOST [9] */,
                      double* lon_out, double* lat_out, void* stream);

/* SkyModel.healpix_interp_transform (skymodel.py:1448-1538): the map at the pixel centres of a NEW frame, bilinearly
 * interpolated (astropy_healpix.HEALPix.interpolate_bilinear_skycoord, :1507-1509) from a map given on pixels of the
 * old frame.  rot_new_to_old: HOST [9] row-major rotation of unit vectors from the new frame to the old one.  The
 * input rows hold n_in pixels: pixel p of the old map sits at column lookup[p] (DEVICE [12 nside^2], -1 = absent), or
 * at p - hpx_first when lookup is NULL.  out [nrows][12 nside^2] in the map's own ordering; valid[j] = 1 when all
 * four neighbours of new pixel j are present (the reference's masked-array result), else 0 and the row values are 0. */
int rsk_healpix_interp(int64_t nside, int nested, const double* rot_new_to_old /* HOST [9] */,
                       const int64_t* lookup /* DEVICE [12 nside^2] or NULL */, int64_t hpx_first, int64_t n_in,
                       const double* in /* [nrows][n_in] */, int64_t nrows, double* out /* [nrows][12 nside^2] */,
                       uint8_t* valid /* [12 nside^2] */, void* stream);

/* SkyModel.calculate_rise_set_lsts (skymodel.py:3284-3329): rise = lon - acos(-tan(tel_lat) tan(lat)) - buffer,
 * set = lon + acos(...) + buffer, each wrapped once into [0, 2 pi]; NaN for components that never rise or never set. */
int rsk_rise_set_lsts(const RskCoords* coords, int64_t n, double tel_lat, double horizon_buffer,
                      double* rise_lst /* [n] */, double* set_lst /* [n] */, void* stream);

/* ---- K1b: horizon compaction ------------------------------------------------------------------ */

/* Order-preserving compaction of the horizon mask = the boolean fancy index of skymodel.py:2685 and
 * the np.isin / list-comprehension bookkeeping of :2693-2699.  idx[k] = index of the k-th component
 * with above != 0; *n_above (device int64) = their number.  work: rsk_workspace_bytes(RSK_OP_COMPACT, n, 0). */
int rsk_compact_index(const uint8_t* above, int64_t n, int64_t* idx /* [n] */,
                      int64_t* n_above /* device [1] */, void* work, size_t work_bytes, void* stream);

/* Per-component bit mask of stored channels at which all four Stokes parameters are non-NaN: the freq_inds_use of
 * skymodel.py:2246-2248 for every component at once.  Channel j is bit (j % 64) of word j / 64;
 * mask is [ceil(nf_in / 64)][n] (one word per component for up to 64 subbands). */
int rsk_valid_channels(const double* stokes /* [4][nf_in][n] */, int64_t nf_in, int64_t n,
                       uint64_t* mask /* [ceil(nf_in / 64)][n] */, void* stream);

/* *out (DEVICE int64) = number of NaNs in a double array (RSK_COUNT_NAN_F64: the np.any(np.isnan(...)) tests of
 * skymodel.py:2157 and :2196) or of non-zero bytes in a byte array (RSK_COUNT_NONZERO_U8: the size of
 * SkyModel._polarized, the above-horizon count of a mask). */
#define RSK_COUNT_NAN_F64 0
#define RSK_COUNT_NONZERO_U8 1
#define RSK_COUNT_NEGATIVE_F64 2 /* the negative-Stokes-I acceptability test of SkyModel.check (skymodel.py:1258-1264) */
int rsk_count(const void* x, int64_t count, int kind, int64_t* out /* device [1] */, void* stream);

/* SkyModel._polarized (skymodel.py:1037-1042): pol[c] = any over stored channels of (Q+U+V != 0). */
int rsk_polarized_mask(const double* stokes /* [4][nf_in][n] */, int64_t nf_in, int64_t n,
                       uint8_t* pol /* [n] */, void* stream);

/* ---- K2: fused spectrum + rotation + coherency ----------------------------------------------- */

/* SkyModel.coherency_calc (skymodel.py:2648-2712) fused with at_frequencies (:2084-2389),
 * stokes_to_coherency (utils.py:17-61), _calc_rotation_matrix (:2557-2601, with sct.vecs2rot
 * spherical_coords_transforms.py:259-310) and _calc_coherency_rotation (:2603-2646, with
 * sct.spherical_basis_vector_rotation_matrix :140-186).  One thread per output column.
 * Column k (0 <= k < ncols) is component idx[k] (idx NULL = identity); its 2x2xnf block is written to
 * out[(a*2+b)*nf*ld_out + f*ld_out + col0 + k] as (re, im).  ld_out/col0 let several shards write
 * disjoint column ranges of one gathered array, and let the caller pad rows: HBM write throughput on
 * B200 is best when ld_out is a multiple of 16 (rows start on 256-byte boundaries).
 * ncols_dev (optional): the live column count in device memory, e.g. the n_above written by
 * rsk_compact_index; columns >= min(ncols, *ncols_dev) are not written.  With it the whole time step
 * (positions -> compaction -> coherency) can be queued without a device->host read of the count;
 * ncols is then only the upper bound that sizes the grid.  alt_az is the [2][n] output of rsk_positions. */
int rsk_coherency(const RskSpectrum* spec, const RskCoords* coords, const double* alt_az /* [2][n] */,
                  const double* ravg /* HOST [9] */, const int64_t* idx /* [ncols] or NULL */,
                  const int64_t* cols /* [ncols] or NULL: thread k computes column cols[k] (a subset) */,
                  int64_t ncols, const int64_t* ncols_dev /* device [1] or NULL */,
                  const double* freqs /* [nf] Hz */, int64_t nf,
                  double* out /* complex [2][2][nf][ld_out] */, int64_t ld_out, int64_t col0,
                  void* stream);

/* The same computation, written in a Hermitian-packed form for transfers over PCIe: the 2x2 coherency is
 * Hermitian with a real diagonal, so 4 real numbers per evaluation carry it.  out is real [4][nf][ld_out] with
 * planes (C00, C11, Re C01 = Re C10, Im C10 = -Im C01): 32 instead of 64 bytes per evaluation. */
int rsk_coherency_packed(const RskSpectrum* spec, const RskCoords* coords, const double* alt_az,
                         const double* ravg, const int64_t* idx, const int64_t* cols, int64_t ncols,
                         const int64_t* ncols_dev, const double* freqs, int64_t nf,
                         double* out /* real [4][nf][ld_out] */, int64_t ld_out, int64_t col0, void* stream);

/* The consumer of coherency_calc fused into the same kernel: the apparent (instrumental) coherency
 * A = J1 C J2^H of every (above-horizon component, frequency), i.e. the contraction
 * einsum("ab...,bc...,dc...->ad...", J, C, conj(J)) that pyuvsim and the reference's golden test
 * (tests/test_skymodel.py:1464-1466) apply to the output of coherency_calc.  The Az/El coherency is never
 * written: per evaluation the kernel reads the Jones matrices (RSK_BEAM_JONES: 64 or 128 bytes) and writes
 * 64 bytes.  Arguments as rsk_coherency; alt_az is required for RSK_BEAM_SHORT_DIPOLE. */
int rsk_apparent_coherency(const RskSpectrum* spec, const RskCoords* coords, const double* alt_az,
                            const double* ravg, const int64_t* idx, const int64_t* cols, int64_t ncols,
                            const int64_t* ncols_dev, const double* freqs, int64_t nf, const RskBeam* beam,
                            double* out /* complex [2][2][nf][ld_out] */, int64_t ld_out, int64_t col0,
                            void* stream);

/* The source sum of the apparent coherency -- what a visibility simulator finally needs from this path:
 *   vis[b][a][d][f] = sum over the ncols above-horizon components s of
 *                     (J1 C J2^H)[a][d][f][s] * exp(2 pi i (f / c) (uvw_b . lmn_s))
 * (pyuvsim's convention: fringe = exp(2j pi uvw_wavelengths . pos_lmn)).  Nothing per-component is written:
 * neither the (4, Nf, N) Stokes cube, nor the Az/El coherency, nor the apparent one.  uvw_m == NULL (or nbl == 0)
 * gives the zero-baseline sum, vis[0].  Spectral modes FLAT, SPECTRAL_INDEX and FULL (evaluate subband models with
 * rsk_at_frequencies first).  pos_lmn is the [3][n] output of rsk_positions.  work: rsk_visibility_workspace_bytes().
 * Deterministic: per-thread running sums in a fixed column order, fixed-shape block reduction, column groups
 * added in order. */
int rsk_visibility(const RskSpectrum* spec, const RskCoords* coords, const double* alt_az, const double* ravg,
                   const int64_t* idx, int64_t ncols, const int64_t* ncols_dev, const double* freqs, int64_t nf,
                   const RskBeam* beam, const double* pos_lmn /* [3][n] or NULL */,
                   const double* uvw_m /* DEVICE [nbl][3] metres, or NULL */, int64_t nbl,
                   double* vis /* complex [max(nbl,1)][2][2][nf] */, void* work, size_t work_bytes, void* stream);
size_t rsk_visibility_workspace_bytes(int64_t ncols, int64_t nf, int64_t nbl);

/* Pitched device-to-device copy of ``height`` rows of ``width_bytes`` (cudaMemcpy2DAsync): how a rank's column block of the
 * gathered (2, 2, Nf, sum N_above) array travels to the same place of a peer GPU's array over NVLink (the optional
 * assembly of north_star; dst may be peer-mapped memory).  Copy engines, asynchronous on ``stream``. */
int rsk_copy_2d(void* dst, size_t dst_pitch_bytes, const void* src, size_t src_pitch_bytes, size_t width_bytes,
                size_t height, void* stream);

/* HOST: expand a packed slab (HOST pointer, e.g. pinned staging memory that a cudaMemcpyAsync has filled)
 * into rows [f0_out, f0_out + nf) of the reference-layout array out = complex [2][2][nf_out][ld_out] in host
 * memory, with nthreads threads.  Pure data movement (duplicate / negate), no arithmetic of the path. */
int rsk_unpack_coherency_host(const double* packed /* HOST [4][nf][ld_packed] */, int64_t nf, int64_t ncols,
                              int64_t ld_packed, double* out /* HOST complex [2][2][nf_out][ld_out] */,
                              int64_t nf_out, int64_t f0_out, int64_t ld_out, int nthreads);

/* Spline coefficients of every stored spectrum: bcoef[s][i][c] = sum_j cmat[i][j] * stokes[s][j][c].
 * cmat is the (nc x nf_in) matrix of scipy's make_interp_spline for the catalogue's freq_array and
 * freq_interp_kind (host-built, uploaded by the caller).  Time independent: run once per catalogue. */
int rsk_spline_coefficients(const double* stokes /* [4][nf_in][n] */, int64_t nf_in, int64_t n,
                            const double* cmat /* [nc][nf_in] */, int64_t nc,
                            double* bcoef /* [4][nc][n] */, void* stream);

/* stokes [4][nf][n] -> out [n][nf][4] (see RskSpectrum.stokes_cm).  Time independent: run once per catalogue. */
int rsk_stokes_component_major(const double* stokes, int64_t nf, int64_t n, double* out, void* stream);

/* ---- K3 / K4: unfused API-parity kernels ------------------------------------------------------- */

/* SkyModel.at_frequencies (skymodel.py:2084-2389): evaluate the spectral model on ``freqs`` ->
 * stokes_out [4][nf][n], for all n components (comps NULL) or only for components comps[0..ncomps)
 * (the per-NaN-pattern groups of :2245-2312; other columns of stokes_out are left untouched). */
int rsk_at_frequencies(const RskSpectrum* spec, const double* freqs /* [nf] */, int64_t nf,
                       const int64_t* comps /* [ncomps] or NULL */, int64_t ncomps,
                       double* stokes_out, void* stream);

/* utils.stokes_to_coherency / SkyModel.calc_frame_coherency (utils.py:17-61, skymodel.py:2481-2497):
 * stokes [4][nf][n] -> out complex [2][2][nf][n]. */
int rsk_stokes_to_coherency(const double* stokes, int64_t nf, int64_t n, double* out, void* stream);

/* utils.coherency_to_stokes (utils.py:64-103), the inverse: coherency [2][2][count] complex128 (interleaved re, im)
 * -> stokes [4][count] = (Re(C00 + C11), Re(C00 - C11), Re(C01 + C10), -Im(C01 - C10)). */
int rsk_coherency_to_stokes(const double* coherency, int64_t count, double* stokes, void* stream);

/* SkyModel.jansky_to_kelvin / kelvin_to_jansky (skymodel.py:1593-1635, :1541-1591) with
 * utils.jy_to_ksr (utils.py:105-126): in-place stokes[s][f][c] *= k(nu) where k = 1e-26 c^2 / (2 k_B nu^2)
 * (or 1/k when to_jansky != 0); nu = freq_per_comp[c] if that pointer is non-NULL (spectral_index /
 * flat-with-reference-frequency branch), else freq_per_chan[f]. */
int rsk_convert_units(double* stokes /* [4][nf][n] */, int64_t nf, int64_t n,
                      const double* freq_per_chan /* [nf] or NULL */,
                      const double* freq_per_comp /* [n] or NULL */, int to_jansky, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RSK_H */
