"""Oracle (test infrastructure): numpy restatement of the ERFA/astropy ICRS -> AltAz chain.

The reference delegates the whole RA/Dec -> Alt/Az step to astropy
(``src/pyradiosky/skymodel.py:2451-2454`` and ``:2518-2527``), i.e. to
``astropy.coordinates.builtin_frames.icrs_observed_transforms.icrs_to_observed`` ->
``erfa_astrom.ErfaAstrom.apco`` + ``builtin_frames.utils.atciqz`` + ``erfa.atioq`` (pyerfa/liberfa).
Those packages are third-party dependencies that are absent from /root/reference and from this
image (astropy>=6.0 per the reference's pyproject.toml:21; pyerfa/liberfa un-pinned), so the functions
below restate the published ERFA (BSD) algorithms by name: ``s2c c2s anp anpm ld ab rxp atioq apcs
apco pvtob gd2gc pom00 c2ixys era00 sp00``.  They are vectorised over sources.

PARITY UNPINNED for absolute Alt/Az against astropy (see oracle/__init__.py); pinned indirectly by
the golden vectors in tests/golden (reference test_skymodel.py:1468-1506, tutorial.rst:1129-1132).

The per-time context ("astrom") is a float64[31] vector laid out exactly like ERFA's ``eraASTROM``
(``erfa.dt_eraASTROM``): pmt, eb[3], eh[3], em, v[3], bm1, bpn[3][3] row-major, along, phi, xpl, ypl,
sphi, cphi, diurab, eral, refa, refb  (1+3+3+1+3+1+9+10 = 31 doubles, 248 bytes).
"""

import numpy as np

# ERFA constants (erfam.h)
D2PI = 6.283185307179586476925287
DAS2R = 4.848136811095359935899141e-6
DJ00 = 2451545.0
DJY = 365.25
DJC = 36525.0
DAYSEC = 86400.0
SRS = 1.97412574336e-8
DAU = 149597870.7e3
AULT = DAU / 299792458.0  # erfam.h: ERFA_AULT = ERFA_DAU / ERFA_CMPS (pinned by t_erfa_c's apcs check values)
WGS84_A = 6378137.0
WGS84_F = 1.0 / 298.257223563
CELMIN = 1e-6
SELMIN = 0.05

# astrom[31] slots
ASTROM_LEN = 31
I_PMT, I_EB, I_EH, I_EM, I_V, I_BM1, I_BPN = 0, 1, 4, 7, 8, 11, 12
I_ALONG, I_PHI, I_XPL, I_YPL, I_SPHI, I_CPHI, I_DIURAB, I_ERAL, I_REFA, I_REFB = range(21, 31)

# TAI-UTC leap-second table (start MJD, TAI-UTC seconds), 1972 onward (IERS Bulletin C).
_LEAP = [
    (41317, 10), (41499, 11), (41683, 12), (42048, 13), (42413, 14), (42778, 15), (43144, 16),
    (43509, 17), (43874, 18), (44239, 19), (44786, 20), (45151, 21), (45516, 22), (46247, 23),
    (47161, 24), (47892, 25), (48257, 26), (48804, 27), (49169, 28), (49534, 29), (50083, 30),
    (50630, 31), (51179, 32), (53736, 33), (54832, 34), (56109, 35), (57204, 36), (57754, 37),
]


def tai_minus_utc(jd_utc):
    """TAI-UTC in seconds (ERFA ``dat`` restated as a table lookup, valid from 1972)."""
    mjd = jd_utc - 2400000.5
    out = 10
    for start, val in _LEAP:
        if mjd >= start:
            out = val
    return float(out)


def anp(a):
    """ERFA ``eraAnp``: normalise angle into [0, 2pi)."""
    w = np.fmod(a, D2PI)
    return np.where(w < 0, w + D2PI, w)


def anpm(a):
    """ERFA ``eraAnpm``: normalise angle into [-pi, pi)."""
    w = np.fmod(a, D2PI)
    return np.where(np.abs(w) >= np.pi, w - np.copysign(D2PI, a), w)


def s2c(theta, phi):
    """ERFA ``eraS2c``: spherical (lon, lat) to unit vector, shape (3, ...)."""
    cp = np.cos(phi)
    return np.stack((np.cos(theta) * cp, np.sin(theta) * cp, np.sin(phi)))


def c2s(p):
    """ERFA ``eraC2s``: vector to spherical (lon, lat)."""
    x, y, z = p[0], p[1], p[2]
    d2 = x * x + y * y
    theta = np.where(d2 == 0.0, 0.0, np.arctan2(y, x))
    phi = np.where(z == 0.0, 0.0, np.arctan2(z, np.sqrt(d2)))
    return theta, phi


def _rx(phi):
    s, c = np.sin(phi), np.cos(phi)
    return np.array([[1.0, 0, 0], [0, c, s], [0, -s, c]])


def _ry(theta):
    s, c = np.sin(theta), np.cos(theta)
    return np.array([[c, 0, -s], [0, 1.0, 0], [s, 0, c]])


def _rz(psi):
    s, c = np.sin(psi), np.cos(psi)
    return np.array([[c, s, 0], [-s, c, 0], [0, 0, 1.0]])


def ld(bm, p, q, e, em, dlim):
    """ERFA ``eraLd``: light deflection by a solar-system body. p, q: (3, N); e: (3,)."""
    e = np.asarray(e, dtype=float).reshape(3, *([1] * (p.ndim - 1)))
    qpe = q + e
    qdqpe = np.sum(q * qpe, axis=0)
    w = bm * SRS / em / np.maximum(qdqpe, dlim)
    eq = np.cross(np.broadcast_to(e, q.shape), q, axis=0)
    peq = np.cross(p, eq, axis=0)
    return p + w * peq


def ab(pnat, v, s, bm1):
    """ERFA ``eraAb``: stellar aberration. pnat: (3, N); v: (3,) in units of c."""
    v = np.asarray(v, dtype=float).reshape(3, *([1] * (pnat.ndim - 1)))
    pdv = np.sum(pnat * v, axis=0)
    w1 = 1.0 + pdv / (1.0 + bm1)
    w2 = SRS / s
    p = pnat * bm1 + w1 * v + w2 * (v - pdv * pnat)
    r = np.sqrt(np.sum(p * p, axis=0))
    return p / r


def rxp(r, p):
    """ERFA ``eraRxp``: r (3,3) times p (3, N)."""
    r = np.asarray(r, dtype=float).reshape(3, 3)
    return np.stack([r[i, 0] * p[0] + r[i, 1] * p[1] + r[i, 2] * p[2] for i in range(3)])


def atciqz(rc, dc, astrom):
    """astropy ``builtin_frames.utils.atciqz`` (distance-less branch): ICRS -> CIRS (ri, di)."""
    pos = s2c(rc, dc)
    pnat = ld(1.0, pos, pos, astrom[I_EH:I_EH + 3], astrom[I_EM], 1e-6)
    ppr = ab(pnat, astrom[I_V:I_V + 3], astrom[I_EM], astrom[I_BM1])
    pi = rxp(astrom[I_BPN:I_BPN + 9], ppr)
    ri, di = c2s(pi)
    return anp(ri), di


def atioq(ri, di, astrom):
    """ERFA ``eraAtioq``: CIRS -> observed. Returns (aob, zob) = (azimuth, zenith distance)."""
    v = s2c(ri - astrom[I_ERAL], di)
    x, y, z = v[0], v[1], v[2]
    sx, cx = np.sin(astrom[I_XPL]), np.cos(astrom[I_XPL])
    sy, cy = np.sin(astrom[I_YPL]), np.cos(astrom[I_YPL])
    xhd = cx * x + sx * z
    yhd = sx * sy * x + cy * y - cx * sy * z
    zhd = -sx * cy * x + sy * y + cx * cy * z
    diurab = astrom[I_DIURAB]
    f = 1.0 - diurab * yhd
    xhdt = f * xhd
    yhdt = f * (yhd + diurab)
    zhdt = f * zhd
    sphi, cphi = astrom[I_SPHI], astrom[I_CPHI]
    xaet = sphi * xhdt - cphi * zhdt
    yaet = yhdt
    zaet = cphi * xhdt + sphi * zhdt
    azobs = np.where((xaet != 0.0) | (yaet != 0.0), np.arctan2(yaet, -xaet), 0.0)
    r = np.maximum(np.sqrt(xaet * xaet + yaet * yaet), CELMIN)
    zz = np.maximum(zaet, SELMIN)
    tz = r / zz
    w = astrom[I_REFB] * tz * tz
    dl = (astrom[I_REFA] + w) * tz / (1.0 + (astrom[I_REFA] + 3.0 * w) / (zz * zz))
    cosdel = 1.0 - dl * dl / 2.0
    f = cosdel - dl * zz / r
    xaeo = xaet * f
    yaeo = yaet * f
    zaeo = cosdel * zaet + dl * r
    zdobs = np.arctan2(np.sqrt(xaeo * xaeo + yaeo * yaeo), zaeo)
    return anp(azobs), zdobs


def icrs_to_altaz(lon, lat, astrom, frame_to_icrs=None):
    """astropy ``icrs_to_observed`` for an AltAz target: returns (alt, az) in radians.

    ``frame_to_icrs`` (3x3) is the matrix astropy applies first when the catalogue frame is not ICRS
    (e.g. FK5 for GLEAM, reference skymodel.py:4182); identity if None.
    """
    lon = np.asarray(lon, dtype=float)
    lat = np.asarray(lat, dtype=float)
    if frame_to_icrs is not None:
        p = rxp(frame_to_icrs, s2c(lon, lat))
        lon, lat = c2s(p)
    ri, di = atciqz(lon, lat, astrom)
    az, zd = atioq(ri, di, astrom)
    return np.pi / 2.0 - zd, az


def altaz_to_icrs(alt, az, astrom, niter=6):
    """Inverse of :func:`icrs_to_altaz` (refraction-free), used to build the reference's fixtures
    (test_skymodel.py:58-69 builds its zenith source by ``AltAz -> ICRS``).  astropy inverts with
    ``atoiq`` + ``aticq``; here the orthogonal steps are inverted exactly and deflection/aberration
    by fixed-point iteration (they are ~1e-4 rad perturbations, so 6 rounds reach 1e-16)."""
    alt = np.atleast_1d(np.asarray(alt, dtype=float))
    az = np.atleast_1d(np.asarray(az, dtype=float))
    ca = np.cos(alt)
    xaet = -np.cos(az) * ca
    yaet = np.sin(az) * ca
    zaet = np.sin(alt)
    sphi, cphi = astrom[I_SPHI], astrom[I_CPHI]
    xhd = sphi * xaet + cphi * zaet
    yhd = yaet
    zhd = -cphi * xaet + sphi * zaet
    sx, cx = np.sin(astrom[I_XPL]), np.cos(astrom[I_XPL])
    sy, cy = np.sin(astrom[I_YPL]), np.cos(astrom[I_YPL])
    # inverse (transpose) of the polar-motion rotation of atioq
    x = cx * xhd + sx * sy * yhd - sx * cy * zhd
    y = cy * yhd + sy * zhd
    z = sx * xhd - cx * sy * yhd + cx * cy * zhd
    mha, di = c2s(np.stack((x, y, z)))
    ri = mha + astrom[I_ERAL]
    pi = s2c(ri, di)
    bpn = astrom[I_BPN:I_BPN + 9].reshape(3, 3)
    ppr = rxp(bpn.T, pi)
    p = ppr.copy()
    for _ in range(niter):
        pnat = ld(1.0, p, p, astrom[I_EH:I_EH + 3], astrom[I_EM], 1e-6)
        fwd = ab(pnat, astrom[I_V:I_V + 3], astrom[I_EM], astrom[I_BM1])
        p = p - (fwd - ppr)
        p = p / np.sqrt(np.sum(p * p, axis=0))
    rc, dc = c2s(p)
    return anp(rc), dc


# --------------------------------------------------------------------------------------------
# host context (once per time step): eraApco and its inputs
# --------------------------------------------------------------------------------------------


def era00(dj1, dj2):
    """ERFA ``eraEra00``: Earth rotation angle from a two-part UT1 JD."""
    d1, d2 = (dj1, dj2) if dj1 < dj2 else (dj2, dj1)
    t = d1 + (d2 - DJ00)
    f = np.fmod(d1, 1.0) + np.fmod(d2, 1.0)
    return float(anp(D2PI * (f + 0.7790572732640 + 0.00273781191135448 * t)))


def sp00(date1, date2):
    """ERFA ``eraSp00``: TIO locator s'."""
    t = ((date1 - DJ00) + date2) / DJC
    return -47e-6 * t * DAS2R


def gd2gc_wgs84(elong, phi, height):
    """ERFA ``eraGd2gc(1, ...)`` / ``eraGd2gce``: WGS84 geodetic -> geocentric metres."""
    sp, cp = np.sin(phi), np.cos(phi)
    w = (1.0 - WGS84_F) ** 2
    d = cp * cp + w * sp * sp
    ac = WGS84_A / np.sqrt(d)
    a_s = w * ac
    r = (ac + height) * cp
    return np.array([r * np.cos(elong), r * np.sin(elong), (a_s + height) * sp])


def pom00(xp, yp, sp):
    """ERFA ``eraPom00``: polar-motion matrix Rx(-yp) Ry(-xp) Rz(sp)."""
    return _rx(-yp) @ _ry(-xp) @ _rz(sp)


def pvtob(elong, phi, hm, xp, yp, sp, theta):
    """ERFA ``eraPvtob``: observer position/velocity (m, m/s) in CIRS."""
    om = 1.00273781191135448 * D2PI / DAYSEC
    xyzm = gd2gc_wgs84(elong, phi, hm)
    x, y, z = pom00(xp, yp, sp).T @ xyzm
    s, c = np.sin(theta), np.cos(theta)
    pos = np.array([c * x - s * y, s * x + c * y, z])
    vel = np.array([om * (-s * x - c * y), om * (c * x - s * y), 0.0])
    return pos, vel


def c2ixys(x, y, s):
    """ERFA ``eraC2ixys``: celestial-to-intermediate matrix from CIP X, Y and CIO locator s."""
    r2 = x * x + y * y
    e = np.arctan2(y, x) if r2 > 0.0 else 0.0
    d = np.arctan(np.sqrt(r2 / (1.0 - r2)))
    return _rz(-(e + s)) @ _ry(d) @ _rz(e)


def apcs(date1, date2, pv, ebpv, ehp):
    """ERFA ``eraApcs``: fills pmt, eb, eh, em, v, bm1 (bpn = identity) of a fresh astrom."""
    astrom = np.zeros(ASTROM_LEN)
    audms = DAU / DAYSEC
    cr = AULT / DAYSEC
    astrom[I_PMT] = ((date1 - DJ00) + date2) / DJY
    dp = pv[0] / DAU
    dv = pv[1] / audms
    pb = ebpv[0] + dp
    vb = ebpv[1] + dv
    ph = ehp + dp
    astrom[I_EB:I_EB + 3] = pb
    em = np.sqrt(np.sum(ph * ph))
    astrom[I_EM] = em
    astrom[I_EH:I_EH + 3] = ph / em
    v = vb * cr
    astrom[I_V:I_V + 3] = v
    astrom[I_BM1] = np.sqrt(1.0 - np.sum(v * v))
    astrom[I_BPN:I_BPN + 9] = np.eye(3).ravel()
    return astrom


def apco(date1, date2, ebpv, ehp, x, y, s, theta, elong, phi, hm, xp, yp, sp, refa, refb):
    """ERFA ``eraApco``: the ICRS<->observed context for a terrestrial observer."""
    r = _rz(elong) @ _rx(-yp) @ _ry(-xp) @ _rz(theta + sp)
    a, b = r[0, 0], r[0, 1]
    eral = np.arctan2(b, a) if (a != 0.0 or b != 0.0) else 0.0
    c = r[0, 2]
    xpl = np.arctan2(c, np.sqrt(a * a + b * b))
    a2, b2 = r[1, 2], r[2, 2]
    ypl = -np.arctan2(a2, b2) if (a2 != 0.0 or b2 != 0.0) else 0.0
    rc2i = c2ixys(x, y, s)
    pos, vel = pvtob(elong, phi, hm, xp, yp, sp, theta)
    pv = np.stack((rc2i.T @ pos, rc2i.T @ vel))
    astrom = apcs(date1, date2, pv, ebpv, ehp)
    astrom[I_ERAL] = eral
    astrom[I_XPL] = xpl
    astrom[I_YPL] = ypl
    astrom[I_ALONG] = float(anpm(eral - theta))
    astrom[I_PHI] = phi
    astrom[I_SPHI] = np.sin(phi)
    astrom[I_CPHI] = np.cos(phi)
    astrom[I_REFA] = refa
    astrom[I_REFB] = refb
    astrom[I_DIURAB] = 0.0
    astrom[I_BPN:I_BPN + 9] = rc2i.ravel()
    return astrom


# --------------------------------------------------------------------------------------------
# medium-accuracy inputs for apco (astropy uses pnm06a/s06/epv00/IERS tables; none available here)
# --------------------------------------------------------------------------------------------


def cip_xys_medium(t):
    """CIP X, Y (7/6 largest terms of the IAU 2006/2000A series) and s ~ -XY/2. t: TT centuries."""
    om = (450160.398036 - 6962890.5431 * t) * DAS2R
    lp = (1287104.79305 + 129596581.0481 * t) * DAS2R
    f = (335779.526232 + 1739527262.8478 * t) * DAS2R
    d = (1072260.70369 + 1602961601.2090 * t) * DAS2R
    x = (2004.191898 * t - 0.4297829 * t * t - 6.844318 * np.sin(om)
         - 0.523908 * np.sin(2 * (f - d + om)) - 0.090552 * np.sin(2 * (f + om))
         + 0.082169 * np.sin(2 * om) + 0.058707 * np.sin(lp))
    y = (-0.006951 - 0.025896 * t - 22.4072747 * t * t + 9.205236 * np.cos(om)
         + 0.573033 * np.cos(2 * (f - d + om)) + 0.097847 * np.cos(2 * (f + om))
         - 0.089618 * np.cos(2 * om))
    x *= DAS2R
    y *= DAS2R
    return x, y, -x * y / 2.0


def _sun_geocentric_j2000(jd_tt):
    """Low-precision solar theory (Astronomical Almanac), rotated to the J2000 equinox."""
    n = jd_tt - DJ00
    deg = np.pi / 180.0
    L = (280.460 + 0.9856474 * n) * deg
    g = (357.528 + 0.9856003 * n) * deg
    lam = L + (1.915 * np.sin(g) + 0.020 * np.sin(2 * g)) * deg
    lam = lam - (1.396971 * (n / DJC)) * deg  # general precession in longitude back to J2000
    R = 1.00014 - 0.01671 * np.cos(g) - 0.00014 * np.cos(2 * g)
    eps = 23.4392911 * deg
    return np.array([R * np.cos(lam), R * np.cos(eps) * np.sin(lam), R * np.sin(eps) * np.sin(lam)])


def earth_pv_medium(jd_tt):
    """Earth position (au) / velocity (au/day); heliocentric used for barycentric as well."""
    h = 0.01
    pos = -_sun_geocentric_j2000(jd_tt)
    vel = -(_sun_geocentric_j2000(jd_tt + h) - _sun_geocentric_j2000(jd_tt - h)) / (2 * h)
    return np.stack((pos, vel)), pos.copy()


def astrom_context(jd_utc1, jd_utc2, elong, phi, height, dut1=0.0, xp=0.0, yp=0.0):
    """astropy ``ErfaAstrom.apco(AltAz(obstime, location))`` with medium-accuracy inputs.

    jd_utc1 + jd_utc2: UTC Julian date; elong/phi: WGS84 geodetic east longitude / latitude (rad);
    height (m); dut1 = UT1-UTC (s); xp, yp polar motion (rad).  Pressure 0 => refa = refb = 0.
    """
    dtai = tai_minus_utc(jd_utc1 + jd_utc2)
    tt1, tt2 = jd_utc1, jd_utc2 + (dtai + 32.184) / DAYSEC
    ut11, ut12 = jd_utc1, jd_utc2 + dut1 / DAYSEC
    t = ((tt1 - DJ00) + tt2) / DJC
    x, y, s = cip_xys_medium(t)
    theta = era00(ut11, ut12)
    sp = sp00(tt1, tt2)
    ebpv, ehp = earth_pv_medium(tt1 + tt2)
    return apco(tt1, tt2, ebpv, ehp, x, y, s, theta, elong, phi, height, xp, yp, sp, 0.0, 0.0)


def frame_to_icrs(frame):
    """3x3 matrix taking FK5(J2000) / Galactic unit vectors to ICRS (identity for 'icrs').

    astropy ``builtin_frames.icrs_fk5_transforms._icrs_to_fk5_matrix`` (frame bias eta0 = -19.9 mas,
    xi0 = 9.1 mas, da0 = -22.9 mas) and ``galactic_transforms.fk5_to_gal`` (NGP 192.8594812065348,
    27.12825118085622 deg; lon0 122.9319185680026 deg); the catalogue -> ICRS leg of the transform the
    reference runs for non-ICRS catalogues such as GLEAM (skymodel.py:4182).
    """
    frame = str(frame).lower()
    if frame == "icrs":
        return np.eye(3)
    deg = np.pi / 180.0
    b = _rx(19.9 / 3600e3 * deg) @ _ry(9.1 / 3600e3 * deg) @ _rz(-22.9 / 3600e3 * deg)
    if frame == "fk5":
        return b.T
    if frame == "galactic":
        g = _rz((180.0 - 122.9319185680026) * deg) @ _ry((90.0 - 27.12825118085622) * deg) @ _rz(192.8594812065348 * deg)
        return (g @ b).T
    raise ValueError(f"unsupported frame {frame}")
