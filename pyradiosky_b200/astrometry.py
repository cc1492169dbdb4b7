"""Host-side astrometry: the once-per-time-step context the position kernel consumes.

The reference passes astropy ``Time`` / ``EarthLocation`` objects to ``SkyModel.update_positions``
(src/pyradiosky/skymodel.py:2413-2454) and lets astropy build the ``erfa.apco`` context.  astropy is
not a dependency of this package (it is absent from the target image), so light-weight stand-ins are
provided here.  When astropy *is* importable, genuine astropy objects are accepted as well and the
real ``eraASTROM`` is taken from ``astropy.coordinates.erfa_astrom`` so results track astropy exactly.

All numerical work is done by librsk's host functions (``rsk_astrom_context``, ``rsk_ravg``,
``rsk_frame_to_icrs``); nothing here touches the GPU.
"""

import ctypes
import datetime as _dt
import math

import numpy as np

from . import _lib

ARCSEC = math.pi / (180.0 * 3600.0)


class EarthLocation:
    """Geodetic WGS84 site: east longitude and latitude in radians, height in metres."""

    def __init__(self, lon, lat, height=0.0):
        self.lon = float(lon)
        self.lat = float(lat)
        self.height = float(height)

    @classmethod
    def from_geodetic(cls, lon_deg, lat_deg, height=0.0):
        return cls(math.radians(lon_deg), math.radians(lat_deg), height)

    @classmethod
    def from_dms(cls, lat_dms, lon_dms, height=0.0):
        """e.g. ``from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)`` -- the HERA site of the
        reference's ``time_location`` fixture (tests/test_skymodel.py:37-42)."""

        def _deg(dms):
            d, m, s = dms
            sign = -1.0 if (d < 0 or (d == 0 and (m < 0 or s < 0))) else 1.0
            return sign * (abs(d) + abs(m) / 60.0 + abs(s) / 3600.0)

        return cls.from_geodetic(_deg(lon_dms), _deg(lat_dms), height)

    @property
    def geocentric(self):
        """Geocentric (x, y, z) in metres (astropy ``EarthLocation.geocentric``)."""
        out = (ctypes.c_double * 3)()
        _lib.check("rsk_geodetic_to_geocentric",
                   _lib.load().rsk_geodetic_to_geocentric(self.lon, self.lat, self.height, out))
        return np.array(out[:])

    def __eq__(self, other):
        return (
            isinstance(other, EarthLocation)
            and self.lon == other.lon
            and self.lat == other.lat
            and self.height == other.height
        )

    def __hash__(self):
        return hash((self.lon, self.lat, self.height))

    def __repr__(self):
        return (f"EarthLocation(lon={math.degrees(self.lon):.8f} deg, "
                f"lat={math.degrees(self.lat):.8f} deg, height={self.height} m)")


class Time:
    """UTC instant as a two-part Julian date plus the Earth-orientation parameters astropy would
    look up in its IERS tables: ``dut1`` = UT1-UTC (s), ``xp``/``yp`` polar motion (rad)."""

    def __init__(self, jd1, jd2=0.0, dut1=0.0, xp=0.0, yp=0.0):
        self.jd1 = float(jd1)
        self.jd2 = float(jd2)
        self.dut1 = float(dut1)
        self.xp = float(xp)
        self.yp = float(yp)

    @classmethod
    def from_utc(cls, text, iers=None, **eop):
        """Parse ``YYYY-MM-DD HH:MM:SS[.fff]`` (UTC).  ``iers``: an :class:`IERSTable` to take UT1-UTC and the polar
        motion from (explicit ``dut1`` / ``xp`` / ``yp`` keywords win)."""
        fmt = "%Y-%m-%d %H:%M:%S.%f" if "." in text else "%Y-%m-%d %H:%M:%S"
        t = _dt.datetime.strptime(text.replace("T", " "), fmt)
        a = (14 - t.month) // 12
        y = t.year + 4800 - a
        m = t.month + 12 * a - 3
        jdn = t.day + (153 * m + 2) // 5 + 365 * y + y // 4 - y // 100 + y // 400 - 32045
        frac = (t.hour * 3600 + t.minute * 60 + t.second + t.microsecond * 1e-6) / 86400.0
        if iers is not None:
            dut1, xp, yp = iers.eop(jdn - 0.5, frac)
            eop = {"dut1": dut1, "xp": xp, "yp": yp, **eop}
        return cls(jdn - 0.5, frac, **eop)

    @property
    def jd(self):
        return self.jd1 + self.jd2

    def __add__(self, days):
        return Time(self.jd1, self.jd2 + float(days), self.dut1, self.xp, self.yp)

    def __eq__(self, other):
        return (
            isinstance(other, Time)
            and (self.jd1, self.jd2, self.dut1, self.xp, self.yp)
            == (other.jd1, other.jd2, other.dut1, other.xp, other.yp)
        )

    def __hash__(self):
        return hash((self.jd1, self.jd2, self.dut1, self.xp, self.yp))

    def __repr__(self):
        return f"Time(jd_utc={self.jd1}+{self.jd2}, dut1={self.dut1}, xp={self.xp}, yp={self.yp})"


class IERSTable:
    """Earth-orientation parameters from an IERS ``finals2000A`` file (``finals2000A.all`` / ``.daily`` / ``.data``):
    the table astropy's ``iers.IERS_A`` reads to turn a UTC instant into UT1-UTC and polar motion before it builds the
    ``eraASTROM`` of ``AltAz(obstime, location)`` (the reference's call at skymodel.py:2451-2454; its tests pin the
    table in tests/conftest.py:10-26).

    Fixed-width columns, as in the IERS ``readme.finals2000A`` (1-based, inclusive): MJD 8-15, Bulletin A PM-x 19-27
    and PM-y 38-46 (arcsec), UT1-UTC 59-68 (s); Bulletin B PM-x 135-144, PM-y 145-154, UT1-UTC 155-165 where present.
    Like astropy, Bulletin B values are preferred where a row has them, rows without a UT1-UTC value are dropped,
    values are interpolated linearly in MJD with the leap-second steps of UT1-UTC removed before and restored after the
    interpolation, and instants outside the table raise.  No IERS file ships on this box, so the column layout is
    restated from the IERS readme and checked on synthetic rows only (tests/test_host.py): parity unpinned.
    """

    def __init__(self, mjd, dut1, xp_arcsec, yp_arcsec):
        order = np.argsort(np.asarray(mjd, dtype=np.float64))
        self.mjd = np.asarray(mjd, dtype=np.float64)[order]
        self.dut1 = np.asarray(dut1, dtype=np.float64)[order]
        self.xp = np.asarray(xp_arcsec, dtype=np.float64)[order]
        self.yp = np.asarray(yp_arcsec, dtype=np.float64)[order]
        if self.mjd.size < 2:
            raise ValueError("an IERS table needs at least two rows")

    @staticmethod
    def _field(line, first, last):
        text = line[first - 1:last].strip()
        return float(text) if text else None

    @classmethod
    def from_finals2000a(cls, path, prefer_bulletin_b=True):
        mjd, dut1, xp, yp = [], [], [], []
        with open(path) as fh:
            for line in fh:
                line = line.rstrip("\n")
                if len(line) < 68:
                    continue
                m = cls._field(line, 8, 15)
                ax, ay, au = cls._field(line, 19, 27), cls._field(line, 38, 46), cls._field(line, 59, 68)
                bx = by = bu = None
                if prefer_bulletin_b and len(line) >= 165:
                    bx, by, bu = cls._field(line, 135, 144), cls._field(line, 145, 154), cls._field(line, 155, 165)
                x, y, u = (bx, by, bu) if None not in (bx, by, bu) else (ax, ay, au)
                if None in (m, x, y, u):
                    continue
                mjd.append(m); xp.append(x); yp.append(y); dut1.append(u)
        if not mjd:
            raise ValueError(f"no Earth-orientation rows found in {path}")
        return cls(mjd, dut1, xp, yp)

    def eop(self, jd1, jd2=0.0):
        """(UT1-UTC [s], xp [rad], yp [rad]) at the UTC Julian date jd1 + jd2."""
        mjd = (jd1 - 2400000.5) + jd2
        if mjd < self.mjd[0] or mjd > self.mjd[-1]:
            raise ValueError(f"MJD {mjd:.3f} is outside the IERS table ({self.mjd[0]:.1f} .. {self.mjd[-1]:.1f})")
        i = int(np.clip(np.searchsorted(self.mjd, mjd, side="right") - 1, 0, self.mjd.size - 2))
        w = (mjd - self.mjd[i]) / (self.mjd[i + 1] - self.mjd[i])
        u0, u1 = self.dut1[i], self.dut1[i + 1]
        # a leap second between the two rows shows as a step of about +1 s in UT1-UTC: take it out of the later row,
        # interpolate, and put it back if the instant is past the step (which falls on the later row's midnight)
        step = np.round(u1 - u0)
        u = u0 + w * ((u1 - step) - u0)
        if step != 0.0 and mjd >= np.floor(self.mjd[i + 1]):
            u += step
        arcsec = np.pi / 180.0 / 3600.0
        x = self.xp[i] + w * (self.xp[i + 1] - self.xp[i])
        y = self.yp[i] + w * (self.yp[i + 1] - self.yp[i])
        return float(u), float(x * arcsec), float(y * arcsec)


_ASTROPY_TYPES = None


def _astropy_types():
    """(astropy Time, astropy EarthLocation) or (None, None); resolved once -- a failing import costs ~100 us and
    this is asked several times per time step."""
    global _ASTROPY_TYPES
    if _ASTROPY_TYPES is None:
        try:
            from astropy.coordinates import EarthLocation as AEL
            from astropy.time import Time as ATime

            _ASTROPY_TYPES = (ATime, AEL)
        except ImportError:
            _ASTROPY_TYPES = (None, None)
    return _ASTROPY_TYPES


def is_time(obj):
    atime, _ = _astropy_types()
    return isinstance(obj, Time) or (atime is not None and isinstance(obj, atime))


def is_earth_location(obj):
    _, ael = _astropy_types()
    return isinstance(obj, EarthLocation) or (ael is not None and isinstance(obj, ael))


def astrom_context(time, location):
    """The 31-double ``eraASTROM`` for (time, location) as a numpy array.

    astropy objects -> ``erfa_astrom.get().apco(AltAz(obstime, location))`` (exactly what the
    reference's transform uses).  Package objects -> librsk's medium-accuracy ``rsk_astrom_context``.
    """
    atime, ael = _astropy_types()
    if atime is not None and isinstance(time, atime) and isinstance(location, ael):
        from astropy.coordinates import AltAz
        from astropy.coordinates.erfa_astrom import erfa_astrom

        rec = erfa_astrom.get().apco(AltAz(obstime=time, location=location))
        return np.frombuffer(np.ascontiguousarray(rec).tobytes(), dtype=np.float64)[: _lib.ASTROM_LEN].copy()
    if not isinstance(time, Time) or not isinstance(location, EarthLocation):
        raise TypeError("time/location must both be pyradiosky_b200.astrometry (or both astropy) objects")
    out = np.empty(_lib.ASTROM_LEN, dtype=np.float64)
    _lib.check(
        "rsk_astrom_context",
        _lib.load().rsk_astrom_context(time.jd1, time.jd2, time.dut1, location.lon, location.lat,
                                       location.height, time.xp, time.yp, out.ctypes.data_as(_lib.c_double_p)),
    )
    return out


def astrom_contexts(times, location):
    """(T, 31) ``eraASTROM`` rows for a sequence of times at one location.

    Package ``Time`` objects that share their Earth-orientation parameters are built in one library call
    (``rsk_astrom_contexts``); anything else falls back to :func:`astrom_context` per time."""
    times = list(times)
    if (times and isinstance(location, EarthLocation) and all(isinstance(t, Time) for t in times)
            and len({(t.dut1, t.xp, t.yp) for t in times}) == 1):
        jd1 = np.ascontiguousarray([t.jd1 for t in times], dtype=np.float64)
        jd2 = np.ascontiguousarray([t.jd2 for t in times], dtype=np.float64)
        out = np.empty((len(times), _lib.ASTROM_LEN), dtype=np.float64)
        t0 = times[0]
        _lib.check("rsk_astrom_contexts", _lib.load().rsk_astrom_contexts(
            jd1.ctypes.data_as(_lib.c_double_p), jd2.ctypes.data_as(_lib.c_double_p), len(times), t0.dut1,
            location.lon, location.lat, location.height, t0.xp, t0.yp, out.ctypes.data_as(_lib.c_double_p)))
        return out
    return np.stack([astrom_context(t, location) for t in times]) if times else np.empty((0, _lib.ASTROM_LEN))


def average_rotation_matrix(astrom):
    """``SkyModel._calc_average_rotation_matrix`` (skymodel.py:2499-2555) for a context: (3, 3)."""
    a = _lib.host_array(np.asarray(astrom, dtype=np.float64).ravel())
    out = np.empty(9, dtype=np.float64)
    _lib.check("rsk_ravg", _lib.load().rsk_ravg(a, out.ctypes.data_as(_lib.c_double_p)))
    return out.reshape(3, 3)


_FRAMES = {"icrs": _lib.FRAME_ICRS, "fk5": _lib.FRAME_FK5, "galactic": _lib.FRAME_GALACTIC}


def frame_to_icrs_matrix(frame, equinox=2000.0):
    """3x3 matrix taking unit vectors of ``frame`` ('icrs', 'fk5', 'galactic') to ICRS."""
    name = str(frame).lower()
    if name not in _FRAMES:
        raise ValueError(f"frame '{frame}' is not supported; use one of {sorted(_FRAMES)}")
    out = (ctypes.c_double * 9)()
    _lib.check("rsk_frame_to_icrs", _lib.load().rsk_frame_to_icrs(_FRAMES[name], float(equinox), out))
    return np.array(out[:]).reshape(3, 3)


def pack_context(astrom, frame_to_icrs=None):
    """Device-ready context (``RSK_CTX_LEN`` doubles) for the sweep kernel."""
    a = _lib.host_array(np.asarray(astrom, dtype=np.float64).ravel())
    m = None if frame_to_icrs is None else _lib.host_array(np.asarray(frame_to_icrs, dtype=np.float64).ravel())
    out = (ctypes.c_double * _lib.CTX_LEN)()
    _lib.check("rsk_pack_context", _lib.load().rsk_pack_context(a, m, out))
    return np.array(out[:])


def pack_contexts(astroms, frame_to_icrs=None):
    """(T, RSK_CTX_LEN) device-ready contexts from (T, 31) eraASTROM rows, in one library call."""
    a = np.ascontiguousarray(astroms, dtype=np.float64).reshape(-1, _lib.ASTROM_LEN)
    out = np.empty((a.shape[0], _lib.CTX_LEN), dtype=np.float64)
    m = None if frame_to_icrs is None else _lib.host_array(np.asarray(frame_to_icrs, dtype=np.float64).ravel())
    _lib.check("rsk_pack_contexts", _lib.load().rsk_pack_contexts(
        a.ctypes.data_as(_lib.c_double_p), a.shape[0], m, out.ctypes.data_as(_lib.c_double_p)))
    return out
