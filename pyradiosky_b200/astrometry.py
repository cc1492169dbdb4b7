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
    def from_utc(cls, text, **eop):
        """Parse ``YYYY-MM-DD HH:MM:SS[.fff]`` (UTC)."""
        fmt = "%Y-%m-%d %H:%M:%S.%f" if "." in text else "%Y-%m-%d %H:%M:%S"
        t = _dt.datetime.strptime(text.replace("T", " "), fmt)
        a = (14 - t.month) // 12
        y = t.year + 4800 - a
        m = t.month + 12 * a - 3
        jdn = t.day + (153 * m + 2) // 5 + 365 * y + y // 4 - y // 100 + y // 400 - 32045
        frac = (t.hour * 3600 + t.minute * 60 + t.second + t.microsecond * 1e-6) / 86400.0
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
