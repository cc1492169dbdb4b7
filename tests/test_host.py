"""Host-side logic of the product that runs without a GPU: the C-ABI library loads and exports every
symbol of include/rsk.h and include/rsk_io.h, the C++ context builder agrees with the oracle, argument errors."""

import ctypes
import os
import re

import numpy as np
import pytest
import scipy.interpolate

from oracle import erfa_np as E
from oracle import skymodel_np as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(librsk):
    header = open(os.path.join(ROOT, "include", "rsk.h")).read() + open(os.path.join(ROOT, "include", "rsk_io.h")).read()
    declared = set(re.findall(r"\b(rsk_[a-z_0-9]+)\s*\(", header))
    declared = {d for d in declared if not d.endswith("_") and d != "rsk_pack_context()"}
    from pyradiosky_b200 import _lib

    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(librsk, name), name
    assert librsk.rsk_version() == 200
    assert librsk.rsk_strerror(0) == b"success"
    assert b"pointer" in librsk.rsk_strerror(-1)


def test_struct_layouts_match_header():
    from pyradiosky_b200 import _lib

    assert ctypes.sizeof(_lib.RskCoords) == 56
    assert ctypes.sizeof(_lib.RskSpectrum) == 168
    assert _lib.ASTROM_LEN == 31 and _lib.CTX_LEN == 34


def test_argument_errors_without_gpu(librsk):
    """Argument validation happens before any CUDA call, so it is testable on a CPU-only box."""
    from pyradiosky_b200 import _lib

    d9 = (ctypes.c_double * 9)()
    assert librsk.rsk_frame_to_icrs(99, 2000.0, d9) == -3
    assert librsk.rsk_ravg(None, d9) == -1
    coords = _lib.RskCoords()
    coords.nside = 0
    a = (ctypes.c_double * 31)()
    assert librsk.rsk_positions(ctypes.byref(coords), 10, None, a, None, None, None, None, None) == -1  # lon/lat NULL
    coords.nside = 12
    coords.nested = 1
    assert librsk.rsk_positions(ctypes.byref(coords), 10, None, a, None, None, None, None, None) == -7  # nested needs 2^k
    spec = _lib.RskSpectrum()
    spec.mode = 7
    spec.stokes = 8
    spec.nf_in = 1
    spec.n = 4
    assert librsk.rsk_coherency(ctypes.byref(spec), ctypes.byref(coords), None, None, None, None, 4, None, None, 1, None, 4, 0, None) == -3
    spec.mode = _lib.FLAT
    assert librsk.rsk_coherency(ctypes.byref(spec), ctypes.byref(coords), None, None, None, None, 4, None, None, 1, ctypes.c_void_p(8), 2, 0, None) == -4
    spec.mode = _lib.SUBBAND
    spec.degree = 5
    spec.breaks = 8
    spec.ppoly = 8
    spec.nb = 3
    assert librsk.rsk_coherency(ctypes.byref(spec), ctypes.byref(coords), None, None, None, None, 4, None, ctypes.c_void_p(8), 1, ctypes.c_void_p(8), 4, 0, None) == -6
    assert librsk.rsk_compact_index(ctypes.c_void_p(8), 100, ctypes.c_void_p(8), ctypes.c_void_p(8), None, 0, None) == -5
    assert librsk.rsk_workspace_bytes(_lib.OP_COMPACT, 5000, 0) == (3 + 1) * 8
    # zero-sized problems are accepted without touching the device
    assert librsk.rsk_polarized_mask(None, 1, 0, None, None) == 0
    assert librsk.rsk_convert_units(None, 0, 0, None, None, 0, None) == 0


def test_context_builder_matches_oracle(librsk, hera):
    from pyradiosky_b200 import astrometry

    lon, lat, h, jd = hera
    loc = astrometry.EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), h)
    assert np.isclose(loc.lon, lon, rtol=0, atol=1e-16) and np.isclose(loc.lat, lat, rtol=0, atol=1e-16)
    t0 = astrometry.Time.from_utc("2015-03-01 00:00:00")
    assert t0.jd == jd
    for frac, dut1, xp, yp in [(0.0, 0.0, 0.0, 0.0), (0.37, -0.47, 1e-7, 1.5e-6), (200.25, 0.3, -2e-7, 1e-6)]:
        t = astrometry.Time(jd, frac, dut1=dut1, xp=xp, yp=yp)
        got = astrometry.astrom_context(t, loc)
        ref = E.astrom_context(jd, frac, lon, lat, h, dut1=dut1, xp=xp, yp=yp)
        assert got.shape == (31,)
        assert np.abs(got - ref).max() < 2e-15
        ravg = astrometry.average_rotation_matrix(got)
        assert np.abs(ravg - S.calc_average_rotation_matrix(ref)).max() < 3e-15
        assert np.abs(ravg @ ravg.T - np.eye(3)).max() < 1e-15
    assert np.allclose(loc.geocentric, [5109342.76, 2005241.904, -3239939.469], atol=1e-3)  # tutorial.rst:1233


def test_time_parsing():
    from pyradiosky_b200.astrometry import Time

    assert Time.from_utc("2000-01-01 12:00:00").jd == 2451545.0
    assert Time.from_utc("2015-03-05 00:00:00").jd == 2457086.5
    assert abs(Time.from_utc("2018-03-01 18:00:00.500").jd2 - (18 * 3600 + 0.5) / 86400) < 1e-15
    t = Time.from_utc("2015-03-01 00:00:00")
    assert (t + 0.25).jd == t.jd + 0.25 and t == Time(2457082.5) and t != t + 1e-9


def test_frame_matrices(librsk):
    from pyradiosky_b200.astrometry import frame_to_icrs_matrix

    assert np.array_equal(frame_to_icrs_matrix("icrs"), np.eye(3))
    fk5 = frame_to_icrs_matrix("fk5")
    gal = frame_to_icrs_matrix("galactic")
    for m in (fk5, gal, frame_to_icrs_matrix("fk5", 1950.0)):
        assert np.abs(m @ m.T - np.eye(3)).max() < 1e-15 and np.isclose(np.linalg.det(m), 1)
    # frame bias off-diagonals (-22.9, -9.1, +19.9) mas (IERS conventions)
    mas = np.pi / 180 / 3600e3
    to_fk5 = fk5.T
    assert np.isclose(to_fk5[0, 1] / mas, -22.9, atol=1e-3)
    assert np.isclose(to_fk5[0, 2] / mas, -9.1, atol=1e-3)
    assert np.isclose(to_fk5[1, 2] / mas, 19.9, atol=1e-3)
    # galactic north pole and centre in FK5 J2000
    pole = fk5.T @ gal @ np.array([0, 0, 1.0])
    assert np.isclose(np.rad2deg(np.arctan2(pole[1], pole[0])) % 360, 192.8594812, atol=1e-6)
    assert np.isclose(np.rad2deg(np.arcsin(pole[2])), 27.1282512, atol=1e-6)
    centre = fk5.T @ gal @ np.array([1.0, 0, 0])
    assert np.isclose(np.rad2deg(np.arctan2(centre[1], centre[0])) % 360, 266.40500, atol=1e-4)
    assert np.isclose(np.rad2deg(np.arcsin(centre[2])), -28.93617, atol=1e-4)
    # precession 2000 -> 1950: ~0.64 deg in RA of the pole-ward axes
    p = frame_to_icrs_matrix("fk5", 1950.0) @ fk5.T
    assert 0.0110 < abs(p[0, 1]) < 0.0113
    with pytest.raises(ValueError, match="not supported"):
        frame_to_icrs_matrix("fk4")


def test_packed_context_reproduces_chain(librsk, hera):
    """rsk_pack_context merges bpn, Earth rotation and polar motion into one matrix: it must map the
    aberrated direction to the same (-HA, Dec) vector as the step-by-step ERFA chain."""
    from pyradiosky_b200 import astrometry

    lon, lat, h, jd = hera
    a = E.astrom_context(jd, 0.2, lon, lat, h, dut1=-0.4, xp=2e-7, yp=1.4e-6)
    ctx = astrometry.pack_context(a)
    assert ctx.shape == (34,)
    a1 = ctx[8:17].reshape(3, 3)
    rng = np.random.default_rng(2)
    p = rng.normal(size=(3, 50)); p /= np.linalg.norm(p, axis=0)
    pi = E.rxp(a[E.I_BPN:E.I_BPN + 9], p)
    ri, di = E.c2s(pi)
    v = E.s2c(E.anp(ri) - a[E.I_ERAL], di)
    sx, cx = np.sin(a[E.I_XPL]), np.cos(a[E.I_XPL]); sy, cy = np.sin(a[E.I_YPL]), np.cos(a[E.I_YPL])
    hd = np.array([cx * v[0] + sx * v[2], sx * sy * v[0] + cy * v[1] - cx * sy * v[2], -sx * cy * v[0] + sy * v[1] + cx * cy * v[2]])
    # no diurnal aberration and no refraction in this context: the latitude rotation is folded in as well
    sphi, cphi = a[E.I_SPHI], a[E.I_CPHI]
    ae = np.array([sphi * hd[0] - cphi * hd[2], hd[1], cphi * hd[0] + sphi * hd[2]])
    assert np.abs(a1 @ p - ae).max() < 1e-15
    assert ctx[18] == 1.0 and ctx[19] == 0.0  # (sphi, cphi)
    # a context with a diurnal-aberration term is left unfolded
    b = a.copy()
    b[E.I_DIURAB] = 1e-6
    ctxb = astrometry.pack_context(b)
    assert ctxb[18] == sphi and ctxb[19] == cphi and ctxb[20] == 1e-6
    assert np.abs(ctxb[8:17].reshape(3, 3) @ p - hd).max() < 1e-15
    assert ctx[3] == E.SRS / a[E.I_EM] and ctx[24] == 1.0  # (w2, m_ident)
    batch = astrometry.pack_contexts(np.stack([a, a]))
    assert batch.shape == (2, 34) and np.array_equal(batch[0], ctx) and np.array_equal(batch[1], ctx)
    assert ctx[17] == 1.0 / (1.0 + a[E.I_BM1]) and ctx[23] == a[E.I_EM]
    m = astrometry.frame_to_icrs_matrix("fk5")
    ctx2 = astrometry.pack_context(a, m)
    assert ctx2[24] == 0.0 and np.array_equal(ctx2[25:34].reshape(3, 3), m)


@pytest.mark.parametrize("kind", ["linear", "slinear", "quadratic", "cubic", 1, 2, 3])
def test_bspline_operator_matches_interp1d(kind):
    from pyradiosky_b200 import spectral

    x = np.array([76, 84, 92, 99, 107, 115, 122, 130, 143, 151, 158, 166, 174, 181, 189, 197, 204, 212, 220, 227]) * 1e6
    rng = np.random.default_rng(1)
    y = rng.normal(size=(20, 7)) + 3
    f = np.concatenate((np.linspace(x[0], x[-1], 555), x))
    degree, breaks, cmat, first, pband = spectral.build_bspline_operator(x, kind)
    assert cmat.shape[1] == 20 and pband.shape == (breaks.size - 1, degree + 1, degree + 1)
    assert first.min() >= 0 and first.max() + degree < cmat.shape[0]
    ref = scipy.interpolate.interp1d(x, y, axis=0, kind=kind)(f)
    assert np.abs(spectral.evaluate_bspline_operator(degree, breaks, cmat, first, pband, y, f) - ref).max() < 2e-14
    perm = rng.permutation(20)
    op = spectral.build_bspline_operator(x[perm], kind)
    ref = scipy.interpolate.interp1d(x[perm], y[perm], axis=0, kind=kind)(f)
    assert np.abs(spectral.evaluate_bspline_operator(*op, y[perm], f) - ref).max() < 2e-14


@pytest.mark.parametrize("kind", ["linear", "slinear", "quadratic", "cubic", 1, 2, 3])
def test_ppoly_operator_matches_interp1d(kind):
    from pyradiosky_b200 import spectral

    x = np.array([76, 84, 92, 99, 107, 115, 122, 130, 143, 151, 158, 166, 174, 181, 189, 197, 204, 212, 220, 227]) * 1e6
    rng = np.random.default_rng(0)
    y = rng.normal(size=(20, 7)) + 3
    f = np.linspace(x[0], x[-1], 777)
    degree, breaks, ppoly = spectral.build_ppoly_operator(x, kind)
    assert ppoly.shape == (breaks.size - 1, degree + 1, 20)
    ref = scipy.interpolate.interp1d(x, y, axis=0, kind=kind)(f)
    assert np.abs(spectral.evaluate_operator(degree, breaks, ppoly, y, f) - ref).max() < 1e-13
    perm = rng.permutation(20)  # unsorted freq_array, as interp1d tolerates
    degree, breaks, ppoly = spectral.build_ppoly_operator(x[perm], kind)
    ref = scipy.interpolate.interp1d(x[perm], y[perm], axis=0, kind=kind)(f)
    assert np.abs(spectral.evaluate_operator(degree, breaks, ppoly, y[perm], f) - ref).max() < 1e-13


@pytest.mark.parametrize("kind", ["nearest", "nearest-up", "zero", "previous", "next", 0])
def test_step_operator_matches_interp1d(kind):
    """The channel-picking interp1d kinds (skymodel.py:2362-2368 passes any kind through): bit-equal to scipy at the
    knots, at the midpoints between knots, one ulp either side of both, and at random frequencies."""
    import scipy.interpolate

    from pyradiosky_b200 import spectral

    x = np.array([76, 84, 92, 99, 107, 115, 122, 130, 143, 151, 158, 166, 174, 181, 189, 197, 204, 212, 220, 227]) * 1e6
    rng = np.random.default_rng(3)
    y = rng.normal(size=(20, 5))
    mids = 0.5 * (x[1:] + x[:-1])
    f = np.concatenate([x, mids, np.nextafter(x[:-1], np.inf), np.nextafter(x[1:], 0), np.nextafter(mids, np.inf),
                        np.nextafter(mids, 0), rng.uniform(x[0], x[-1], 3000)])
    for xs, ys in ((x, y), (x[::-1].copy(), y[::-1].copy())):
        degree, breaks, ppoly = spectral.build_ppoly_operator(xs, kind)
        assert degree == 0 and ppoly.shape == (breaks.size - 1, 1, 20)
        ref = scipy.interpolate.interp1d(xs, ys, axis=0, kind=kind)(f)
        assert np.array_equal(spectral.evaluate_operator(degree, breaks, ppoly, ys, f), ref)


def test_ppoly_operator_errors():
    from pyradiosky_b200 import spectral

    with pytest.raises(NotImplementedError):
        spectral.build_ppoly_operator(np.linspace(1, 2, 10), "akima")
    with pytest.raises(NotImplementedError):
        spectral.build_ppoly_operator(np.linspace(1, 2, 10), 5)
    with pytest.raises(ValueError):
        spectral.build_ppoly_operator(np.linspace(1, 2, 3), "cubic")


def test_no_cpu_fallback(librsk):
    """Without a CUDA device the product must refuse to run rather than fall back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from pyradiosky_b200 import SkyModel

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SkyModel(ra=[0.1], dec=[0.2], stokes=np.ones((4, 1, 1)), spectral_type="flat")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pyradiosky_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


@pytest.mark.filterwarnings("ignore:Some Stokes values are NaNs")
@pytest.mark.parametrize("kind", ["cubic", "linear"])
@pytest.mark.parametrize("nan_handling", ["clip", "interp", "propagate"])
def test_nan_group_operator_matches_reference_branch(kind, nan_handling):
    """spectral.build_nan_group_operator == the per-component NaN branch of at_frequencies
    (skymodel.py:2245-2312, via the oracle) for a range of valid-channel patterns."""
    from pyradiosky_b200 import spectral

    x = np.array([76, 84, 92, 99, 107, 115, 122, 130, 143, 151, 158, 166, 174, 181, 189, 197, 204, 212, 220, 227]) * 1e6
    rng = np.random.default_rng(3)
    f = np.concatenate((np.linspace(76e6, 227e6, 211), x))
    for use in (np.arange(2, 20), np.arange(0, 17), np.array([0, 1, 2, 5, 6, 9, 15, 19]), np.array([3, 7]),
                np.array([4, 8, 11]), np.array([5]), np.array([], dtype=int)):
        st = rng.normal(size=(4, 20, 1)) + 3
        gone = np.ones(20, bool)
        gone[use] = False
        st[rng.integers(0, 4), gone, 0] = np.nan
        ref = S.at_frequencies(st, "subband", f, freq_array=x, freq_interp_kind=kind, nan_handling=nan_handling)[:, :, 0].T
        degree, breaks, ppoly, chan, _ = spectral.build_nan_group_operator(x, use, kind, nan_handling)
        with np.errstate(all="ignore"):
            got = spectral.evaluate_operator(degree, breaks, ppoly, st[:, chan, 0].T, f)
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        ok = np.isfinite(ref)
        assert np.abs(got[ok] - ref[ok]).max() < 1e-13 if ok.any() else True
    with pytest.raises(ValueError, match="nan_handling must be one of"):
        spectral.build_nan_group_operator(x, np.arange(20), kind, "foo")


def test_batched_contexts_equal_single_calls(librsk):
    """rsk_astrom_contexts == rsk_astrom_context per time, bit for bit; mixed inputs fall back per time."""
    from pyradiosky_b200 import EarthLocation, Time, astrometry

    loc = EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
    t0 = Time(2457082.5, 0.0, dut1=-0.3, xp=1e-7, yp=2e-7)
    times = [t0 + k * 0.01 for k in range(37)]
    batch = astrometry.astrom_contexts(times, loc)
    assert batch.shape == (37, 31)
    for k in (0, 1, 17, 36):
        assert np.array_equal(batch[k], astrometry.astrom_context(times[k], loc))
    mixed = times[:3] + [Time(2457083.5, 0.0, dut1=0.1)]
    assert np.array_equal(astrometry.astrom_contexts(mixed, loc)[3], astrometry.astrom_context(mixed[3], loc))
    assert astrometry.astrom_contexts([], loc).shape == (0, 31)


def _finals_row(mjd, pmx, pmy, dut1, bull_b=None):
    """One line of finals2000A in the column layout of the IERS readme (1-based: MJD 8-15, PM-x 19-27, PM-y 38-46,
    UT1-UTC 59-68; Bulletin B 135-144, 145-154, 155-165)."""
    line = [" "] * 188
    def put(first, last, text):
        width = last - first + 1
        line[first - 1:last] = list(text.rjust(width)[-width:])
    put(1, 6, "150301")
    put(8, 15, f"{mjd:.2f}")
    put(17, 17, "I")
    put(19, 27, f"{pmx:.6f}")
    put(28, 36, f"{0.00004:.6f}")
    put(38, 46, f"{pmy:.6f}")
    put(47, 55, f"{0.00005:.6f}")
    put(58, 58, "I")
    put(59, 68, f"{dut1:.7f}")
    put(69, 78, f"{0.00001:.7f}")
    if bull_b is not None:
        put(135, 144, f"{bull_b[0]:.6f}")
        put(145, 154, f"{bull_b[1]:.6f}")
        put(155, 165, f"{bull_b[2]:.7f}")
    return "".join(line).rstrip()


def test_iers_table(tmp_path):
    """IERSTable: fixed-width parsing, Bulletin B preferred, linear interpolation, the leap second of 2015-07-01
    (MJD 57204) in UT1-UTC, range errors, Time.from_utc(iers=...)."""
    from pyradiosky_b200 import IERSTable, Time

    rows = [
        _finals_row(57082.0, 0.010000, 0.300000, -0.4800000, bull_b=(0.011000, 0.301000, -0.4801000)),
        _finals_row(57083.0, 0.012000, 0.304000, -0.4812000, bull_b=(0.013000, 0.305000, -0.4813000)),
        _finals_row(57084.0, 0.014000, 0.308000, -0.4824000),  # no Bulletin B yet: Bulletin A is used
        _finals_row(57202.0, 0.100000, 0.400000, -0.6750000),
        _finals_row(57203.0, 0.102000, 0.402000, -0.6760000),
        _finals_row(57204.0, 0.104000, 0.404000, 0.3230000),   # after the leap second
        _finals_row(57205.0, 0.106000, 0.406000, 0.3220000),
        "150709 57212.00 P                                                        ",  # prediction row without values
    ]
    path = tmp_path / "finals2000A.test"
    path.write_text("\n".join(rows) + "\n")
    tab = IERSTable.from_finals2000a(str(path))
    assert tab.mjd.size == 7
    arcsec = np.pi / 180 / 3600
    u, x, y = tab.eop(2400000.5, 57082.0)
    assert (u, x, y) == (-0.4801, 0.011 * arcsec, 0.301 * arcsec)
    u, x, y = tab.eop(2400000.5 + 57082.0, 0.25)
    assert abs(u - (-0.4801 - 0.25 * 0.0012)) < 1e-12 and abs(x - 0.0115 * arcsec) < 1e-18
    u, x, y = tab.eop(2400000.5 + 57083.0, 0.5)   # between a Bulletin B row and a Bulletin A row
    assert abs(u - 0.5 * (-0.4813 - 0.4824)) < 1e-12 and abs(y - 0.5 * (0.305 + 0.308) * arcsec) < 1e-18
    # across the leap second the interpolant follows UT1-UTC without the step until midnight
    u, _, _ = tab.eop(2400000.5 + 57203.0, 0.75)
    assert abs(u - (-0.6760 + 0.75 * (-0.6770 + 0.6760))) < 1e-12
    u, _, _ = tab.eop(2400000.5 + 57204.0, 0.0)
    assert u == 0.3230
    u, _, _ = tab.eop(2400000.5 + 57204.0, 0.5)
    assert abs(u - 0.3225) < 1e-12
    with pytest.raises(ValueError, match="outside the IERS table"):
        tab.eop(2400000.5 + 57300.0)
    tab_a = IERSTable.from_finals2000a(str(path), prefer_bulletin_b=False)
    assert tab_a.eop(2400000.5, 57082.0)[0] == -0.48
    t = Time.from_utc("2015-03-01 06:00:00", iers=tab)
    assert abs(t.dut1 - (-0.4801 - 0.25 * 0.0012)) < 1e-12 and abs(t.xp - 0.0115 * arcsec) < 1e-18
    t2 = Time.from_utc("2015-03-01 06:00:00", iers=tab, dut1=0.1)
    assert t2.dut1 == 0.1 and t2.xp == t.xp
    (tmp_path / "empty").write_text("nothing here\n")
    with pytest.raises(ValueError, match="no Earth-orientation rows"):
        IERSTable.from_finals2000a(str(tmp_path / "empty"))


def test_headers_compile_as_c_and_layouts_match_ctypes(tmp_path):
    """include/rsk.h and include/rsk_io.h are plain C (what a cgo / JNI / ctypes binding would consume): a C program
    that includes them reports every struct's size and field offsets, which must be what pyradiosky_b200._lib declares."""
    import subprocess

    from pyradiosky_b200 import _lib

    structs = {"RskCoords": _lib.RskCoords, "RskSpectrum": _lib.RskSpectrum, "RskBeam": _lib.RskBeam,
               "RskSelect": _lib.RskSelect, "RskH5Info": _lib.RskH5Info}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "rsk.h"', '#include "rsk_io.h"', "int main(void) {"]
    for name, cls in structs.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for field, _ in cls._fields_:
            lines.append(f'  printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)],
                   check=True, capture_output=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for name, cls in structs.items():
        assert int(out[name]) == ctypes.sizeof(cls), name
        for field, _ in cls._fields_:
            assert int(out[f"{name}.{field}"]) == getattr(cls, field).offset, f"{name}.{field}"


def test_bench_catalogue_is_the_same_for_every_rank_count():
    """bench.py shards ONE catalogue (strong scaling): the blocks of 2, 4 and 8 ranks concatenate to the single-rank
    catalogue, for point and subband workloads, and the GPU and reference arms describe the same config."""
    import bench

    for wl in ("c2", "c1"):
        n = bench.WORKLOADS[wl][1]
        whole = bench.synth_catalog(wl)
        for world in (2, 4, 8):
            parts = [bench.synth_catalog(wl, 2, *bench.shard_bounds(n, world, r)) for r in range(world)]
            assert sum(p["n"] for p in parts) == n
            for key in ("stokes", "lon", "lat"):
                assert np.array_equal(np.concatenate([p[key] for p in parts], axis=-1), whole[key])
    n = bench.WORKLOADS["c4"][1]
    lo8, hi8 = bench.shard_bounds(n, 8, 5)
    lo2, hi2 = bench.shard_bounds(n, 2, 1)
    a, b = bench.synth_catalog("c4", 2, lo8, hi8), bench.synth_catalog("c4", 2, lo2, hi2)
    assert a["stokes"].shape == (4, 20, n // 8)
    assert np.array_equal(a["stokes"], b["stokes"][..., lo8 - lo2:hi8 - lo2])
    assert np.array_equal(a["lon"], b["lon"][lo8 - lo2:hi8 - lo2])
    cfg1, cfg8 = bench.config_record("n1m", 1), bench.config_record("n1m", 8)
    assert set(cfg1) == set(cfg8) and cfg8["components_per_gpu"] == 125_000 and cfg1["components"] == 1_000_000
    # roofline bytes: 64 B per evaluation dominate; the stored-cube path counts 96
    cat = bench.synth_catalog("c2", 2, 0, 1000)
    assert bench.k2_algorithmic_bytes(cat, 500) == 64 * 1024 * 500 + (8 + 1 + 24 + 32 + 16) * 500 + 8 * 1024
    assert bench.k2_algorithmic_bytes(cat, 500, stored_full=True) == 96 * 1024 * 500 + 33 * 500


def test_bench_traffic_record_is_never_stale(tmp_path, monkeypatch):
    """roofline.traffic comes from the committed ncu capture only while the kernel sources hash to the capture's."""
    import json

    import bench

    rec = bench.traffic_record("n1m")
    assert set(rec) == {"traffic", "traffic_source"}
    entry = json.load(open(os.path.join(bench.ROOT, "profiles", "k_eval_traffic.json")))["n1m"]
    if rec["traffic"] is not None:
        assert rec["traffic"] == entry["traffic_bytes"] and abs(rec["traffic"] / entry["algorithmic_bytes"] - 1) < 0.01
    assert bench.traffic_record("c1")["traffic"] is None
