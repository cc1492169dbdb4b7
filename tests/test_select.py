"""Component down-selection: SkyModel.select / cut_nonrising / calculate_rise_set_lsts.

CPU part: the numpy oracle (oracle/select_np.py) against the reference's own known answers.
GPU part (``-m gpu``): the reference's tests for these methods (tests/test_skymodel.py:2174-2760), line for line where
astropy types are not essential, on the CUDA path (rsk_select_mask -> rsk_compact_index -> rsk_gather_rows,
rsk_rise_set_lsts), plus random parity against the oracle and the C-ABI error paths.
"""

import ctypes
import os
import warnings

import numpy as np
import pytest

from oracle import erfa_np as E
from oracle import h5_np
from oracle import healpix_np as hp
from oracle import select_np as SEL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLEAM50 = os.path.join(ROOT, "tests", "golden", "ref_fixture_gleam50_old_header.skyh5")


def _gleam_with_nans():
    f = h5_np.File(GLEAM50)
    st = np.array(f["/Data/stokes"].read(), dtype=np.float64)
    # tests/test_skymodel.py:2181-2185
    st[0, 0:2, 0] = np.nan
    st[0, 10:11, 1] = np.nan
    st[0, -2:, 2] = np.nan
    st[0, :, 3] = np.nan
    lon = np.deg2rad(np.array(f["/Header/ra"].read(), dtype=np.float64))
    lat = np.deg2rad(np.array(f["/Header/dec"].read(), dtype=np.float64))
    fa = np.array(f["/Header/freq_array"].read(), dtype=np.float64)
    return lon, lat, st, fa


# ---------------------------------------------------------------------------------------------------
# CPU: the oracle is pinned by the reference's known answers
# ---------------------------------------------------------------------------------------------------
def test_oracle_reproduces_reference_counts():
    """tests/test_skymodel.py:2193-2223: 46 / 49 / 32 / 31 / 29 / 4 components survive."""
    lon, lat, st, fa = _gleam_with_nans()
    kw = dict(spectral_type="subband", freq_array=fa)
    assert SEL.select_indices(lon, lat, st, non_nan="any", **kw).size == 46
    assert SEL.select_indices(lon, lat, st, non_nan="all", **kw).size == 49
    assert SEL.select_indices(lon, lat, st, non_nan=None, non_negative=True, **kw).size == 32
    assert SEL.select_indices(lon, lat, st, non_nan="all", non_negative=True, **kw).size == 31
    assert SEL.select_indices(lon, lat, st, non_nan="any", non_negative=True, **kw).size == 29
    assert SEL.select_indices(lon, lat, st, component_inds=np.arange(10), non_nan="any", non_negative=True, **kw).size == 4
    assert SEL.select_indices(lon, lat, st, **kw) is None


def test_oracle_rise_set_identities():
    """tests/test_skymodel.py:2726-2737: circumpolar <=> tan tan > 1 (set/rise NaN), never rising <=> < -1."""
    rng = np.random.default_rng(3)
    lon = rng.uniform(0, 2 * np.pi, 2000)
    lat = np.arcsin(rng.uniform(-1, 1, 2000))
    tel = -0.5361
    rise, sett = SEL.rise_set_lsts(lon, lat, tel)
    tans = np.tan(tel) * np.tan(lat)
    assert np.array_equal(np.isnan(rise), np.abs(tans) > 1) and np.array_equal(np.isnan(sett), np.abs(tans) > 1)
    ok = ~np.isnan(rise)
    assert rise[ok].min() >= 0 and rise[ok].max() <= 2 * np.pi and sett[ok].min() >= 0 and sett[ok].max() <= 2 * np.pi
    keep = SEL.nonrising_keep(lat, tel)
    assert np.array_equal(keep, np.nonzero(~(tans < -1))[0])
    # the time above the (buffered) horizon is 2 acos(-tans) + 2 buffer
    up = np.mod(sett[ok] - rise[ok], 2 * np.pi)
    want = np.mod(2 * np.arccos(-tans[ok]) + 2 * 0.04364, 2 * np.pi)
    assert np.abs(up - want).max() < 1e-12


# ---------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def rsk(librsk):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pyradiosky_b200

    return pyradiosky_b200


@pytest.fixture(scope="module")
def site(rsk):
    loc = rsk.EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
    time = rsk.Time.from_utc("2015-03-01 00:00:00")
    return time, loc


def _gleam_model(rsk, with_nans=True):
    lon, lat, st, fa = _gleam_with_nans()
    if not with_nans:
        st = np.array(h5_np.File(GLEAM50)["/Data/stokes"].read(), dtype=np.float64)
    names = np.array([f"GLEAM {k}" for k in range(st.shape[2])])
    sky = rsk.SkyModel(name=names, ra=lon, dec=lat, frame="fk5", stokes=st, spectral_type="subband", freq_array=fa,
                       history="read from the fixture.")
    return sky, lon, lat, st, fa


@pytest.mark.gpu
def test_cut_nan_neg(rsk):
    """tests/test_skymodel.py:2174-2223."""
    sky, *_ = _gleam_model(rsk)
    assert sky.select(non_nan="any", inplace=False).Ncomponents == 46
    assert sky.select(non_nan="all", inplace=False).Ncomponents == 49
    assert sky.select(non_nan=None, non_negative=True, inplace=False).Ncomponents == 32
    assert sky.select(non_nan="all", non_negative=True, inplace=False).Ncomponents == 31
    assert sky.select(non_nan="any", non_negative=True, inplace=False).Ncomponents == 29
    sky3 = sky.select(component_inds=np.arange(10), non_nan="any", non_negative=True, inplace=False)
    assert sky3.Ncomponents == 4
    assert sky.Ncomponents == 50  # inplace=False left the object alone


def _assert_same_model(a, b):
    import torch

    assert a.Ncomponents == b.Ncomponents and a.Nfreqs == b.Nfreqs
    for attr in ("stokes", "lon", "lat", "spectral_index", "reference_frequency", "stokes_error", "beam_amp", "alt_az",
                 "pos_lmn", "above_horizon", "hpx_inds"):
        x, y = getattr(a, attr), getattr(b, attr)
        assert (x is None) == (y is None), attr
        if x is not None:
            assert torch.equal(torch.nan_to_num(x.double(), nan=-7.0), torch.nan_to_num(y.double(), nan=-7.0)), attr
    assert (a.name is None) == (b.name is None) and (a.name is None or np.array_equal(a.name, b.name))
    assert (a.extended_model_group is None) == (b.extended_model_group is None)
    assert (a.extra_columns is None) == (b.extra_columns is None)
    if a.extra_columns:
        for k in a.extra_columns:
            assert np.array_equal(a.extra_columns[k], b.extra_columns[k])


@pytest.mark.gpu
def test_flux_source_cuts(rsk):
    """tests/test_skymodel.py:2226-2258."""
    sky, *_ = _gleam_model(rsk, with_nans=False)
    sky.select(non_negative=True)
    sky2 = sky.select(min_brightness=0.2, max_brightness=1.5, inplace=False)
    assert sky2.Ncomponents < sky.Ncomponents
    s0 = sky2.stokes[0].cpu().numpy()
    assert np.all((s0 > 0.2) & (s0 < 1.5))
    st = sky.stokes[0].cpu().numpy()
    keep = np.where((st.min(axis=0) > 0.2) & (st.max(axis=0) < 1.5))[0]
    sky3 = sky.select(component_inds=keep, inplace=False)
    expected = sky.history + "  Downselected to specific components using pyradiosky."
    assert sky2.history == expected and sky3.history == expected
    _assert_same_model(sky2, sky3)


@pytest.mark.gpu
def test_select(rsk, site):
    """tests/test_skymodel.py:2261-2285: every per-component attribute follows, also the time-dependent ones."""
    import torch

    time, loc = site
    sky, *_ = _gleam_model(rsk, with_nans=False)
    sky.select(non_negative=True)
    n = sky.Ncomponents
    sky.beam_amp = torch.ones((4, sky.Nfreqs, n), dtype=torch.float64, device=sky.device)
    sky.stokes_error = sky.stokes * 0.1
    sky.extended_model_group = np.full(n, "", dtype="<U10")
    sky.update_positions(time, loc)
    sky.extra_columns = {"foo": np.arange(n, dtype=float), "bar": np.arange(n, dtype=int),
                         "gah": np.array(["gah " + str(i) for i in range(n)])}
    before = {k: getattr(sky, k).cpu().numpy().copy() for k in ("stokes", "lon", "lat", "alt_az", "pos_lmn", "above_horizon")}
    sky2 = sky.select(component_inds=np.arange(10), inplace=False)
    sky.select(component_inds=np.arange(10), run_check=False)
    _assert_same_model(sky, sky2)
    assert sky.extended_model_group is None and sky.Ncomponents == 10
    for k, v in before.items():
        assert np.array_equal(getattr(sky, k).cpu().numpy(), v[..., :10], equal_nan=True), k
    assert np.array_equal(sky.extra_columns["gah"], np.array(["gah " + str(i) for i in range(10)]))
    full = np.array(h5_np.File(GLEAM50)["/Data/stokes"].read(), dtype=np.float64)
    nonneg = np.nonzero(~np.any(full[0] < 0, axis=0))[0]
    assert np.array_equal(sky.name, np.array([f"GLEAM {k}" for k in nonneg[:10]]))
    # the selected object keeps working: a new time step, coherencies of the right width
    sky.update_positions(time + 0.01, loc)
    out = sky.coherency_calc(freqs=np.array([150e6, 160e6]))
    assert out.shape[:3] == (2, 2, 2) and out.shape[3] == int(sky.above_horizon.sum())
    # unsorted indices with repeats are kept in the order given (numpy fancy indexing)
    sky4, *_ = _gleam_model(rsk, with_nans=False)
    pick = np.array([7, 3, 3, 49, 0, -1])
    ref = sky4.stokes.cpu().numpy()[..., pick]
    sky4.select(component_inds=pick)
    assert np.array_equal(sky4.stokes.cpu().numpy(), ref)
    with pytest.raises(IndexError):
        sky4.select(component_inds=[0, 99])


@pytest.mark.gpu
def test_select_none(rsk):
    """tests/test_skymodel.py:2288-2299."""
    sky, *_ = _gleam_model(rsk, with_nans=False)
    sky2 = sky.select(non_nan=None, component_inds=None, inplace=False)
    _assert_same_model(sky, sky2)
    sky2 = sky.select(non_nan="all", component_inds=None, inplace=False)
    _assert_same_model(sky, sky2)
    assert sky2.history == sky.history  # nothing removed, history untouched
    assert sky.select(non_nan=None, component_inds=None) is None
    _assert_same_model(sky, sky2)


def _twenty(rsk, spec_type, init_kwargs, frame="fk5"):
    """The 20-source model of tests/test_skymodel.py:2328-2350 (angles in radians)."""
    nsrcs, minflux, maxflux = 20, 0.5, 3.0
    ras = np.deg2rad(np.linspace(0, 360.0, nsrcs))
    decs = np.deg2rad(np.linspace(-90, 90, nsrcs))
    if spec_type in ("flat", "spectral_index"):
        stokes = np.zeros((4, 1, nsrcs))
        stokes[0, :, :] = np.linspace(minflux, maxflux, nsrcs)
    else:
        stokes = np.zeros((4, 2, nsrcs))
        stokes[0, 0, :] = np.linspace(minflux, maxflux / 2.0, nsrcs)
        stokes[0, 1, :] = np.linspace(minflux * 2.0, maxflux, nsrcs)
    ucomp = maxflux + 1.3
    stokes[2, :, :] = ucomp
    sky = rsk.SkyModel(name=[f"src{i}" for i in range(nsrcs)], ra=ras, dec=decs, frame=frame, stokes=stokes,
                       spectral_type=spec_type, **init_kwargs)
    return sky, ucomp


@pytest.mark.gpu
@pytest.mark.parametrize(
    "spec_type, init_kwargs, cut_kwargs, cut_type",
    [
        ("flat", {}, {}, "min"),
        ("flat", {"reference_frequency": np.ones(20) * 200e6}, {}, "max"),
        ("full", {"freq_array": np.array([1e8, 1.5e8])}, {}, "both"),
        ("subband", {"freq_array": np.array([1e8, 1.5e8])}, {"freq_range": np.array([0.9e8, 2e8])}, "both"),
        ("subband", {"freq_array": np.array([1e8, 1.5e8])}, {"freq_range": np.array([1.1e8, 2e8])}, "both"),
        ("flat", {"freq_array": np.array([1e8])}, {"freq_range": np.array([0.9e8, 2e8])}, "both"),
    ],
)
def test_select_flux(rsk, spec_type, init_kwargs, cut_kwargs, cut_type):
    """tests/test_skymodel.py:2302-2384."""
    sky, ucomp = _twenty(rsk, spec_type, init_kwargs)
    st0 = sky.stokes.cpu().numpy().copy()
    min_cut = 1.0 if cut_type in ("min", "both") else None
    max_cut = 2.3 if cut_type in ("max", "both") else None
    freq_range = cut_kwargs.get("freq_range")
    sky.select(min_brightness=min_cut, max_brightness=max_cut, brightness_freq_range=freq_range, non_nan=None)
    s0 = sky.stokes[0].cpu().numpy()
    if freq_range is not None and max_cut is not None and np.min(freq_range) > np.min(init_kwargs["freq_array"]):
        assert np.all(s0 <= max_cut)
    else:
        if min_cut is not None:
            assert np.all(s0 >= min_cut)
        if max_cut is not None:
            assert np.all(s0 <= max_cut)
    assert np.all(sky.stokes[2].cpu().numpy() == ucomp)
    want = SEL.select_indices(sky.lon.cpu().numpy() * 0, sky.lat.cpu().numpy() * 0, st0, spectral_type=spec_type,
                              freq_array=init_kwargs.get("freq_array"), min_brightness=min_cut, max_brightness=max_cut,
                              brightness_freq_range=freq_range)
    assert np.array_equal(sky.stokes.cpu().numpy(), st0[..., want])
    assert np.array_equal(sky.name, np.array([f"src{i}" for i in range(20)])[want])


@pytest.mark.gpu
def test_select_flux_cut_error(rsk):
    """tests/test_skymodel.py:2387-2485 (the unit-checking decorator of astropy is not mirrored)."""
    sky, _ = _twenty(rsk, "subband", {"freq_array": np.array([1e8, 1.5e8])}, frame="icrs")
    with pytest.raises(ValueError, match=r"If set, non_nan can only be set to one of: \['any', 'all'\]"):
        sky.select(min_brightness=1.0, max_brightness=2.3, non_nan=True)
    si, _ = _twenty(rsk, "spectral_index", {"reference_frequency": np.ones(20) * 200e6, "spectral_index": np.ones(20) * 0.8},
                    frame="icrs")
    with pytest.raises(NotImplementedError, match="Flux cuts with spectral index type objects is not supported yet."):
        si.select(min_brightness=1.0, max_brightness=2.3, non_nan="all")
    with pytest.raises(ValueError, match="brightness_freq_range must have 2 elements."):
        sky.select(min_brightness=1.0, max_brightness=2.3, brightness_freq_range=0.9e8, non_nan="all")
    with pytest.raises(ValueError, match="No object frequencies in specified range for flux cuts."):
        sky.select(min_brightness=1.0, max_brightness=2.3, brightness_freq_range=np.array([1.1e8, 1.4e8]), non_nan="all")
    assert sky.Ncomponents == 20


@pytest.mark.gpu
@pytest.mark.parametrize(
    "spec_type, init_kwargs",
    [
        ("flat", {}),
        ("flat", {"reference_frequency": np.ones(20) * 200e6}),
        ("full", {"freq_array": np.array([1e8, 1.5e8])}),
        ("subband", {"freq_array": np.array([1e8, 1.5e8])}),
        ("flat", {"freq_array": np.array([1e8])}),
    ],
)
def test_select_field(rsk, spec_type, init_kwargs):
    """tests/test_skymodel.py:2517-2599."""
    sky, _ = _twenty(rsk, spec_type, init_kwargs, frame="icrs")
    lon_range = np.deg2rad([90.0, 240.0])
    sky2 = sky.copy()
    sky2.select(lon_range=lon_range)
    ra = sky2.lon.cpu().numpy()
    assert ra.size > 0 and np.all(ra >= lon_range[0]) and np.all(ra <= lon_range[1])
    lat_range = np.deg2rad([-45.0, 45.0])
    sky2 = sky.copy()
    sky2.select(lat_range=lat_range)
    dec = sky2.lat.cpu().numpy()
    assert dec.size > 0 and np.all(dec >= lat_range[0]) and np.all(dec <= lat_range[1])
    sky2 = sky.copy()
    sky2.select(lon_range=lon_range, lat_range=lat_range)
    ra, dec = sky2.lon.cpu().numpy(), sky2.lat.cpu().numpy()
    assert np.all(ra >= lon_range[0]) and np.all(ra <= lon_range[1])
    assert np.all(dec >= lat_range[0]) and np.all(dec <= lat_range[1])
    # wrapping longitude
    lon_range = np.deg2rad([270.0, 90.0])
    sky2 = sky.copy()
    sky2.select(lon_range=lon_range)
    ra = sky2.lon.cpu().numpy()
    assert np.nonzero((ra > lon_range[1]) & (ra < lon_range[0]))[0].size == 0
    full = sky.lon.cpu().numpy()
    assert ra.size == np.count_nonzero((full >= lon_range[0]) | (full <= lon_range[1]))


@pytest.mark.gpu
def test_select_field_error(rsk):
    """tests/test_skymodel.py:2602-2651 (type errors need astropy's Latitude / Longitude classes: mimicked)."""
    sky, _ = _twenty(rsk, "flat", {"reference_frequency": np.ones(20) * 200e6})

    class Longitude(np.ndarray):
        pass

    class Latitude(np.ndarray):
        pass

    lon_range = np.deg2rad([90.0, 240.0]).view(Longitude)
    lat_range = np.deg2rad([-45.0, 45.0]).view(Latitude)
    with pytest.raises(TypeError, match="lat_range must be an astropy Latitude object."):
        sky.select(lat_range=lon_range)
    with pytest.raises(ValueError, match="lat_range must be 2 element range with the second component larger than the first."):
        sky.select(lat_range=lat_range[0:1])
    with pytest.raises(ValueError, match="lat_range must be 2 element range with the second component larger than the first."):
        sky.select(lat_range=np.array([lat_range[1], lat_range[0]]))
    with pytest.raises(TypeError, match="lon_range must be an astropy Longitude object."):
        sky.select(lon_range=lat_range)
    with pytest.raises(ValueError, match="lon_range must be 2 element range."):
        sky.select(lon_range=lon_range[0:1])
    sky.select(lat_range=lat_range)
    with pytest.raises(ValueError, match="Select would result in an empty object."):
        sky.select(lat_range=np.deg2rad([-90.0, -75.0]))


def _gmst06(jd_ut1, jd_tt):
    """Greenwich mean sidereal time, IAU 2006 (eraGmst06): what Time.sidereal_time("mean") is built on."""
    t = (jd_tt - E.DJ00) / E.DJC
    poly = 0.014506 + (4612.156534 + (1.3915817 + (-0.00000044 + (-0.000029956 - 0.0000000368 * t) * t) * t) * t) * t
    return E.anp(E.era00(jd_ut1, 0.0) + poly * E.DAS2R)


@pytest.mark.gpu
def test_circumpolar_nonrising(rsk, site):
    """tests/test_skymodel.py:2661-2740: cut_nonrising, the rise/set LST coarse cut against 100 position updates,
    circumpolar / never-rising sets against tan(lat) tan(dec)."""
    time, loc = site
    ntimes, nras, ndecs = 100, 20, 20
    nsrcs = nras * ndecs
    lon_deg = np.rad2deg(loc.lon)
    ra = np.linspace(lon_deg - 90, lon_deg + 90, nras)
    dec = np.linspace(-90, 90, ndecs)
    ra, dec = map(np.ndarray.flatten, np.meshgrid(ra, dec))
    ra, dec = np.deg2rad(np.mod(ra, 360.0)), np.deg2rad(dec)
    names = np.array([f"src{i}" for i in range(nsrcs)])
    st = np.zeros((4, 1, nsrcs)); st[0] = 1.0
    sky = rsk.SkyModel(name=names, ra=ra, dec=dec, frame="icrs", stokes=st, spectral_type="flat")
    sky2 = sky.cut_nonrising(loc.lat, inplace=False)
    sky3 = sky.copy()
    sky3.cut_nonrising(loc.lat)
    _assert_same_model(sky2, sky3)
    nonrising = np.array([sky.name[i] not in sky2.name for i in range(nsrcs)])
    sky2.calculate_rise_set_lsts(loc.lat)
    rise, sett = sky2._rise_lst.cpu().numpy(), sky2._set_lst.cpu().numpy()
    r_ref, s_ref = SEL.rise_set_lsts(sky2.lon.cpu().numpy(), sky2.lat.cpu().numpy(), loc.lat)
    assert np.allclose(rise, r_ref, rtol=0, atol=1e-13, equal_nan=True)
    assert np.allclose(sett, s_ref, rtol=0, atol=1e-13, equal_nan=True)
    is_below = np.zeros(nsrcs, dtype=bool)
    is_below[nonrising] = True
    alts = []
    for ti in range(ntimes):
        t = time + ti / (ntimes - 1.0)
        sky.update_positions(t, loc)
        alt = sky.alt_az[0].cpu().numpy()
        alts.append(alt)
        jd_utc = t.jd1 + t.jd2
        lst = E.anp(_gmst06(jd_utc, jd_utc + 67.184 / 86400.0) + loc.lon)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            dt0 = lst - rise
            dt1 = sett - rise
            dt0[dt0 < 0] += 2 * np.pi
            dt1[dt1 < 0] += 2 * np.pi
            is_below[~nonrising] = dt0 > dt1
        assert np.all(alt[is_below] < 0.0)
    alts = np.array(alts)
    nonrising_test = np.where(np.all(alts < 0, axis=0))[0]
    circumpolar_test = np.where(np.all(alts > 0, axis=0))[0]
    tans = np.tan(loc.lat) * np.tan(dec)
    assert np.array_equal(np.where(tans > 1)[0], circumpolar_test)
    assert np.array_equal(np.where(tans < -1)[0], nonrising_test)
    assert np.array_equal(np.where(nonrising)[0], nonrising_test)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["point", "healpix_ring", "healpix_nested_subset"])
def test_select_matches_oracle_random(rsk, kind):
    """Random catalogues with NaNs and negative fluxes, random criteria: same index list as the oracle, and every
    array is the numpy fancy index of the original."""
    import torch

    rng = np.random.default_rng(11)
    nf = 7
    fa = np.sort(rng.uniform(100e6, 200e6, nf))
    if kind == "point":
        n = 20011
        lon, lat = rng.uniform(0, 2 * np.pi, n), np.arcsin(rng.uniform(-1, 1, n))
        coords = dict(ra=lon, dec=lat, frame="icrs")
    else:
        nside = 32
        nested = kind.endswith("subset")
        pix = np.sort(rng.choice(12 * nside * nside, 9000, replace=False)) if nested else np.arange(12 * nside * nside)
        n = pix.size
        lon, lat = hp.healpix_to_lonlat(nside, pix, nested=nested)
        coords = dict(nside=nside, hpx_order="nested" if nested else "ring", hpx_frame="icrs",
                      hpx_inds=pix if nested else None)
    st = rng.normal(1.0, 1.0, (4, nf, n))
    st[0][rng.random((nf, n)) < 0.02] = np.nan
    st[2][rng.random((nf, n)) < 0.01] = np.nan
    st[:, :, rng.choice(n, 50, replace=False)] = np.nan
    err = rng.random((4, nf, n))
    cases = [
        dict(non_nan="any"),
        dict(non_nan="all", non_negative=True),
        dict(lat_range=np.array([-0.3, 0.9])),
        dict(lon_range=np.array([1.0, 4.0]), min_brightness=0.1),
        dict(lon_range=np.array([5.5, 0.7]), max_brightness=2.5, brightness_freq_range=np.array([fa[1], fa[4]])),
        dict(component_inds=rng.integers(0, n, 5000), lat_range=np.array([-1.0, 0.2])),
        dict(component_inds=rng.integers(0, n, 5000), lon_range=np.array([6.0, 0.5]), non_negative=True),
        dict(component_inds=rng.permutation(n)[: n // 2], min_brightness=-0.5, max_brightness=3.0, non_nan="all",
             lat_range=np.array([-1.2, 1.2]), lon_range=np.array([0.2, 6.0])),
    ]
    for kw in cases:
        sky = rsk.SkyModel(stokes=st, spectral_type="full", freq_array=fa, stokes_error=err, **coords)
        want = SEL.select_indices(lon, lat, st, spectral_type="full", freq_array=fa, **kw)
        sky.select(**kw)
        assert sky.Ncomponents == want.size, kw
        assert np.array_equal(sky.stokes.cpu().numpy(), st[..., want], equal_nan=True), kw
        assert np.array_equal(sky.stokes_error.cpu().numpy(), err[..., want]), kw
        glon, glat = (x.cpu().numpy() for x in sky.get_lon_lat())
        assert np.abs(glon - lon[want]).max() < 1e-14 and np.abs(glat - lat[want]).max() < 1e-14, kw
        if kind != "point":
            assert torch.equal(sky.hpx_inds.cpu(), torch.as_tensor(pix[want])), kw


@pytest.mark.gpu
def test_select_large_and_abi_errors(rsk):
    """A million components through the three kernels, checked against torch boolean indexing; argument errors."""
    import torch
    from pyradiosky_b200 import _lib

    rng = np.random.default_rng(12)
    n, nf = 1_000_003, 3
    lon, lat = rng.uniform(0, 2 * np.pi, n), np.arcsin(rng.uniform(-1, 1, n))
    st = np.abs(rng.normal(1.0, 1.0, (4, nf, n)))
    sky = rsk.SkyModel(ra=lon, dec=lat, stokes=st, spectral_type="full", freq_array=np.array([1e8, 1.1e8, 1.2e8]))
    keep = (sky.lat >= -0.5) & (sky.lat <= 0.7) & (sky.stokes[0].min(dim=0).values >= 0.3)
    ref = sky.stokes[..., keep].clone()
    sky.select(lat_range=np.array([-0.5, 0.7]), min_brightness=0.3)
    assert torch.equal(sky.stokes, ref) and sky.Ncomponents == int(keep.sum())
    lib = _lib.load()
    sel = _lib.RskSelect()
    sel.non_nan = 5
    buf = torch.zeros(16, dtype=torch.uint8, device=sky.device)
    assert lib.rsk_select_mask(None, _lib.ptr(sky.stokes), 1, 16, None, 16, ctypes.byref(sel), _lib.ptr(buf), None) == -3
    sel.non_nan = 1
    assert lib.rsk_select_mask(None, None, 1, 16, None, 16, ctypes.byref(sel), _lib.ptr(buf), None) == -1
    assert lib.rsk_select_mask(None, None, 1, 16, None, 16, None, _lib.ptr(buf), None) == -1
    sel.non_nan, sel.use_lat = 0, 1
    assert lib.rsk_select_mask(None, None, 1, 16, None, 16, ctypes.byref(sel), _lib.ptr(buf), None) == -1  # no coords
    assert lib.rsk_select_mask(None, None, 1, 16, None, -1, ctypes.byref(sel), _lib.ptr(buf), None) == -2
    idx = torch.zeros(4, dtype=torch.int64, device=sky.device)
    assert lib.rsk_gather_rows(_lib.ptr(sky.stokes), 1, 16, _lib.ptr(idx), 4, 4, _lib.ptr(buf), None) == -3
    assert lib.rsk_gather_rows(None, 1, 16, _lib.ptr(idx), 4, 8, _lib.ptr(buf), None) == -1
    assert lib.rsk_gather_rows(None, 0, 16, None, 4, 8, None, None) == 0
    assert lib.rsk_rise_set_lsts(None, 4, 0.1, 0.0, None, None, None) == -1


# ---------------------------------------------------------------------------------------------------
# concat (skymodel.py:2714-3010): select into thirds, concatenate, get the original back
# ---------------------------------------------------------------------------------------------------
GSM = os.path.join(ROOT, "tests", "golden", "ref_fixture_gsm_icrs_nside8.skyh5")


def _full_model(rsk, comp_type, spec_type="subband"):
    import torch

    if comp_type == "point":
        sky, lon, lat, st, fa = _gleam_model(rsk, with_nans=False)
        if spec_type == "flat":
            sky = rsk.SkyModel(name=sky.name, ra=lon, dec=lat, frame="fk5", stokes=st[:, 9:10, :], spectral_type="flat",
                               history="read from the fixture.")
        elif spec_type == "spectral_index":
            sky = rsk.SkyModel(name=sky.name, ra=lon, dec=lat, frame="fk5", stokes=st[:, 9:10, :],
                               spectral_type="spectral_index", reference_frequency=np.full(50, 200e6),
                               spectral_index=np.linspace(-1.2, 0.2, 50), history="read from the fixture.")
        sky.filename = ["gleam_50srcs.vot"]
        sky.select(non_negative=True, non_nan="all")
    else:
        sky = rsk.SkyModel.from_skyh5(GSM)
        sky.filename = ["healpix_disk.skyh5"]
    return sky


@pytest.mark.gpu
@pytest.mark.parametrize("comp_type, spec_type", [("point", "subband"), ("point", "spectral_index"), ("point", "flat"),
                                                  ("healpix", "full")])
def test_concat(rsk, comp_type, spec_type):
    """tests/test_skymodel.py:1629-1722."""
    import torch

    full = _full_model(rsk, comp_type, spec_type)
    n = full.Ncomponents
    full.extended_model_group = np.array([f"g{k}" for k in range(n)]) if full.name is None else full.name.copy()
    full.beam_amp = torch.ones((4, full.Nfreqs, n), dtype=torch.float64, device=full.device)
    full.calc_frame_coherency()
    full.extra_columns = {"foo": np.arange(n, dtype=float), "bar": np.arange(n, dtype=int),
                          "gah": np.array(["gah " + str(i) for i in range(n)])}
    s1 = full.select(component_inds=np.arange(n // 3), inplace=False)
    s2 = full.select(component_inds=np.arange(n // 3, 2 * n // 3), inplace=False)
    s3 = full.select(component_inds=np.arange(2 * n // 3, n), inplace=False)
    new = s1.concat(s2, inplace=False)
    assert new.Ncomponents == s1.Ncomponents + s2.Ncomponents and s1.Ncomponents == n // 3
    new.concat(s3)
    assert new.filename == full.filename
    expected = (full.history + "  Downselected to specific components using pyradiosky."
                + " Combined skymodels along the component axis using pyradiosky."
                + " Combined skymodels along the component axis using pyradiosky.")
    assert new.history != full.history and rsk.SkyModel._histories_match(new.history, expected)
    _assert_same_model(new, full)
    assert torch.equal(new.frame_coherency, full.frame_coherency)
    assert np.array_equal(new.extended_model_group, full.extended_model_group)
    assert torch.equal(new._polarized_mask, full._polarized_mask) and new._n_polarized == full._n_polarized
    # the combined object works on the hot path exactly like the original
    loc = rsk.EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
    t = rsk.Time.from_utc("2015-03-01 00:00:00")
    freqs = np.array([110e6, 150e6]) if spec_type != "full" else None
    outs = []
    for obj in (new, full):
        obj.update_positions(t, loc)
        outs.append(obj.coherency_calc(freqs=freqs) if freqs is not None else obj.coherency_calc())
    assert torch.equal(outs[0], outs[1])
    # history handling
    s2.history += " testing the history."
    new = s1.concat(s2, inplace=False, run_check=False)
    new.concat(s3)
    expected = (full.history + "  Downselected to specific components using pyradiosky."
                + " Combined skymodels along the component axis using pyradiosky. "
                + "Unique part of next object history follows.  testing history."
                + " Combined skymodels along the component axis using pyradiosky.")
    assert rsk.SkyModel._histories_match(new.history, expected), new.history
    new = s1.concat(s2, inplace=False, verbose_history=True)
    new.concat(s3, verbose_history=True)
    expected = (full.history + "  Downselected to specific components using pyradiosky."
                + " Combined skymodels along the component axis using pyradiosky. "
                + "Next object history follows. " + s2.history
                + " Combined skymodels along the component axis using pyradiosky. "
                + "Next object history follows. " + s3.history)
    assert rsk.SkyModel._histories_match(new.history, expected)


@pytest.mark.gpu
@pytest.mark.parametrize("param", ["reference_frequency", "extended_model_group", "beam_amp", "stokes_error", "name",
                                   "spectral_index", "frame_coherency"])
def test_concat_optional_params(rsk, param):
    """tests/test_skymodel.py:1725-1896: a parameter only one object has is filled with NaNs / empty strings (or dropped,
    for name and frame_coherency), with the reference's warning."""
    import torch

    full = _full_model(rsk, "healpix" if param == "name" else "point", "flat")
    n = full.Ncomponents
    dev = full.device
    if param == "name":
        full.name = np.array(["hpx" + str(i) for i in range(n)])
    elif param == "extended_model_group":
        full.extended_model_group = full.name.copy()
    elif param == "beam_amp":
        full.beam_amp = torch.ones((4, full.Nfreqs, n), dtype=torch.float64, device=dev)
    elif param == "stokes_error":
        full.stokes_error = full.stokes * 0.05
    elif param == "spectral_index":
        full.spectral_index = torch.full((n,), -0.7, dtype=torch.float64, device=dev)
    elif param == "reference_frequency":
        full.reference_frequency = torch.full((n,), 2e8, dtype=torch.float64, device=dev)
    elif param == "frame_coherency":
        full.calc_frame_coherency()
    assert getattr(full, param) is not None
    base = full.filename[0]
    for drop_first in (True, False):
        s1 = full.select(component_inds=np.arange(n // 2), inplace=False)
        s1.filename = [base + "_1"]
        s2 = full.select(component_inds=np.arange(n // 2, n), inplace=False)
        s2.filename = [base + "_2"]
        new = s1.concat(s2, inplace=False)
        _assert_same_model(new, full)
        setattr(s1 if drop_first else s2, param, None)
        with pytest.warns(UserWarning, match=f"Only one object has {param}"):
            new = s1.concat(s2, inplace=False)
        assert new.filename == [base + "_1", base + "_2"]
        got = getattr(new, param)
        if param in ("name", "frame_coherency"):
            assert got is None
            continue
        assert got is not None
        want = getattr(full, param)
        kept = slice(n // 2, n) if drop_first else slice(0, n // 2)
        lost = slice(0, n // 2) if drop_first else slice(n // 2, n)
        if param == "extended_model_group":
            assert got[kept].tolist() == want[kept].tolist() and np.all(got[lost] == "")
        else:
            assert torch.equal(got[..., kept], want[..., kept]) and bool(torch.isnan(got[..., lost]).all())
        setattr(new, param, want)
        _assert_same_model(new, full)


@pytest.mark.gpu
def test_concat_errors(rsk, site):
    """tests/test_skymodel.py:1899-2060."""
    time, loc = site
    for comp_type, msg in (("point", "The two SkyModel objects contain components with the same name."),
                           ("healpix", "The two SkyModel objects contain overlapping Healpix pixels.")):
        full = _full_model(rsk, comp_type)
        n = full.Ncomponents
        s1 = full.select(component_inds=np.arange(n // 2), inplace=False)
        s2 = full.select(component_inds=np.arange(n // 3, n), inplace=False)
        with pytest.raises(ValueError, match=msg):
            s1.concat(s2)
    sub = _full_model(rsk, "point", "subband")
    spi = _full_model(rsk, "point", "spectral_index")
    hpx = _full_model(rsk, "healpix")
    with pytest.raises(ValueError, match="Only SkyModel"):
        sub.concat("foo")
    with pytest.raises(ValueError, match="UVParameter component_type does not match. "):
        sub.concat(hpx)
    with pytest.raises(ValueError, match="UVParameter spectral_type does not match. "):
        sub.concat(spi)
    n = sub.Ncomponents
    s1 = sub.select(component_inds=np.arange(n // 2), inplace=False)
    s2 = sub.select(component_inds=np.arange(n // 2, n), inplace=False)
    s2.freq_array = s2.freq_array * 2.0
    with pytest.raises(ValueError, match="UVParameter freq_array does not match. "):
        s1.concat(s2)
    m = hpx.Ncomponents
    h1 = hpx.select(component_inds=np.arange(m // 2), inplace=False)
    h2 = hpx.select(component_inds=np.arange(m // 2, m), inplace=False)
    h2.nside = h2.nside * 2
    with pytest.raises(ValueError, match="UVParameter nside does not match. "):
        h1.concat(h2)
    hpx.update_positions(time, loc)
    h1 = hpx.select(component_inds=np.arange(m // 2), inplace=False)
    h2 = hpx.select(component_inds=np.arange(m // 2, m), inplace=False)
    h2.update_positions(time + 1.0, loc)
    with pytest.raises(ValueError, match="UVParameter time does not match. "):
        h1.concat(h2, clear_time_position=False)
    h1.concat(h2, clear_time_position=True)
    hpx.clear_time_position_specific_params()
    _assert_same_model(h1, hpx)
    s1 = sub.select(component_inds=np.arange(n // 2), inplace=False)
    s2 = sub.select(component_inds=np.arange(n // 2, n), inplace=False)
    s1.extra_columns = {"foo": np.arange(s1.Ncomponents, dtype=float), "bar": np.arange(s1.Ncomponents, dtype=int)}
    with pytest.raises(ValueError, match="One object has extra_columns and the other does not. Cannot combine objects."):
        s1.concat(s2)
    s2.extra_columns = {"foo": np.arange(s2.Ncomponents, dtype=float), "blech": np.arange(s2.Ncomponents, dtype=int)}
    with pytest.raises(ValueError, match="Both objects have extra_columns but the column names do not match."):
        s1.concat(s2)
    s2.extra_columns = {"foo": np.arange(s2.Ncomponents, dtype=float), "bar": np.arange(s2.Ncomponents, dtype=float)}
    with pytest.raises(ValueError, match="Both objects have extra_columns but the dtypes for column bar do not match."):
        s1.concat(s2)
    # positions kept when both halves were updated for the same time
    s1 = sub.select(component_inds=np.arange(n // 2), inplace=False)
    s2 = sub.select(component_inds=np.arange(n // 2, n), inplace=False)
    for obj in (sub, s1, s2):
        obj.update_positions(time, loc)
    with pytest.raises(ValueError, match="UVParameter alt_az does not match. "):
        s1.concat(s2, clear_time_position=False)  # (arrays of different halves never compare equal, as upstream)
    lib_rc = __import__("pyradiosky_b200")._lib.load().rsk_concat_rows(None, 1, None, 1, 1, 3, None, None)
    assert lib_rc == -3


# ---------------------------------------------------------------------------------------------------
# transform_to (skymodel.py:1419-1446)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_transform_to(rsk, site):
    import torch

    rng = np.random.default_rng(21)
    n = 50021
    lon, lat = rng.uniform(0, 2 * np.pi, n), np.arcsin(rng.uniform(-1, 1, n))
    lon[:4], lat[:4] = [0.0, np.pi / 2, np.pi, 3 * np.pi / 2], [0.0, 0.0, np.pi / 2 - 1e-9, -np.pi / 2 + 1e-9]
    st = np.zeros((4, 1, n)); st[0] = 1.0; st[1] = 0.1
    sky = rsk.SkyModel(ra=lon, dec=lat, frame="icrs", stokes=st, spectral_type="flat")
    time, loc = site
    sky.update_positions(time, loc)
    alt_az0 = sky.alt_az.clone()
    coh0 = sky.coherency_calc()
    for frame in ("galactic", "fk5", "icrs"):
        old = sky.frame
        prev = np.stack([x.cpu().numpy() for x in (sky.lon, sky.lat)])
        sky.transform_to(frame)
        assert sky.frame == frame
        m = E.frame_to_icrs(frame).T @ E.frame_to_icrs(old)
        v = m @ E.s2c(prev[0], prev[1])
        want_lon, want_lat = np.mod(np.arctan2(v[1], v[0]), 2 * np.pi), np.arctan2(v[2], np.hypot(v[0], v[1]))
        got_lon, got_lat = sky.lon.cpu().numpy(), sky.lat.cpu().numpy()
        dlon = np.abs(np.mod(got_lon - want_lon + np.pi, 2 * np.pi) - np.pi) * np.cos(want_lat)
        assert dlon.max() < 1e-15 and np.abs(got_lat - want_lat).max() < 1e-15
        assert got_lon.min() >= 0 and got_lon.max() <= 2 * np.pi
        # the sky has not moved: same Alt/Az from the new coordinates (the frame matrix goes with them)
        sky.clear_time_position_specific_params()
        sky.update_positions(time, loc)
        d = (sky.alt_az - alt_az0).abs()
        d[1] = torch.minimum(d[1], 2 * np.pi - d[1]) * torch.cos(alt_az0[0])
        assert float(d.max()) < 1e-14
    # back in ICRS after the round trip
    dl = np.abs(np.mod(sky.lon.cpu().numpy() - lon + np.pi, 2 * np.pi) - np.pi) * np.cos(lat)
    assert dl.max() < 5e-15 and np.abs(sky.lat.cpu().numpy() - lat).max() < 5e-15  # (three rotations)
    assert coh0.shape == sky.coherency_calc().shape
    # known answers: the galactic centre and the north galactic pole in ICRS (IAU 1958 definition via FK5 J2000)
    gal = rsk.SkyModel(ra=np.array([0.0, 0.0]), dec=np.array([0.0, np.pi / 2]), frame="galactic",
                       stokes=np.ones((4, 1, 2)), spectral_type="flat")
    gal.transform_to("icrs")
    ra, dec = np.rad2deg(gal.lon.cpu().numpy()), np.rad2deg(gal.lat.cpu().numpy())
    assert abs(ra[0] - 266.40499) < 2e-4 and abs(dec[0] + 28.93617) < 2e-4
    assert abs(ra[1] - 192.85948) < 2e-4 and abs(dec[1] - 27.12825) < 2e-4
    hpx = rsk.SkyModel(nside=2, stokes=np.ones((4, 1, 48)), spectral_type="flat", hpx_frame="icrs")
    with pytest.raises(ValueError, match="Direct coordinate transformation between frames is not valid"):
        hpx.transform_to("galactic")
    with pytest.raises(ValueError, match="not supported"):
        sky.transform_to("fk4")


@pytest.mark.gpu
def test_equality_operator(rsk, site):
    """``==`` as the reference's tests use it (UVBase.__eq__): copies, select / concat round trips, differences."""
    time, loc = site
    full = _full_model(rsk, "point", "subband")
    n = full.Ncomponents
    other = full.copy()
    assert full == other and not (full != other)
    s1 = full.select(component_inds=np.arange(n // 2), inplace=False)
    s2 = full.select(component_inds=np.arange(n // 2, n), inplace=False)
    assert s1 != full and s1 != s2
    new = s1.concat(s2, inplace=False)
    assert new != full  # histories differ
    new.history = full.history
    assert new == full
    new.stokes[0, 0, 0] *= 1.001
    assert new != full
    other.update_positions(time, loc)
    assert other != full
    full.update_positions(time, loc)
    assert other == full
    other.name = other.name.copy(); other.name[0] = "different"
    assert other != full and full != "foo"
    hp_a = rsk.SkyModel(nside=4, stokes=np.ones((4, 1, 192)), spectral_type="flat", hpx_frame="icrs")
    hp_b = rsk.SkyModel(nside=4, stokes=np.ones((4, 1, 192)), spectral_type="flat", hpx_frame="icrs", hpx_inds=np.arange(192))
    assert hp_a == hp_b
    hp_c = rsk.SkyModel(nside=4, stokes=np.ones((4, 1, 192)), spectral_type="flat", hpx_frame="icrs", hpx_order="nested")
    assert hp_a != hp_c


# ---------------------------------------------------------------------------------------------------
# healpix_interp_transform (skymodel.py:1448-1538): bilinear HEALPix interpolation into a new frame
# ---------------------------------------------------------------------------------------------------
GSM_GAL = os.path.join(ROOT, "tests", "golden", "ref_fixture_gsm_galactic_nside8.skyh5")


def test_oracle_interpolation_reproduces_reference_golden_pair():
    """tests/test_skymodel.py:4423-4452: gsm_galactic.skyh5 transformed to ICRS must equal gsm_icrs.skyh5 (UVParameter
    tolerance rtol 1e-5).  This pins oracle/healpix_np.interp_weights and the galactic -> ICRS rotation."""
    g, i = h5_np.File(GSM_GAL), h5_np.File(GSM)
    sg, si = np.array(g["/Data/stokes"].read(), float), np.array(i["/Data/stokes"].read(), float)
    hg, hi = np.array(g["/Header/hpx_inds"].read()), np.array(i["/Header/hpx_inds"].read())
    assert bytes(g["/Header/frame"].read()) == b"galactic" and bytes(i["/Header/frame"].read()) == b"icrs"
    rot = E.frame_to_icrs("galactic").T  # icrs (new) -> galactic (old)
    worst = 0.0
    for f in range(sg.shape[1]):
        inds, vals = hp.interp_transform(8, False, hg, sg[0, f], rot)
        assert np.array_equal(inds, hi)
        worst = max(worst, (np.abs(vals - si[0, f]) / np.abs(si[0, f])).max())
    assert worst < 2e-6  # observed 9.7e-7; the reference's own comparison allows 1e-5


def test_oracle_interpolation_weights_properties():
    rng = np.random.default_rng(41)
    for nside in (1, 2, 8, 64):
        npix = 12 * nside * nside
        lon = np.concatenate([rng.uniform(0, 2 * np.pi, 4000), [0.0, 1e-12, 2 * np.pi - 1e-12, 3.0, 3.0]])
        lat = np.concatenate([np.arcsin(rng.uniform(-1, 1, 4000)), [0.0, 0.0, 0.0, np.pi / 2 - 1e-9, -np.pi / 2 + 1e-9]])
        pix, wgt = hp.interp_weights(nside, lon, lat)
        assert pix.min() >= 0 and pix.max() < npix
        assert np.abs(wgt.sum(axis=0) - 1).max() < 1e-12 and wgt.min() > -1e-12 and wgt.max() < 1 + 1e-12
        # a constant map is reproduced, and so is the map value at a pixel centre
        assert np.abs((wgt * np.ones(npix)[pix]).sum(axis=0) - 1).max() < 1e-12
        clon, clat = hp.healpix_to_lonlat(nside, np.arange(npix))
        cp, cw = hp.interp_weights(nside, clon, clat)
        vals = rng.normal(size=npix)
        assert np.abs((cw * vals[cp]).sum(axis=0) - vals).max() < 1e-9
        # nested numbering names the same four pixels
        pn, wn = hp.interp_weights(nside, lon, lat, nested=True) if nside & (nside - 1) == 0 else (None, None)
        if pn is not None:
            assert np.array_equal(pn, hp.ring2nest(nside, pix.ravel()).reshape(pix.shape)) and np.array_equal(wn, wgt)
        # continuity: a small step moves the interpolated value of a smooth map by a small amount
        zmap = np.sin(clat) + 0.3 * np.cos(clon) * np.cos(clat)
        v0 = (wgt * zmap[pix]).sum(axis=0)
        p1, w1 = hp.interp_weights(nside, lon + 1e-7, lat)
        assert np.abs((w1 * zmap[p1]).sum(axis=0) - v0).max() < 1e-5


@pytest.mark.gpu
def test_healpix_interp_transform_golden_pair(rsk):
    """tests/test_skymodel.py:4423-4452 on the GPU: galactic GSM map -> ICRS equals the reference's ICRS file."""
    import torch

    gal = rsk.SkyModel.from_skyh5(GSM_GAL)
    icrs = rsk.SkyModel.from_skyh5(GSM)
    assert gal.frame == "galactic" and icrs.frame == "icrs" and gal.Ncomponents == 768
    gal.calc_frame_coherency()
    before = gal.copy()
    new = gal.healpix_interp_transform("icrs", inplace=False)
    assert gal == before and new != gal
    assert new.frame == "icrs" and new.Ncomponents == 768 and new.frame_coherency is not None
    rel = ((new.stokes[0] - icrs.stokes[0]).abs() / icrs.stokes[0].abs()).max()
    assert float(rel) < 2e-6
    new.history, new.filename = icrs.history, icrs.filename
    new.frame_coherency = None
    assert new == icrs  # the reference's assertion, to its own (rtol 1e-5) tolerance
    gal.healpix_interp_transform("icrs")
    assert torch.equal(gal.stokes, new.stokes)
    # the oracle, value for value
    g = h5_np.File(GSM_GAL)
    sg = np.array(g["/Data/stokes"].read(), float)
    rot = E.frame_to_icrs("galactic").T
    for f in (0, 9):
        _, vals = hp.interp_transform(8, False, np.arange(768), sg[0, f], rot)
        assert np.abs(new.stokes[0, f].cpu().numpy() - vals).max() < 1e-9 * np.abs(vals).max()


@pytest.mark.gpu
@pytest.mark.parametrize("order", ["ring", "nested"])
def test_healpix_interp_transform_partial_sky(rsk, order):
    """Partial maps: pixels whose four neighbours are not all present drop out; full_sky keeps them as zeros
    (tests/test_skymodel.py:4488-4503); ring and nested; errors (:4455-4485)."""
    rng = np.random.default_rng(43)
    nside, nf = 16, 3
    npix = 12 * nside * nside
    nested = order == "nested"
    lon, lat = hp.healpix_to_lonlat(nside, np.arange(npix), nested=nested)
    have = np.nonzero(lat > -0.3)[0]                      # a cap-and-belt map
    have = have[rng.permutation(have.size)]               # in scrambled order
    vals = 5.0 + np.sin(lat) + 0.3 * np.cos(lon) * np.cos(lat)
    st = np.zeros((4, nf, have.size))
    st[0] = vals[have][None, :] * np.array([1.0, 2.0, 3.0])[:, None]
    sky = rsk.SkyModel(nside=nside, hpx_order=order, hpx_frame="icrs", hpx_inds=have, stokes=st, spectral_type="full",
                       freq_array=np.array([1e8, 1.1e8, 1.2e8]), stokes_unit="K")
    rot = E.frame_to_icrs("icrs").T @ E.frame_to_icrs("galactic")  # galactic (new) -> icrs (old)
    want_inds, want0 = hp.interp_transform(nside, nested, have, st[0, 0], rot)
    new = sky.healpix_interp_transform("galactic", inplace=False)
    assert 0 < new.Ncomponents < npix
    assert np.array_equal(new.hpx_inds.cpu().numpy(), want_inds)
    got = new.stokes.cpu().numpy()
    assert np.abs(got[0, 0] - want0).max() < 1e-12 and np.abs(got[0, 2] - 3 * want0).max() < 1e-11
    assert not got[1:].any()
    # a smooth map is interpolated to within the pixel scale of its true value at the new centres
    glon, glat = (x.cpu().numpy() for x in new.get_lon_lat())
    v = E.frame_to_icrs("galactic") @ E.s2c(glon, glat)
    truth = 5.0 + v[2] + 0.3 * v[0]
    assert np.abs(got[0, 0] - truth).max() < 5e-3
    full = sky.healpix_interp_transform("galactic", full_sky=True, inplace=False)
    assert full.Ncomponents == npix and full.hpx_inds is None
    fs = full.stokes.cpu().numpy()[0, 0]
    assert np.array_equal(np.nonzero(fs)[0], want_inds) and np.abs(fs[want_inds] - want0).max() < 1e-12
    # errors
    pt = rsk.SkyModel(ra=[0.1], dec=[0.2], stokes=np.ones((4, 1, 1)), spectral_type="flat")
    with pytest.raises(ValueError, match="Healpix frame interpolation is not valid for point source catalogs."):
        pt.healpix_interp_transform("galactic")
    sky.stokes[1] = sky.stokes[0]
    with pytest.raises(NotImplementedError, match="Healpix map transformations are currently not implemented for"):
        sky.healpix_interp_transform("galactic")
    from pyradiosky_b200 import _lib

    assert _lib.load().rsk_healpix_interp(0, 0, None, None, 0, 0, None, 0, None, None, None) == -7
    assert _lib.load().rsk_healpix_interp(8, 0, None, None, 0, 0, None, 0, None, None, None) == -1


@pytest.mark.gpu
def test_extra_columns(rsk):
    """skymodel.py:1279-1376 (tests/test_skymodel.py add / remove extra columns): they follow select and concat."""
    sky = _full_model(rsk, "point", "flat")
    n = sky.Ncomponents
    sky.add_extra_columns(names="foo", values=np.arange(n, dtype=float))
    sky.add_extra_columns(names=["bar", "gah"], values=[np.arange(n), np.array(["g" + str(k) for k in range(n)])],
                          dtype=[int, str])
    assert list(sky.extra_columns) == ["foo", "bar", "gah"] and sky.extra_columns["bar"].dtype.kind == "i"
    with pytest.raises(ValueError, match="foo column already exists, use a different name."):
        sky.add_extra_columns(names="foo", values=np.zeros(n))
    with pytest.raises(ValueError, match="Must provide the same number of names and values."):
        sky.add_extra_columns(names=["a", "b"], values=[np.zeros(n)])
    with pytest.raises(ValueError, match="If dtype is set, it must be the same length as `name`."):
        sky.add_extra_columns(names=["a"], values=[np.zeros(n)], dtype=[float, int])
    with pytest.raises(ValueError, match="The value array in index 0 is not the right shape."):
        sky.add_extra_columns(names="a", values=np.zeros(n + 1))
    half = sky.select(component_inds=np.arange(0, n, 2), inplace=False)
    assert half.extra_columns["gah"].tolist() == ["g" + str(k) for k in range(0, n, 2)]
    rest = sky.select(component_inds=np.arange(1, n, 2), inplace=False)
    both = half.concat(rest, inplace=False)
    assert both.extra_columns["foo"].tolist() == list(range(0, n, 2)) + list(range(1, n, 2))
    sky.remove_extra_columns(["bar", "gah"])
    assert list(sky.extra_columns) == ["foo"]
    sky.remove_extra_columns("foo")
    assert sky.extra_columns is None


@pytest.mark.gpu
def test_shard_sky_and_reassemble(rsk):
    """dist.shard_sky: the per-rank blocks of a loaded catalogue, concatenated in rank order, are the catalogue."""
    from pyradiosky_b200 import dist

    for comp_type in ("point", "healpix"):
        full = _full_model(rsk, comp_type)
        world = 3
        parts = [dist.shard_sky(full, world, r) for r in range(world)]
        assert sum(p.Ncomponents for p in parts) == full.Ncomponents
        assert [p.Ncomponents for p in parts] == [b - a for a, b in (dist.shard_bounds(full.Ncomponents, world, r) for r in range(world))]
        whole = parts[0]
        for p in parts[1:]:
            whole = whole.concat(p, inplace=False)
        whole.history = full.history
        assert whole == full
    one = dist.shard_sky(full, 1, 0)  # a single rank keeps everything (select removes nothing)
    assert one == full and one is not full
    with pytest.raises(ValueError, match="Select would result in an empty object."):
        dist.shard_sky(full.select(component_inds=[0, 1], inplace=False), 4, 3)


@pytest.mark.gpu
def test_check_acceptability_warnings(rsk):
    """skymodel.py:1184-1277; the warnings tests/test_skymodel.py:2174-2223 expects around the NaN / negative cuts."""
    sky, *_ = _gleam_model(rsk)  # NaNs injected, GLEAM has negative fluxes
    with pytest.warns(UserWarning) as rec:
        assert sky.check()
    msgs = [str(w.message) for w in rec]
    assert any("Some Stokes values are NaNs." in m for m in msgs) and any("Some Stokes I values are negative" in m for m in msgs)
    clean = sky.select(non_nan="any", non_negative=True, inplace=False)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert clean.check() and sky.check(run_check_acceptability=False)
    only_nan = sky.select(non_negative=True, inplace=False)
    with pytest.warns(UserWarning, match="Some Stokes values are NaNs.") as rec:
        only_nan.check()
    assert not any("negative" in str(w.message) for w in rec)
    spi = _full_model(rsk, "point", "spectral_index")
    spi.spectral_index[3] = float("nan")
    with pytest.warns(UserWarning, match="Some spectral index values are NaNs."):
        spi.check()
    spi.freq_array = np.array([1e8])
    with pytest.raises(ValueError, match="Only one of freq_array and reference_frequency can be specified, not both."):
        spi.check()
    hpx = rsk.SkyModel(nside=2, stokes=np.ones((4, 1, 48)), spectral_type="flat", hpx_frame="icrs", stokes_unit="Jy")
    with pytest.raises(ValueError, match="For healpix component types, the stokes parameter must have a unit"):
        hpx.check()


@pytest.mark.gpu
def test_frame_coordinate_attributes(rsk):
    """skymodel.py:1065-1090 / tests/test_skymodel.py:2575-2583: ``ra`` / ``dec`` or ``l`` / ``b`` follow the frame."""
    import torch

    sky, _ = _twenty(rsk, "flat", {}, frame="icrs")
    assert torch.equal(sky.ra, sky.lon) and torch.equal(sky.dec, sky.lat)
    with pytest.raises(AttributeError, match="'SkyModel' object has no attribute 'b'"):
        sky.b  # noqa: B018
    sky.transform_to("galactic")
    assert torch.equal(sky.l, sky.lon) and torch.equal(sky.b, sky.lat)
    with pytest.raises(AttributeError, match="'SkyModel' object has no attribute 'ra'"):
        sky.ra  # noqa: B018
    with pytest.raises(AttributeError):
        sky.no_such_thing  # noqa: B018
    hpx = rsk.SkyModel(nside=2, stokes=np.ones((4, 1, 48)), spectral_type="flat", hpx_frame="icrs")
    with pytest.warns(UserWarning, match="It is more efficient to use the `get_lon_lat` method"):
        ra = hpx.ra
    assert torch.equal(ra, hpx.get_lon_lat()[0])
    assert "stokes" in sky.ncomponent_length_params and "alt_az" not in sky.ncomponent_length_params
    import copy as _c

    clone = _c.copy(sky)  # __getattr__ must not recurse on half-built objects
    assert clone.frame == "galactic" and sky.copy() == sky
