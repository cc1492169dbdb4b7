"""GPU parity tests added in round 2: cache invalidation after in-place unit changes, ``at_frequencies``
book-keeping, the channel-picking interpolation kinds, NaN handling beyond 64 subbands, subband ``visibility_calc``,
SkyH5 round trips of point objects that came from HEALPix maps, the component-major stored-cube path, ERFA's check
vectors through the position kernels, the overlap of positions with the previous step's coherency kernel."""

import warnings

import numpy as np
import pytest
import scipy.interpolate
import torch

from oracle import skymodel_np as S

from test_gpu_parity import (GLEAM_BANDS, assert_coherency_close, make_model, polarized_stokes, random_sky, rsk,  # noqa: F401
                             site)

pytestmark = pytest.mark.gpu


def test_unit_change_invalidates_spline_cache(rsk, site):
    """the banded spline coefficients are cached per catalogue; an in-place Jy -> K conversion (a
    librsk kernel, invisible to torch's version counter) must drop them."""
    time, loc = site
    rng = np.random.default_rng(5)
    sky, lon, lat, st, kw = make_model(rsk, rng, "subband", 1501)
    freqs = np.linspace(80e6, 220e6, 33)
    sky.update_positions(time, loc)
    sky.coherency_calc(freqs=freqs)  # fills the coefficient cache in Jy
    sky.jansky_to_kelvin()
    st_k = st * S.jy_to_ksr(kw["freq_array"])[None, :, None]
    new = sky.at_frequencies(freqs, inplace=False)
    ref = S.at_frequencies(st_k, "subband", freqs, freq_array=kw["freq_array"])
    scale = np.max(np.abs(st_k), axis=(0, 1))[None, None, :]
    assert np.max(np.abs(new.stokes.cpu().numpy() - ref) / scale) < 1e-12
    # and the fused path sees the new units too
    got = sky.coherency_calc(freqs=freqs, pad=False).cpu().numpy()
    above = sky.above_horizon.cpu().numpy()
    want = S.coherency_calc(ref, lon, lat, sky.alt_az.cpu().numpy(), above, S.calc_average_rotation_matrix(sky._astrom),
                            polarized=S.polarized_indices(st), vectorized=True)
    assert_coherency_close(got, want, ref[0][:, above])
    # healpix_to_point scales the Stokes in place the same way
    hp_st = polarized_stokes(rng, 48, 4)
    hp_sky = rsk.SkyModel(nside=2, hpx_order="ring", stokes=hp_st, spectral_type="subband", freq_array=GLEAM_BANDS[:4],
                          stokes_unit="Jy / sr")
    f2 = np.linspace(80e6, 95e6, 7)
    a = hp_sky.at_frequencies(f2, inplace=False).stokes.cpu().numpy()
    hp_sky.healpix_to_point(to_jy=False)
    b = hp_sky.at_frequencies(f2, inplace=False).stokes.cpu().numpy()
    assert np.allclose(b, a * (4 * np.pi / 48), rtol=1e-13, atol=0)


@pytest.mark.filterwarnings("ignore:freq_edge_array not set")
def test_at_frequencies_bookkeeping(rsk):
    """skymodel.py:2179-2180, :2375-2386: freq_edge_array follows the channel subset of a 'full' model and is dropped
    when a subband model becomes 'full'; the request becomes freq_array; run_check runs SkyModel.check."""
    rng = np.random.default_rng(6)
    n = 300
    lon, lat = random_sky(rng, n)
    edges = np.stack([GLEAM_BANDS - 4e6, GLEAM_BANDS + 4e6])
    sub = rsk.SkyModel(lon=lon, lat=lat, stokes=polarized_stokes(rng, n, 20), spectral_type="subband",
                       freq_array=GLEAM_BANDS, freq_edge_array=edges)
    full = sub.at_frequencies(np.linspace(80e6, 200e6, 5), inplace=False)
    assert full.freq_edge_array is None and sub.freq_edge_array is not None
    assert np.array_equal(full.freq_array, np.linspace(80e6, 200e6, 5))
    stored = rsk.SkyModel(lon=lon, lat=lat, stokes=polarized_stokes(rng, n, 20), spectral_type="full",
                          freq_array=GLEAM_BANDS, freq_edge_array=edges)
    stored.at_frequencies(GLEAM_BANDS[[2, 5, 11]])
    assert np.array_equal(stored.freq_edge_array, edges[:, [2, 5, 11]])
    assert stored.Nfreqs == 3 and stored.stokes.shape == (4, 3, n)
    # run_check: negative Stokes I gives the reference's warning, run_check=False stays silent
    neg = polarized_stokes(rng, n, 1)
    neg[0, 0, 3] = -1.0
    flat = rsk.SkyModel(lon=lon, lat=lat, stokes=neg, spectral_type="flat")
    with pytest.warns(UserWarning, match="Some Stokes I values are negative"):
        flat.at_frequencies([1e8, 2e8], inplace=False)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        flat.at_frequencies([1e8, 2e8], inplace=False, run_check=False)


@pytest.mark.parametrize("kind", ["nearest", "nearest-up", "zero", "previous", "next", 0])
def test_at_frequencies_step_kinds(rsk, site, kind):
    """Any interp1d kind goes through at_frequencies upstream (skymodel.py:2362-2368); the channel-picking ones are
    bit-exact: the value is one of the stored values, picked as scipy picks it (ties at knots and midpoints included)."""
    time, loc = site
    rng = np.random.default_rng(8)
    sky, lon, lat, st, kw = make_model(rsk, rng, "subband", 777)
    fa = kw["freq_array"]
    mids = 0.5 * (fa[1:] + fa[:-1])
    freqs = np.concatenate([fa, mids, np.nextafter(fa[:-1], np.inf), np.nextafter(mids, 0), rng.uniform(fa[0], fa[-1], 200)])
    new = sky.at_frequencies(freqs, inplace=False, freq_interp_kind=kind)
    ref = scipy.interpolate.interp1d(fa, st, axis=1, kind=kind)(freqs)
    assert np.array_equal(new.stokes.cpu().numpy(), ref)
    sky.update_positions(time, loc)
    got = sky.coherency_calc(freqs=freqs[:40], freq_interp_kind=kind, pad=False).cpu().numpy()
    above = sky.above_horizon.cpu().numpy()
    want = S.coherency_calc(ref[:, :40], lon, lat, sky.alt_az.cpu().numpy(), above,
                            S.calc_average_rotation_matrix(sky._astrom), polarized=S.polarized_indices(st), vectorized=True)
    assert_coherency_close(got, want, ref[0][:40][:, above])


@pytest.mark.filterwarnings("ignore:Some Stokes values are NaNs")
@pytest.mark.filterwarnings("ignore:invalid value encountered")
@pytest.mark.filterwarnings("ignore:.*components had")
def test_nan_handling_with_more_than_64_subbands(rsk, site):
    """The valid-channel patterns are multi-word bit masks: 100 stored channels, NaNs on either side of the 64-bit
    word boundary."""
    time, loc = site
    rng = np.random.default_rng(9)
    n, nfin = 400, 100
    sky0, lon, lat, st, kw = make_model(rsk, rng, "subband", n, nf_in=nfin)
    for c in range(0, n, 3):
        st[rng.integers(0, 4), rng.choice(nfin, rng.integers(1, 9), replace=False), c] = np.nan
    st[2, 60:70, 5] = np.nan
    st[0, 64:, 7] = np.nan
    st[:, :, 11] = np.nan
    sky = rsk.SkyModel(lon=lon, lat=lat, stokes=st, spectral_type="subband", freq_array=kw["freq_array"])
    freqs = np.linspace(kw["freq_array"][0], kw["freq_array"][-1], 57)
    got = sky.at_frequencies(freqs, inplace=False, nan_handling="clip", run_check=False).stokes.cpu().numpy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = S.at_frequencies(st, "subband", freqs, freq_array=kw["freq_array"], nan_handling="clip")
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = np.isfinite(ref)
    scale = np.broadcast_to(np.nanmax(np.abs(st), axis=(0, 1))[None, None, :], ref.shape)
    assert np.max(np.abs(got[ok] - ref[ok]) / scale[ok]) < 1e-11
    assert np.isnan(ref).any()


@pytest.mark.filterwarnings("ignore:Some Stokes values are NaNs")
@pytest.mark.filterwarnings("ignore:invalid value encountered")
@pytest.mark.filterwarnings("ignore:.*components had")
@pytest.mark.parametrize("with_nans", [False, True])
def test_visibility_calc_subband(rsk, site, with_nans):
    """visibility_calc on a subband model == at_frequencies() followed by visibility_calc on the 'full' model."""
    time, loc = site
    rng = np.random.default_rng(10)
    sky, lon, lat, st, kw = make_model(rsk, rng, "subband", 3001)
    if with_nans:
        st[1, 3:6, ::17] = np.nan
        st[0, :2, 5::29] = np.nan
        sky = rsk.SkyModel(lon=lon, lat=lat, stokes=st, spectral_type="subband", freq_array=kw["freq_array"])
        keep = ~np.isnan(st).any(axis=(0, 1))
    freqs = np.linspace(90e6, 210e6, 13)
    uvw = np.array([[14.0, 3.0, 0.5], [-30.0, 100.0, 2.0]])
    sky.update_positions(time, loc)
    full = sky.at_frequencies(freqs, inplace=False, run_check=False)
    full.update_positions(time, loc)
    if with_nans:  # a NaN-free sum needs the all-finite components only: compare through select
        cube = full.stokes.cpu().numpy()
        assert np.isfinite(cube).all()  # 'clip' fills every gap of these patterns
    a = sky.visibility_calc(beam="short_dipole", freqs=freqs, uvw=uvw).cpu().numpy()
    b = full.visibility_calc(beam="short_dipole", uvw=uvw).cpu().numpy()
    assert a.shape == (2, 2, 2, 13)
    assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(b))


def test_skyh5_roundtrip_keeps_healpix_description_of_point_objects(rsk, tmp_path):
    """a point object made by healpix_to_point keeps nside / hpx_order / hpx_inds
    (skymodel.py:1716-1727); they survive write_skyh5 -> from_skyh5, so _point_to_healpix still works."""
    rng = np.random.default_rng(12)
    st = polarized_stokes(rng, 192, 3)
    sky = rsk.SkyModel(nside=4, hpx_order="nested", stokes=st, spectral_type="full", freq_array=[1e8, 1.1e8, 1.2e8],
                       stokes_unit="K", hpx_frame="icrs")
    ref = sky.copy()
    sky.healpix_to_point()
    assert sky.component_type == "point" and sky.nside == 4 and sky.hpx_order == "nested"
    path = tmp_path / "pt.skyh5"
    sky.write_skyh5(path)
    back = rsk.SkyModel.from_skyh5(path)
    assert back.component_type == "point" and back.nside == 4 and back.hpx_order == "nested"
    assert torch.equal(back.hpx_inds, torch.arange(192, device=back.hpx_inds.device))
    back.filename = sky.filename  # (the reader records the file name, as upstream)
    assert back == sky
    back._point_to_healpix()
    assert back.component_type == "healpix" and back.stokes_unit == "K"
    assert torch.allclose(back.stokes, ref.stokes, rtol=1e-13, atol=0)
    # healpix_to_point leaves time / position attributes alone, as upstream
    loc = rsk.EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
    ref.update_positions(rsk.Time.from_utc("2015-03-01 00:00:00"), loc)
    ref.healpix_to_point()
    assert ref.alt_az is not None and ref.time is not None


def test_copy_does_not_share_private_state(rsk, site):
    time, loc = site
    rng = np.random.default_rng(13)
    sky = make_model(rsk, rng, "spectral_index", 900)[0]
    freqs = np.linspace(1e8, 2e8, 9)
    sky.update_positions(time, loc)
    host = torch.empty(4 * 9 * 900, dtype=torch.complex128).pin_memory()
    sky.coherency_calc_to_host(freqs, host, slab=4, packed=True, nthreads=1)
    assert "_staging" in sky.__dict__
    other = sky.copy()
    assert "_staging" not in other.__dict__ and "_copy_stream" not in other.__dict__
    a = sky.coherency_calc(freqs=freqs, pad=False)
    b = other.coherency_calc(freqs=freqs, pad=False)  # uses the cloned rotation coefficients
    assert torch.equal(a, b)


@pytest.mark.parametrize("density", [0.0, 0.07, 0.5, 0.93, 1.0])
@pytest.mark.parametrize("n,nf", [(5003, 37), (256, 1), (70001, 8), (3001, 13)])
def test_full_coherency_component_major_path(rsk, site, density, n, nf):
    """coherency_calc() on stored 'full' Stokes (the reference-shaped call, skymodel.py:2685, :2704-2710) reading the
    component-major copy of the cube (RskSpectrum.stokes_cm): equal, bit for bit, to the kernel that gathers from the
    (4, Nf, N) cube, for every horizon-mask density, with subsets of channels and column offsets."""
    time, loc = site
    rng = np.random.default_rng(n + nf)
    lon, lat = random_sky(rng, n)
    st = polarized_stokes(rng, n, nf) * (1 + 0.1 * rng.normal(size=(4, nf, n)))
    st[1:, :, rng.uniform(size=n) < 0.3] = 0.0
    fa = np.linspace(100e6, 200e6, nf)
    sky = rsk.SkyModel(lon=lon, lat=lat, stokes=st, spectral_type="full", freq_array=fa)
    sky.update_positions(time, loc)
    above = rng.uniform(size=n) < density
    sky.above_horizon = torch.as_tensor(above).cuda()
    import os
    os.environ["RSK_FULL_CM"] = "0"
    try:
        ref = sky.coherency_calc(pad=False)
    finally:
        os.environ["RSK_FULL_CM"] = "1"
    try:
        got = sky.coherency_calc(pad=False)
        assert got.shape == (2, 2, nf, int(above.sum()))
        assert torch.equal(got, ref)
        # padded rows, a column offset, and the count left on the device
        wide = torch.full((2, 2, nf, n + 48), complex(3, 3), dtype=torch.complex128, device="cuda")
        out, cnt = sky.coherency_calc(out=wide, col0=16, sync=False)
        c = int(cnt.item())
        assert c == int(above.sum())
        assert torch.equal(out[..., 16:16 + c], ref)
        assert bool((out[..., :16] == complex(3, 3)).all()) and bool((out[..., 16 + c:] == complex(3, 3)).all())
        if nf >= 8:
            sub = sky.coherency_calc(freqs=fa[[1, 4, 6]], pad=False)
            assert torch.equal(sub, ref[:, :, [1, 4, 6]])
    finally:
        os.environ.pop("RSK_FULL_CM", None)
    want = S.coherency_calc(st, lon, lat, sky.alt_az.cpu().numpy(), above, S.calc_average_rotation_matrix(sky._astrom),
                            polarized=S.polarized_indices(st), vectorized=True)
    assert_coherency_close(got.cpu().numpy(), want, st[0][:, above])


def test_device_chain_reproduces_erfa_check_values(rsk, site):
    """The position kernel against ERFA's own published test vectors (tests/golden/erfa_check_values.json), not against
    this repo's restatement: (a) eraAtciqz on the eraApci13 context -- light deflection, aberration, bias-precession-
    nutation -- read off through a trivial observer (pole site, no Earth rotation: alt = di, az = pi - ri); (b) eraAtioq
    on the eraApio13 observer -- polar motion, diurnal aberration, latitude rotation, A/B refraction -- fed through a
    transparent star part (no deflection, no aberration, identity bpn)."""
    import json
    import os

    from oracle import erfa_np as E

    time, loc = site
    with open(os.path.join(os.path.dirname(__file__), "golden", "erfa_check_values.json")) as fh:
        k = json.load(fh)
    st = np.zeros((4, 1, 1)); st[0] = 1.0
    # (a) folded context
    star = k["apci13"]
    a = np.zeros(31)
    a[E.I_PMT] = star["pmt"]; a[E.I_EB:E.I_EB + 3] = star["eb"]; a[E.I_EH:E.I_EH + 3] = star["eh"]; a[E.I_EM] = star["em"]
    a[E.I_V:E.I_V + 3] = star["v"]; a[E.I_BM1] = star["bm1"]; a[E.I_BPN:E.I_BPN + 9] = np.asarray(star["bpn"]).ravel()
    a[E.I_PHI] = np.pi / 2; a[E.I_SPHI] = 1.0; a[E.I_CPHI] = 0.0
    q = k["atciqz"]
    sky = rsk.SkyModel(ra=[q["rc"]], dec=[q["dc"]], stokes=st, spectral_type="flat")
    sky.update_positions(time, loc, astrom=a)
    alt, az = sky.alt_az.cpu().numpy()[:, 0]
    assert abs(alt - q["di"]) < 1e-14, alt - q["di"]
    assert abs(az - (np.pi - q["ri"]) % (2 * np.pi)) < 1e-14
    # (b) general context: refraction and diurnal aberration
    b = np.zeros(31)
    b[E.I_EH:E.I_EH + 3] = [1.0, 0.0, 0.0]; b[E.I_EM] = 1e30; b[E.I_BM1] = 1.0
    b[E.I_BPN:E.I_BPN + 9] = np.eye(3).ravel()
    obs = k["apio13"]
    for key, slot in (("along", E.I_ALONG), ("phi", E.I_PHI), ("xpl", E.I_XPL), ("ypl", E.I_YPL), ("sphi", E.I_SPHI),
                      ("cphi", E.I_CPHI), ("diurab", E.I_DIURAB), ("eral", E.I_ERAL), ("refa", E.I_REFA), ("refb", E.I_REFB)):
        b[slot] = obs[key]
    o = k["atioq"]
    sky = rsk.SkyModel(ra=[o["ri"]], dec=[o["di"]], stokes=st, spectral_type="flat")
    sky.update_positions(time, loc, astrom=b)
    alt, az = sky.alt_az.cpu().numpy()[:, 0]
    assert abs(az - o["aob"]) < 1e-14, az - o["aob"]
    assert abs(alt - (np.pi / 2 - o["zob"])) < 1e-14, alt - (np.pi / 2 - o["zob"])
    # the same two through the sweep kernel (T = 2 contexts of each kind)
    for astrom, ra, dec, want_alt, want_az in ((a, q["rc"], q["dc"], q["di"], (np.pi - q["ri"]) % (2 * np.pi)),
                                               (b, o["ri"], o["di"], np.pi / 2 - o["zob"], o["aob"])):
        sky = rsk.SkyModel(ra=[ra], dec=[dec], stokes=st, spectral_type="flat")
        for lmn in (True, False):  # False: the sweep-grade angle routine (below 1e-13 rad by construction)
            alt_az, _, _ = sky.positions_sweep([time, time], loc, astroms=np.stack([astrom, astrom]), want_lmn=lmn)
            got = alt_az.cpu().numpy()
            tol = 1e-14 if lmn else 1e-13
            assert np.abs(got[:, 0, 0] - want_alt).max() < tol and np.abs(got[:, 1, 0] - want_az).max() < tol


def test_astropy_context_bytes_give_astropy_alt_az(rsk):
    """With astropy present (not on the build / GPU image: skipped there) the reference's own call,
    ``skycoord.transform_to(AltAz(obstime, location))`` (skymodel.py:2451-2454), and this package fed the real
    ``eraASTROM`` of ``erfa_astrom.get().apco`` must agree to 1 uas -- the N2 row closed by astropy itself."""
    astropy = pytest.importorskip("astropy")  # noqa: F841
    import astropy.units as u
    from astropy.coordinates import AltAz, EarthLocation, SkyCoord
    from astropy.time import Time

    rng = np.random.default_rng(14)
    n = 5000
    ra, dec = random_sky(rng, n)
    loc = EarthLocation.from_geodetic(lon=21.428305555 * u.deg, lat=-30.72152777 * u.deg, height=1073.0 * u.m)
    time = Time("2015-03-01 00:00:00", scale="utc", location=loc)
    want = SkyCoord(ra=ra * u.rad, dec=dec * u.rad, frame="icrs").transform_to(AltAz(obstime=time, location=loc))
    st = np.zeros((4, 1, n)); st[0] = 1.0
    sky = rsk.SkyModel(ra=ra, dec=dec, stokes=st, spectral_type="flat")
    sky.update_positions(time, loc)  # astropy objects: the context is astropy's own eraASTROM
    got = sky.alt_az.cpu().numpy()
    uas = np.pi / 180 / 3600 * 1e-6
    assert np.abs(got[0] - want.alt.rad).max() < uas
    d_az = np.abs(got[1] - want.az.rad)
    assert (np.minimum(d_az, 2 * np.pi - d_az) * np.cos(want.alt.rad)).max() < uas
    assert np.array_equal(sky.above_horizon.cpu().numpy(), want.alt.rad > 0)


@pytest.mark.parametrize("spectral_type,frame", [("spectral_index", "fk5"), ("subband", "icrs"), ("flat", "icrs")])
def test_position_overlap_is_bit_identical(rsk, site, spectral_type, frame):
    """enable_position_overlap(): K1 + compaction of step k + 1 on a side stream beside K2 of step k, two buffer sets.
    Every step's positions, mask and coherency equal the plain sequence bit for bit, also when the caller reads the
    position arrays between the two calls and when a big kernel is still running on the caller's stream."""
    time, loc = site
    rng = np.random.default_rng(99)
    n, nf = 60_000, 64
    sky = make_model(rsk, rng, spectral_type, n, frame=frame)[0]
    ref = sky.copy()
    freqs = np.linspace(100e6, 160e6, nf)
    sky.enable_position_overlap()
    buf = torch.empty((2, 2, nf, -(-n // 16) * 16), dtype=torch.complex128, device="cuda")
    outs, alts, masks, cnts = [], [], [], []
    for k in range(7):
        sky.update_positions(time + 0.013 * k, loc)
        if k % 2:
            alts.append(sky.alt_az.clone())  # a read on the caller's stream right after the update
        else:
            alts.append(None)
        _, cnt = sky.coherency_calc(freqs=freqs, out=buf, sync=False)
        if alts[-1] is None:
            alts[-1] = sky.alt_az.clone()
        masks.append(sky.above_horizon.clone())
        outs.append(buf.clone())
        cnts.append(cnt.clone())
    torch.cuda.synchronize()
    for k in range(7):
        ref.update_positions(time + 0.013 * k, loc)
        want = ref.coherency_calc(freqs=freqs, pad=False)
        c = int(cnts[k].item())
        assert c == want.shape[-1]
        assert torch.equal(alts[k], ref.alt_az) and torch.equal(masks[k], ref.above_horizon)
        assert torch.equal(outs[k][..., :c], want)
    # the exactly-sized call and a copy of the model work as before
    sky.update_positions(time + 0.5, loc)
    ref.update_positions(time + 0.5, loc)
    assert torch.equal(sky.coherency_calc(freqs=freqs, pad=False), ref.coherency_calc(freqs=freqs, pad=False))
    other = sky.copy()
    assert "_overlap" not in other.__dict__
    other.update_positions(time + 0.6, loc)
    ref.update_positions(time + 0.6, loc)
    assert torch.equal(other.alt_az, ref.alt_az)
    sky.enable_position_overlap(False)
    sky.update_positions(time + 0.7, loc)
    ref.update_positions(time + 0.7, loc)
    assert torch.equal(sky.alt_az, ref.alt_az)


def test_position_overlap_survives_select_and_healpix(rsk, site):
    """The overlap's buffer sets follow the component list (select) and work for HEALPix maps (pixel centres generated
    in the kernel on the side stream)."""
    time, loc = site
    rng = np.random.default_rng(101)
    sky = make_model(rsk, rng, "spectral_index", 9000)[0]
    ref = sky.copy()
    sky.enable_position_overlap()
    freqs = np.linspace(1e8, 2e8, 5)
    sky.update_positions(time, loc)
    keep = torch.arange(0, 9000, 3, device="cuda")
    sky.select(component_inds=keep)
    ref.select(component_inds=keep)
    for k in range(3):
        sky.update_positions(time + 0.02 * (k + 1), loc)
        ref.update_positions(time + 0.02 * (k + 1), loc)
        assert sky.alt_az.shape == (2, 3000) and torch.equal(sky.alt_az, ref.alt_az)
        assert torch.equal(sky.coherency_calc(freqs=freqs, pad=False), ref.coherency_calc(freqs=freqs, pad=False))
    st = polarized_stokes(rng, 12 * 16 * 16)
    hp_sky = rsk.SkyModel(nside=16, hpx_order="nested", stokes=st, spectral_type="flat", stokes_unit="K")
    hp_ref = hp_sky.copy()
    hp_sky.enable_position_overlap()
    for k in range(3):
        hp_sky.update_positions(time + 0.1 * k, loc)
        hp_ref.update_positions(time + 0.1 * k, loc)
        assert torch.equal(hp_sky.alt_az, hp_ref.alt_az) and torch.equal(hp_sky.above_horizon, hp_ref.above_horizon)
        assert torch.equal(hp_sky.coherency_calc(pad=False), hp_ref.coherency_calc(pad=False))


@pytest.mark.filterwarnings("ignore:Some Stokes values are NaNs")
@pytest.mark.filterwarnings("ignore:invalid value encountered")
@pytest.mark.filterwarnings("ignore:.*components had")
def test_apparent_coherency_with_nan_spectra(rsk, site):
    """apparent_coherency_calc on a NaN-bearing subband catalogue (round 1: NotImplementedError): J C J^H of what
    coherency_calc gives, NaN columns included."""
    time, loc = site
    rng = np.random.default_rng(102)
    sky, lon, lat, st, kw = make_model(rsk, rng, "subband", 2000)
    st[2, 4:9, ::13] = np.nan
    st[:, :, 7] = np.nan
    sky = rsk.SkyModel(lon=lon, lat=lat, stokes=st, spectral_type="subband", freq_array=kw["freq_array"])
    freqs = np.linspace(90e6, 200e6, 9)
    sky.update_positions(time, loc)
    coh = sky.coherency_calc(freqs=freqs, pad=False)
    app = sky.apparent_coherency_calc(beam="uniform", freqs=freqs, pad=False)
    # identity Jones: the same numbers wherever the whole 2x2 matrix is finite; a NaN in any entry of C spreads over
    # all of J C J^H (0 * NaN = NaN in IEEE arithmetic), exactly as the einsum of tests/test_skymodel.py:1464-1466 does
    bad = torch.isnan(torch.view_as_real(coh)).any(dim=-1).any(dim=0).any(dim=0)  # (Nf, N_above)
    assert bool(bad.any()) and not bool(bad.all())
    assert torch.equal(app[:, :, ~bad], coh[:, :, ~bad])
    assert bool(torch.isnan(torch.view_as_real(app[:, :, bad])).any(dim=-1).all())
