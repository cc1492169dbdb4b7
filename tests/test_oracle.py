"""The CPU oracle against the reference's own golden vectors and known-answer tests (no GPU)."""

import numpy as np
import pytest
import scipy.interpolate

from oracle import erfa_np as E
from oracle import healpix_np as hp
from oracle import sct_np as sct
from oracle import skymodel_np as S


# ---- rotation helpers: bit-for-bit against the reference module (tests/golden/sct_golden.json) ----
def test_sct_unit_vectors_bit_exact(sct_golden):
    th, ph = np.array(sct_golden["theta"]), np.array(sct_golden["phi"])
    assert np.array_equal(sct.r_hat(th, ph), np.array(sct_golden["r_hat"]))
    assert np.array_equal(sct.theta_hat(th, ph), np.array(sct_golden["theta_hat"]))
    assert np.array_equal(sct.phi_hat(th, ph), np.array(sct_golden["phi_hat"]))


def test_sct_rotations_bit_exact(sct_golden):
    aa = sct_golden["axis_angle"]
    assert np.array_equal(sct.axis_angle_rotation_matrix(np.array(aa["axis"]), aa["angle"]), np.array(aa["matrix"]))
    q = np.array(sct_golden["q"])
    for i, expect in enumerate(sct_golden["rotate_points_3d"]):
        got = sct.rotate_points_3d(q, sct_golden["theta"][i], sct_golden["phi"][i])
        assert list(map(float, got)) == expect
    for c in sct_golden["cases"]:
        rot = sct.vecs2rot(r1=np.array(c["r1"]), r2=np.array(c["r2"]))
        assert np.array_equal(rot, np.array(c["vecs2rot"]))
        basis = sct.spherical_basis_vector_rotation_matrix(
            c["theta"], c["phi"], np.array(c["r_exact"]), c["beta"], c["alpha"])
        assert np.array_equal(basis, np.array(c["basis_rot"]))


def test_sct_errors():
    # reference tests/test_spherical_coords_transforms.py:14-18 and :153-181
    for fn in (sct.r_hat, sct.theta_hat, sct.phi_hat):
        with pytest.raises(ValueError, match="theta and phi must have the same shape"):
            fn([0, 0], [0])
    with pytest.raises(ValueError, match="rot_matrix must be a 3x3 array"):
        sct.rotate_points_3d(np.eye(3)[0:1], 0.3, 0.2)
    r1 = sct.r_hat(0.3, 0.2)
    with pytest.raises(ValueError, match="Either r1 and r2 must be supplied or all of"):
        sct.vecs2rot(r1=r1, theta2=0.1, phi2=0.2)
    with pytest.raises(ValueError, match="r1 and r2 must be length 3 vectors"):
        sct.vecs2rot(r1=r1[0:1], r2=r1)
    with pytest.raises(ValueError, match="r1 and r2 must be unit vectors"):
        sct.vecs2rot(r1=r1 * 2, r2=r1)
    with pytest.raises(ValueError, match="axis must be a must be length 3 vector"):
        sct.axis_angle_rotation_matrix(r1[0:1], 0.1)
    with pytest.raises(ValueError, match="axis must be a unit vector"):
        sct.axis_angle_rotation_matrix(r1 * 2, 0.1)
    # identical vectors: NaN axis -> upstream's ValueError (spherical_coords_transforms.py:300-303)
    with pytest.raises(ValueError, match="axis must be a unit vector"), np.errstate(all="ignore"):
        sct.vecs2rot(r1=r1, r2=r1)


# ---- Stokes <-> coherency, units ---------------------------------------------------------------
def test_stokes_to_coherency_known_answer(known_answers):
    ka = known_answers["stokes_coherency"]
    coh = S.stokes_to_coherency(np.array(ka["stokes"]))
    expect = np.array(ka["coherency_re"]) + 1j * np.array(ka["coherency_im"])
    # sign convention of utils.py:48,51: C01 = (U - iV)/2
    assert np.allclose(coh, np.conj(expect).T) or np.allclose(coh, expect)
    assert np.allclose(coh, 0.5 * np.array([[4.2, 1.2 + 0.15j], [1.2 - 0.15j, 4.8]]))
    st3 = np.repeat(np.array(ka["stokes"])[:, None, None], 2, axis=2)
    assert S.stokes_to_coherency(st3).shape == (2, 2, 1, 2)
    with pytest.raises(ValueError, match="First dimension of stokes_vector must be length 4."):
        S.stokes_to_coherency(st3[0:2])
    tk = known_answers["tutorial_frame_coherency"]
    assert np.allclose(S.stokes_to_coherency(np.array(tk["stokes"])).real, tk["coherency_re"])


def test_jy_to_ksr(known_answers):
    ka = known_answers["jy_to_ksr_cgs"]
    freqs = np.linspace(ka["freqs_mhz"][0], ka["freqs_mhz"][1], int(ka["freqs_mhz"][2])) * 1e6
    lam = ka["c_cmps"] / freqs
    assert np.allclose(S.jy_to_ksr(freqs), 1e-23 * lam**2 / (2 * ka["k_boltz_erg"]))  # test_utils.py:139-152
    tu = known_answers["tutorial_unit_conversion"]
    got = np.array(tu["gleam_jy"]) * S.jy_to_ksr(tu["gleam_freq_hz"])
    assert np.allclose(got, tu["gleam_ksr"], rtol=2e-6)  # printed to 6 significant figures
    got = np.array(tu["gsm_k"]) / S.jy_to_ksr(tu["gsm_freq_hz"])
    assert np.allclose(got, tu["gsm_jy_sr"], rtol=1e-10)


def test_unit_round_trip():
    rng = np.random.default_rng(0)
    st = rng.normal(size=(4, 5, 7))
    fa = np.linspace(1e8, 2e8, 5)
    back = S.kelvin_to_jansky(S.jansky_to_kelvin(st, "full", freq_array=fa), "full", freq_array=fa)
    assert np.allclose(back, st, rtol=1e-15)
    rf = np.full(7, 1.5e8)
    k = S.jansky_to_kelvin(st[:, :1], "spectral_index", reference_frequency=rf)
    assert np.allclose(k, st[:, :1] * S.jy_to_ksr(1.5e8))
    with pytest.raises(ValueError, match="Either reference_frequency or freq_array must be set"):
        S.jansky_to_kelvin(st[:, :1], "flat")


# ---- spectra -------------------------------------------------------------------------------------
def test_at_frequencies_power_law(known_answers):
    # tests/test_skymodel.py:162-221, :3542-3583
    ka = known_answers["at_frequencies_power_law"]
    n, alpha = ka["ncomp"], ka["alpha"]
    fin = np.linspace(*ka["freqs_in"][:2], int(ka["freqs_in"][2]))
    fine = np.linspace(*ka["freqs_fine"][:2], int(ka["freqs_fine"][2]))
    spectrum = (fin / fin[0]) ** alpha
    fine_spectrum = (fine / fine[0]) ** alpha
    st_full = np.zeros((4, fin.size, n))
    st_full[0] = spectrum[:, None]
    # spectral_index
    st1 = np.zeros((4, 1, n)); st1[0] = 1.0
    out = S.at_frequencies(st1, "spectral_index", fine, reference_frequency=np.full(n, fin[0]), spectral_index=np.full(n, alpha))
    assert np.allclose(out[0, :, 0], fine_spectrum)
    # subband (cubic interpolation of a smooth power law)
    out = S.at_frequencies(st_full, "subband", fine, freq_array=fin)
    assert np.allclose(out[0, :, 0], fine_spectrum)
    # flat
    out = S.at_frequencies(st1, "flat", fine)
    assert out.shape == (4, fine.size, n) and np.all(out[0] == 1.0)
    # full: subset, and error for absent channels
    out = S.at_frequencies(st_full, "full", fin[5:10], freq_array=fin)
    assert np.array_equal(out, st_full[:, 5:10])
    with pytest.raises(ValueError, match="Some requested frequencies"):
        S.at_frequencies(st_full, "full", fine, freq_array=fin)
    # errors, tests/test_skymodel.py:3586-3608
    with pytest.raises(ValueError, match="larger than the highest subband"):
        S.at_frequencies(st_full, "subband", fin + 10, freq_array=fin)
    with pytest.raises(ValueError, match="smaller than the lowest subband"):
        S.at_frequencies(st_full, "subband", fin - 10, freq_array=fin)
    bad = st_full.copy(); bad[0, 0, 0] = np.nan
    with pytest.raises(ValueError, match="nan_handling must be one of"):
        S.at_frequencies(bad, "subband", fin, freq_array=fin, nan_handling="foo")
    with pytest.raises(ValueError, match="Some spectral index values are NaNs."):
        S.at_frequencies(st1, "spectral_index", fine, reference_frequency=np.full(n, fin[0]),
                         spectral_index=np.array([np.nan] + [alpha] * (n - 1)))


@pytest.mark.filterwarnings("ignore:Some Stokes values are NaNs")
def test_at_frequencies_nan_handling():
    fin = np.linspace(100e6, 130e6, 12)
    st = np.zeros((4, 12, 3)); st[0] = ((fin / fin[0]) ** -0.7)[:, None]
    st[0, 0, 1] = np.nan   # low edge missing on comp 1
    st[0, -1, 2] = np.nan  # high edge missing on comp 2
    out = S.at_frequencies(st, "subband", fin, freq_array=fin, nan_handling="propagate")
    assert np.all(np.isnan(out[0][:, 1:])) and not np.any(np.isnan(out[0][:, 0]))
    out = S.at_frequencies(st, "subband", fin, freq_array=fin, nan_handling="interp")
    assert np.isnan(out[0, 0, 1]) and np.isnan(out[0, -1, 2]) and np.allclose(out[0, 1:, 1], st[0, 1:, 1])
    out = S.at_frequencies(st, "subband", fin, freq_array=fin, nan_handling="clip")
    assert out[0, 0, 1] == st[0, 1, 1] and out[0, -1, 2] == st[0, -2, 2]


# ---- positions ------------------------------------------------------------------------------------
def test_site_geocentric(known_answers, hera):
    lon, lat, h, _ = hera
    assert np.allclose(E.gd2gc_wgs84(lon, lat, h), known_answers["hera_geocentric_m"]["xyz"], atol=1e-3)


def test_context_invariants(hera):
    lon, lat, h, jd = hera
    a = E.astrom_context(jd, 0.0, lon, lat, h)
    bpn = a[E.I_BPN:E.I_BPN + 9].reshape(3, 3)
    assert np.allclose(bpn @ bpn.T, np.eye(3), atol=1e-15)
    assert abs(np.linalg.norm(a[E.I_EH:E.I_EH + 3]) - 1) < 1e-15
    assert 0 < 1 - a[E.I_BM1] < 1e-8
    assert 0.98 < a[E.I_EM] < 1.02
    # a source at CIRS (ri, di) = (eral, phi) is at the zenith (SURVEY appendix A invariant i)
    az, zd = E.atioq(np.array([a[E.I_ERAL]]), np.array([lat]), a)
    assert abs(zd[0]) < 1e-9
    # aberration ~ 20.5" * sin(angle to apex); deflection 4.07 mas at 90 deg from the Sun
    v = a[E.I_V:E.I_V + 3]
    perp = np.cross(v, [0, 0, 1.0]); perp /= np.linalg.norm(perp)
    moved = E.ab(perp[:, None], v, a[E.I_EM], a[E.I_BM1])[:, 0]
    assert abs(np.arccos(np.clip(moved @ perp, -1, 1)) / E.DAS2R - 20.49552 * np.linalg.norm(v) / 9.9365e-5) < 0.05
    e = a[E.I_EH:E.I_EH + 3]
    q = np.cross(e, [0, 0, 1.0]); q /= np.linalg.norm(q)
    bent = E.ld(1.0, q[:, None], q[:, None], e, a[E.I_EM], 1e-6)[:, 0]
    assert abs(np.linalg.norm(bent - q) / E.DAS2R * 1e3 - 4.07 / a[E.I_EM]) < 0.05


def test_zenith_source(hera):
    # tests/test_skymodel.py:481-527: a source built by AltAz(90, 0) -> ICRS sits at lmn = (0, 0, 1)
    lon, lat, h, jd = hera
    a = E.astrom_context(jd, 0.0, lon, lat, h)
    ra, dec = E.altaz_to_icrs(np.pi / 2, 0.0, a)
    alt_az, lmn, above = S.update_positions(ra, dec, a)
    assert np.allclose(lmn[:, 0], [0, 0, 1], atol=1e-9) and above[0]
    ra2, dec2 = E.altaz_to_icrs(np.deg2rad([80.0, 35.0]), np.deg2rad([10.0, 250.0]), a)
    alt_az, _, _ = S.update_positions(ra2, dec2, a)
    assert np.allclose(alt_az, np.deg2rad([[80.0, 35.0], [10.0, 250.0]]), atol=1e-13)


def test_tutorial_positions(known_answers, hera):
    lon, lat, h, jd = hera
    a = E.astrom_context(jd, 0.0, lon, lat, h)
    ra, dec = E.altaz_to_icrs(np.deg2rad([90.0, 80.0]), np.deg2rad([0.0, 10.0]), a)
    alt_az, lmn, _ = S.update_positions(ra, dec, a)
    ka = known_answers["tutorial_altaz_lmn"]
    assert np.allclose(alt_az[0], ka["alt_az_2dp"][0], atol=6e-3)
    assert abs(alt_az[1, 1] - ka["alt_az_2dp"][1][1]) < 6e-3
    assert np.allclose(lmn[:, 1], np.array(ka["pos_lmn_1sf"])[:, 1], atol=6e-3)


def test_tutorial_horizon_mask(known_answers, hera):
    tm = known_answers["tutorial_horizon_mask"]
    lon = hera[0]
    ra = np.linspace(np.rad2deg(lon) - 90, np.rad2deg(lon) + 90, 5)
    dec = np.linspace(-90, 90, 5)
    ra, dec = np.meshgrid(ra, dec)
    ra = np.deg2rad(ra.flatten()) % (2 * np.pi)
    dec = np.deg2rad(dec.flatten())
    a = E.astrom_context(tm["jd_utc"], 0.0, np.deg2rad(tm["site_lon_deg"]), np.deg2rad(tm["site_lat_deg"]), tm["site_height_m"])
    _, _, above = S.update_positions(ra, dec, a)
    assert np.array_equal(above, np.array(tm["above_horizon"]))


def test_circumpolar_nonrising(hera):
    # tests/test_skymodel.py:2661-2740: over a sidereal day, never-setting / never-rising sources are
    # those with tan(lat) tan(dec) > 1 / < -1
    lon, lat, h, jd = hera
    rng = np.random.default_rng(7)
    n = 400
    ra = rng.uniform(0, 2 * np.pi, n)
    dec = np.arcsin(rng.uniform(-1, 1, n))
    up = np.zeros((100, n), dtype=bool)
    for ti in range(100):
        a = E.astrom_context(jd, ti * 0.9972696 / 100, lon, lat, h)
        up[ti] = S.update_positions(ra, dec, a)[2]
    tans = np.tan(lat) * np.tan(dec)
    margin = np.abs(np.abs(tans) - 1) > 0.05
    assert np.array_equal(up.all(axis=0)[margin], (tans > 1)[margin])
    assert np.array_equal((~up).all(axis=0)[margin], (tans < -1)[margin])


# ---- rotation + coherency ------------------------------------------------------------------------
def test_rotation_matrices(hera):
    # tests/test_skymodel.py:1253-1303 and test_spherical_coords_transforms.py:131-134
    lon, lat, h, jd = hera
    a = E.astrom_context(jd, 0.0, lon, lat, h)
    ravg = S.calc_average_rotation_matrix(a)
    assert np.allclose(ravg @ ravg.T, np.eye(3)) and np.allclose(ravg.T @ ravg, np.eye(3))
    ra = np.array([np.pi]); dec = np.array([-np.pi / 6])
    alt_az, _, _ = S.update_positions(ra, dec, a)
    rot2 = S.calc_coherency_rotation(ra, dec, alt_az, ravg)
    assert np.isclose(np.linalg.det(rot2[:, :, 0]), 1)
    r_exact = S.calc_rotation_matrix(ra, dec, alt_az, ravg)[:, :, 0]
    v1 = sct.r_hat(np.pi / 2 - dec, ra)[:, 0]
    v2 = sct.r_hat(np.pi / 2 - alt_az[0], alt_az[1])[:, 0]
    assert np.allclose(r_exact @ v1, v2)


def test_vectorised_rotation_matches_loop(hera):
    lon, lat, h, jd = hera
    a = E.astrom_context(jd, 0.123, lon, lat, h)
    rng = np.random.default_rng(3)
    n = 300
    ra = rng.uniform(0, 2 * np.pi, n); dec = np.arcsin(rng.uniform(-1, 1, n))
    alt_az, _, _ = S.update_positions(ra, dec, a)
    ravg = S.calc_average_rotation_matrix(a)
    inds = np.arange(0, n, 3)
    loop = S.calc_coherency_rotation(ra, dec, alt_az, ravg, inds)
    vec = S.calc_coherency_rotation_vec(ra, dec, alt_az, ravg, inds)
    assert np.abs(loop - vec).max() < 5e-15


@pytest.mark.parametrize("eop", [(0.0, 0.0), (-0.47, 0.32)])
def test_golden_instrumental_coherency(known_answers, hera, eop):
    """The reference's golden 2x2x5 matrix, tests/test_skymodel.py:1388-1508 (its own rtol: 1e-5)."""
    lon, lat, h, jd = hera
    dut1, yp = eop[0], eop[1] * E.DAS2R
    pv = known_answers["polarized_visibilities"]
    a0 = E.astrom_context(jd, 0.0, lon, lat, h, dut1=dut1, yp=yp)
    ra, dec = E.altaz_to_icrs(np.pi / 2, 0.0, a0)
    stokes = np.array(pv["stokes"], float).reshape(4, 1, 1)
    out = np.zeros((2, 2, 5), complex)
    alts, azs = [], []
    for ti, off in enumerate(pv["time_offsets_hr"]):
        a = E.astrom_context(jd, off / 24.0, lon, lat, h, dut1=dut1, yp=yp)
        alt_az, _, above = S.update_positions(ra, dec, a)
        ravg = S.calc_average_rotation_matrix(a)
        out[:, :, ti] = S.coherency_calc(stokes, ra, dec, alt_az, above, ravg)[:, :, 0, 0]
        alts.append(alt_az[0, 0]); azs.append(alt_az[1, 0])
    jones = S.short_dipole_jones(np.array(azs), np.pi / 2 - np.array(alts))
    jones = np.flip(np.flip(np.transpose(jones, (1, 0, 2)), axis=1), axis=0)  # test_skymodel.py:1456-1462
    inst = np.einsum("ab...,bc...,dc...->ad...", jones, out, np.conj(jones))
    expect = np.array(pv["expected_re"]) + 1j * np.array(pv["expected_im"])
    assert np.allclose(inst, expect)  # the reference's own tolerance (rtol 1e-5, atol 1e-8)
    assert np.max(np.abs(inst - expect) / np.abs(expect)) < 5e-6


def test_tutorial_local_coherency(known_answers, hera):
    lon, lat, h, jd = hera
    tk = known_answers["tutorial_local_coherency"]
    a = E.astrom_context(jd, 0.0, lon, lat, h, dut1=-0.47, yp=0.32 * E.DAS2R)
    ra, dec = E.altaz_to_icrs(np.deg2rad(tk["alt_deg"]), np.deg2rad(tk["az_deg"]), a)
    alt_az, _, above = S.update_positions(ra, dec, a)
    ravg = S.calc_average_rotation_matrix(a)
    c = S.coherency_calc(np.array(tk["stokes"]).reshape(4, 1, 1), ra, dec, alt_az, above, ravg)[:, :, 0, 0]
    expect = np.array(tk["coherency_re"])
    assert np.allclose(np.diag(c.real), np.diag(expect), atol=2e-9)
    assert abs(c[0, 1].real / expect[0, 1] - 1) < 0.01  # IERS-limited, see the fixture's note
    assert np.abs(c.imag).max() == 0


@pytest.mark.filterwarnings("ignore:Horizon cutoff undefined")
def test_pol_rotator(hera):
    # tests/test_skymodel.py:1306-1385
    lon, lat, h, jd = hera
    a = E.astrom_context(jd, 0.0, lon, lat, h)
    n = 50
    ra = np.deg2rad(np.linspace(0, 24, n) * 15) % (2 * np.pi)
    dec = np.deg2rad(np.linspace(-90, 90, n))
    st = np.zeros((4, 1, n)); st[:, 0, :] = np.array([5.5, 0.7, 0.3, 0.0])[:, None]
    st[:, 0, -1] = [1.0, 0, 0, 0]
    assert S.polarized_indices(st).size == n - 1
    alt_az, _, above = S.update_positions(ra, dec, a)
    ravg = S.calc_average_rotation_matrix(a)
    rots1 = S.calc_rotation_matrix(ra, dec, alt_az, ravg)
    inds = np.array([25, 45, 16])
    assert np.allclose(rots1[..., inds], S.calc_rotation_matrix(ra, dec, alt_az, ravg, inds))
    frame = S.stokes_to_coherency(st)
    with pytest.warns(UserWarning, match="Horizon cutoff undefined"):
        local = S.coherency_calc(st, ra, dec, alt_az, None, ravg)
    assert not np.all(np.isclose(local[..., :-1], frame[..., :-1]))
    assert np.allclose(local[..., -1], frame[..., -1])
    assert S.coherency_calc(st, ra, dec, alt_az, np.zeros(n, bool), ravg).size == 0


def test_rotation_closed_form(hera):
    """R^T C R of skymodel.py:2704-2710 == (I', Q', U', V') = (n2 I, c2 Q - s2 U, s2 Q + c2 U, n2 V)."""
    lon, lat, h, jd = hera
    a = E.astrom_context(jd, 0.3, lon, lat, h)
    rng = np.random.default_rng(11)
    n = 64
    ra = rng.uniform(0, 2 * np.pi, n); dec = np.arcsin(rng.uniform(-1, 1, n))
    st = rng.normal(size=(4, 3, n)); st[0] = np.abs(st[0]) + 2
    alt_az, _, above = S.update_positions(ra, dec, a)
    ravg = S.calc_average_rotation_matrix(a)
    ref = S.coherency_calc(st, ra, dec, alt_az, above, ravg)
    rot = S.calc_coherency_rotation(ra, dec, alt_az, ravg, np.nonzero(above)[0])
    cx, sx = rot[0, 0], rot[0, 1]
    n2, c2, s2 = cx * cx + sx * sx, cx * cx - sx * sx, 2 * sx * cx
    i, q, u, v = (st[k][:, above] for k in range(4))
    closed = S.stokes_to_coherency(np.array([n2 * i, c2 * q - s2 * u, s2 * q + c2 * u, n2 * v]))
    assert np.abs(closed - ref).max() < 1e-14


# ---- HEALPix ---------------------------------------------------------------------------------------
def test_healpix_known_pixel(known_answers):
    ka = known_answers["healpix_pixel"]
    z, _, phi = hp.pix2zphi(ka["nside"], np.array([ka["ipix"]]))
    assert np.isclose(z[0], ka["z"]) and np.isclose(phi[0], ka["phi"])


@pytest.mark.parametrize("nside", [1, 2, 4, 8, 16, 64])
def test_healpix_invariants(nside):
    npix = 12 * nside * nside
    lon_r, lat_r = hp.healpix_to_lonlat(nside, np.arange(npix))
    lon_n, lat_n = hp.healpix_to_lonlat(nside, np.arange(npix), nested=True)
    assert np.all(np.diff(lat_r) <= 1e-15)  # ring order runs north -> south
    ring = sorted(zip(np.round(lat_r, 12), np.round(lon_r, 12)))
    nest = sorted(zip(np.round(lat_n, 12), np.round(lon_n, 12)))
    assert np.allclose(ring, nest)
    vec = np.array([np.cos(lat_r) * np.cos(lon_r), np.cos(lat_r) * np.sin(lon_r), np.sin(lat_r)])
    assert np.abs(vec.sum(axis=1)).max() < 1e-12 * npix
    assert lon_r.min() >= 0 and lon_r.max() < 2 * np.pi
    # nested: the 4 children of a coarse pixel surround its centre
    if nside >= 2:
        lon_c, lat_c = hp.healpix_to_lonlat(nside // 2, np.arange(npix // 4), nested=True)
        kids = np.array([np.cos(lat_n) * np.cos(lon_n), np.cos(lat_n) * np.sin(lon_n), np.sin(lat_n)]).reshape(3, -1, 4).mean(axis=2)
        par = np.array([np.cos(lat_c) * np.cos(lon_c), np.cos(lat_c) * np.sin(lon_c), np.sin(lat_c)])
        assert np.all(np.sum(kids * par, axis=0) / np.linalg.norm(kids, axis=0) > np.cos(0.7 / nside))


def test_interp1d_is_linear_operator():
    """The fact the GPU subband path rests on: interp1d is linear in the data."""
    x = np.linspace(70e6, 230e6, 20)
    rng = np.random.default_rng(5)
    y = rng.normal(size=(20, 6))
    f = np.linspace(80e6, 220e6, 333)
    w = scipy.interpolate.interp1d(x, np.eye(20), axis=0, kind="cubic")(f)
    assert np.abs(w @ y - scipy.interpolate.interp1d(x, y, axis=0, kind="cubic")(f)).max() < 1e-13


def test_lonlat_to_healpix_inverts_pixel_centres():
    """Size-independent property of the ang2pix restatement: it inverts pix2ang on every pixel, ring and
    nested, and the two orderings agree on random directions."""
    from oracle import healpix_np as hp

    for nside in (1, 2, 8, 32, 128):
        pix = np.arange(12 * nside * nside)
        for nested in (False, True):
            lon, lat = hp.healpix_to_lonlat(nside, pix, nested=nested)
            assert np.array_equal(hp.lonlat_to_healpix(nside, lon, lat, nested=nested), pix)
    for nside in (3, 12, 100):  # ring ordering has no power-of-two restriction
        pix = np.arange(12 * nside * nside)
        lon, lat = hp.healpix_to_lonlat(nside, pix)
        assert np.array_equal(hp.lonlat_to_healpix(nside, lon, lat), pix)
    rng = np.random.default_rng(8)
    lon, lat = rng.uniform(0, 2 * np.pi, 20000), np.arcsin(rng.uniform(-1, 1, 20000))
    cr = hp.healpix_to_lonlat(64, hp.lonlat_to_healpix(64, lon, lat))
    cn = hp.healpix_to_lonlat(64, hp.lonlat_to_healpix(64, lon, lat, nested=True), nested=True)
    assert np.allclose(cr[0], cn[0]) and np.allclose(cr[1], cn[1])
    # the reference's fixture (tests/test_skymodel.py:234-244): +-0.01 deg around the centre of ring pixel 25
    # at nside 32 stays in pixel 25
    ra, dec = hp.healpix_to_lonlat(32, np.array([25]))
    off = np.deg2rad(0.01)
    assert list(hp.lonlat_to_healpix(32, ra + np.array([-off, off]), dec + np.array([-off, off]))) == [25, 25]


# ---------------------------------------------------------------------------------------------------
# ERFA's own known-answer vectors (src/t_erfa_c.c): the anchor of the position chain outside this repo
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def erfa_kat():
    import json
    import os

    with open(os.path.join(os.path.dirname(__file__), "golden", "erfa_check_values.json")) as fh:
        return json.load(fh)


def _astrom_from(star, obs=None):
    """eraASTROM (31 doubles) from the star-independent part of one check record and the observer part of another."""
    a = np.zeros(E.ASTROM_LEN)
    a[E.I_PMT] = star["pmt"]
    a[E.I_EB:E.I_EB + 3] = star["eb"]
    a[E.I_EH:E.I_EH + 3] = star["eh"]
    a[E.I_EM] = star["em"]
    a[E.I_V:E.I_V + 3] = star["v"]
    a[E.I_BM1] = star["bm1"]
    a[E.I_BPN:E.I_BPN + 9] = np.asarray(star["bpn"]).ravel()
    if obs is not None:
        for key, slot in (("along", E.I_ALONG), ("xpl", E.I_XPL), ("ypl", E.I_YPL), ("sphi", E.I_SPHI), ("cphi", E.I_CPHI),
                          ("diurab", E.I_DIURAB), ("eral", E.I_ERAL), ("refa", E.I_REFA), ("refb", E.I_REFB)):
            a[slot] = obs[key]
        a[E.I_PHI] = obs.get("phi", 0.0)
    return a


def test_erfa_check_values(erfa_kat):
    """oracle/erfa_np.py against ERFA's published test vectors: eraAb, eraLd, eraAtciqz (on the eraApci13 context),
    eraAtioq (on the eraApio13 context: refraction and diurnal aberration present), eraApcs, eraApco, eraPvtob,
    eraGd2gc, eraC2ixys, eraEra00, eraSp00.  Tolerances are t_erfa_c.c's own (1e-12) or tighter."""
    k = erfa_kat
    ab = k["ab"]
    got = E.ab(np.asarray(ab["pnat"])[:, None], ab["v"], ab["s"], ab["bm1"])[:, 0]
    assert np.abs(got - ab["ppr"]).max() < 1e-15
    ld = k["ld"]
    got = E.ld(ld["bm"], np.asarray(ld["p"])[:, None], np.asarray(ld["q"])[:, None], ld["e"], ld["em"], ld["dlim"])[:, 0]
    assert np.abs(got - ld["p1"]).max() < 1e-15
    # apci13's context is internally consistent (orthogonal bpn, unit eh, bm1 = sqrt(1 - v^2)) ...
    star = k["apci13"]
    bpn = np.asarray(star["bpn"])
    assert np.abs(bpn @ bpn.T - np.eye(3)).max() < 1e-15
    assert abs(np.linalg.norm(star["eh"]) - 1.0) < 1e-15
    assert abs(np.sqrt(1.0 - np.sum(np.square(star["v"]))) - star["bm1"]) < 1e-15
    # ... and carries (2.71, 0.174) to ERFA's CIRS answer through light deflection, aberration and bpn
    q = k["atciqz"]
    ri, di = E.atciqz(np.array([q["rc"]]), np.array([q["dc"]]), _astrom_from(star))
    assert abs(ri[0] - q["ri"]) < 1e-15 and abs(di[0] - q["di"]) < 1e-15
    # CIRS -> observed with polar motion, diurnal aberration and A/B refraction
    o = k["atioq"]
    obs = np.zeros(E.ASTROM_LEN)
    for key, slot in (("along", E.I_ALONG), ("xpl", E.I_XPL), ("ypl", E.I_YPL), ("sphi", E.I_SPHI), ("cphi", E.I_CPHI),
                      ("diurab", E.I_DIURAB), ("eral", E.I_ERAL), ("refa", E.I_REFA), ("refb", E.I_REFB)):
        obs[slot] = k["apio13"][key]
    az, zd = E.atioq(np.array([o["ri"]]), np.array([o["di"]]), obs)
    assert abs(az[0] - o["aob"]) < 1e-15 and abs(zd[0] - o["zob"]) < 1e-15
    # the star-independent parameters themselves
    c = k["apcs"]
    a = E.apcs(*c["date"], np.asarray(c["pv"]), np.asarray(c["ebpv"]), np.asarray(c["ehp"]))
    assert abs(a[E.I_PMT] - c["pmt"]) < 1e-11 and abs(a[E.I_EM] - c["em"]) < 1e-15 and abs(a[E.I_BM1] - c["bm1"]) < 1e-15
    assert np.abs(a[E.I_EB:E.I_EB + 3] - c["eb"]).max() < 1e-15 and np.abs(a[E.I_EH:E.I_EH + 3] - c["eh"]).max() < 1e-15
    assert np.abs(a[E.I_V:E.I_V + 3] - c["v"]).max() < 1e-19  # (with AULT = 499.004782 this is 3.3e-13)
    c = k["apco"]
    a = E.apco(*c["date"], np.asarray(k["apcs"]["ebpv"]), np.asarray(k["apcs"]["ehp"]), c["x"], c["y"], c["s"], c["theta"],
               c["elong"], c["phi"], c["hm"], c["xp"], c["yp"], c["sp"], c["refa"], c["refb"])
    want = _astrom_from(c, c)
    want[E.I_PHI] = c["phi"]
    assert np.abs(a - want).max() < 1e-15
    c = k["pvtob"]
    p, v = E.pvtob(c["elong"], c["phi"], c["hm"], c["xp"], c["yp"], c["sp"], c["theta"])
    assert np.abs(p - c["p"]).max() < 1e-5 and np.abs(v - c["v"]).max() < 1e-9  # t_erfa_c: 1e-5 m, 1e-9 m/s
    c = k["gd2gc_wgs84"]
    assert np.abs(E.gd2gc_wgs84(c["elong"], c["phi"], c["height"]) - c["xyz"]).max() < 1e-7
    c = k["c2ixys"]
    assert np.abs(E.c2ixys(c["x"], c["y"], c["s"]) - c["rc2i"]).max() < 1e-15
    assert abs(E.era00(*k["era00"]["date"]) - k["era00"]["era"]) < 1e-12
    assert abs(E.sp00(*k["sp00"]["date"]) - k["sp00"]["sp"]) < 1e-24


def test_apio13_observer_part_from_oracle_builder(erfa_kat):
    """The Earth-rotation / site half of a context (eral, along, xpl, ypl, sphi, cphi) needs no series: eraApio13's
    published values follow from era00, sp00 and the polar-motion composition the oracle's apco already uses."""
    c = erfa_kat["apio13"]
    utc1, utc2 = c["utc"]
    tai = utc2 + E.tai_minus_utc(utc1 + utc2) / E.DAYSEC
    tt = tai + 32.184 / E.DAYSEC
    ut1 = utc2 + c["dut1"] / E.DAYSEC
    sp = E.sp00(utc1, tt)
    theta = E.era00(utc1, ut1)
    dummy = np.array([[1.0, 0, 0], [0, 0.01, 0]])
    a = E.apco(utc1, tt, dummy, np.array([1.0, 0, 0]), 0.0, 0.0, 0.0, theta, c["elong"], c["phi"], c["hm"], c["xp"],
               c["yp"], sp, c["refa"], c["refb"])
    for key, slot in (("along", E.I_ALONG), ("xpl", E.I_XPL), ("ypl", E.I_YPL), ("sphi", E.I_SPHI), ("cphi", E.I_CPHI),
                      ("eral", E.I_ERAL)):
        assert abs(a[slot] - c[key]) < 1e-12, key


def test_aberration_and_deflection_magnitudes(hera):
    """SURVEY Appendix A invariant (ii), independent of any recalled number: annual aberration displaces a source 90 deg
    from the Earth's velocity by v/c ~ 20.5 arcsec towards the apex, and the Sun deflects light from a source 90 deg
    away from it by 4.07 mas, radially away from the Sun."""
    lon, lat, h, jd = hera
    a = E.astrom_context(jd, 0.3, lon, lat, h)
    v = a[E.I_V:E.I_V + 3]
    vhat = v / np.linalg.norm(v)
    perp = np.cross(vhat, [0.0, 0.0, 1.0]); perp /= np.linalg.norm(perp)
    out = E.ab(perp[:, None], v, a[E.I_EM], a[E.I_BM1])[:, 0]
    shift = np.arccos(np.clip(out @ perp, -1, 1))
    assert abs(shift - np.linalg.norm(v)) < 1e-9 and 20.0 < np.rad2deg(shift) * 3600 < 21.0
    assert out @ vhat > 0  # towards the apex
    eh = a[E.I_EH:E.I_EH + 3]  # Sun -> observer
    p = np.cross(eh, [0.0, 0.0, 1.0]); p /= np.linalg.norm(p)  # 90 deg from the Sun
    d = E.ld(1.0, p[:, None], p[:, None], eh, a[E.I_EM], 1e-6)[:, 0]
    defl = np.linalg.norm(d - p)
    assert abs(np.rad2deg(defl) * 3600e3 - 4.07 / a[E.I_EM]) < 0.02  # 2 SRS / em at 90 deg = 4.07 mas at 1 au
    assert (d - p) @ eh > 0  # away from the Sun (eh points from the Sun to the observer)
