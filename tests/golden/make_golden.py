"""Generate the golden fixtures under tests/golden/.

Run in the build container only (it needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Two files are produced and committed:

* ``sct_golden.json`` -- outputs of the REFERENCE's own ``spherical_coords_transforms.py`` (the one
  reference module that imports without astropy; loaded by file path) on seeded random inputs.
* ``reference_known_answers.json`` -- known-answer values transcribed from the reference's tests and
  doctests, each with the file:line it comes from.  astropy/pyuvdata are absent in this image, so the
  reference's SkyModel itself cannot be run to produce more.
"""

import importlib.util
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/pyradiosky/spherical_coords_transforms.py"


def _load_ref_sct():
    spec = importlib.util.spec_from_file_location("ref_sct", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_sct_golden():
    sct = _load_ref_sct()
    rng = np.random.default_rng(20261019)
    n = 64
    theta = np.arccos(rng.uniform(-1, 1, n))
    phi = rng.uniform(0, 2 * np.pi, n)
    beta = np.arccos(rng.uniform(-1, 1, n))
    alpha = rng.uniform(0, 2 * np.pi, n)
    # a rotation near the identity composed with a random rotation, like R_perturb . R_avg
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] *= -1
    cases = []
    for i in range(n):
        r1 = q @ sct.r_hat(theta[i], phi[i])
        # second vector ~1e-4 rad away (the aberration scale) for half the cases, far for the rest
        if i % 2 == 0:
            d = rng.normal(size=3) * 1e-4
            r2 = r1 + d
            r2 = r2 / np.sqrt(np.dot(r2, r2))
        else:
            r2 = sct.r_hat(beta[i], alpha[i])
        r1 = r1 / np.sqrt(np.dot(r1, r1))
        rot = sct.vecs2rot(r1=r1, r2=r2)
        r_exact = rot @ q
        # angles of r2 in the (beta, alpha) chart
        b = float(np.arccos(np.clip(r2[2], -1, 1)))
        a = float(np.arctan2(r2[1], r2[0]) % (2 * np.pi))
        basis = sct.spherical_basis_vector_rotation_matrix(theta[i], phi[i], r_exact, b, a)
        cases.append(
            {
                "theta": float(theta[i]),
                "phi": float(phi[i]),
                "beta": b,
                "alpha": a,
                "r1": r1.tolist(),
                "r2": r2.tolist(),
                "vecs2rot": rot.tolist(),
                "r_exact": r_exact.tolist(),
                "basis_rot": basis.tolist(),
            }
        )
    out = {
        "source": "reference src/pyradiosky/spherical_coords_transforms.py imported by file path",
        "q": q.tolist(),
        "theta": theta.tolist(),
        "phi": phi.tolist(),
        "r_hat": sct.r_hat(theta, phi).tolist(),
        "theta_hat": sct.theta_hat(theta, phi).tolist(),
        "phi_hat": sct.phi_hat(theta, phi).tolist(),
        "axis_angle": {
            "axis": [0.6, 0.0, 0.8],
            "angle": 0.3,
            "matrix": sct.axis_angle_rotation_matrix(np.array([0.6, 0.0, 0.8]), 0.3).tolist(),
        },
        "rotate_points_3d": [
            list(map(float, sct.rotate_points_3d(q, float(theta[i]), float(phi[i])))) for i in range(8)
        ],
        "cases": cases,
    }
    with open(os.path.join(HERE, "sct_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)


def make_known_answers():
    ka = {
        "hera_site": {
            "cite": "tests/test_skymodel.py:37-42",
            "lat_dms": [-30, 43, 17.5],
            "lon_dms": [21, 25, 41.9],
            "height_m": 1073.0,
            "utc": "2015-03-01 00:00:00",
            "jd_utc": 2457082.5,
        },
        "hera_geocentric_m": {
            "cite": "docs/tutorial.rst:1231-1233",
            "xyz": [5109342.76, 2005241.904, -3239939.469],
        },
        "stokes_coherency": {
            "cite": "tests/test_utils.py:14-23",
            "stokes": [4.5, -0.3, 1.2, -0.15],
            "coherency_re": [[2.1, 0.6], [0.6, 2.4]],
            "coherency_im": [[0.0, 0.075], [-0.075, 0.0]],
        },
        "tutorial_frame_coherency": {
            "cite": "docs/tutorial.rst:1165-1177",
            "stokes": [1.0, 0.2, 0.0, 0.0],
            "coherency_re": [[0.6, 0.0], [0.0, 0.4]],
        },
        "tutorial_local_coherency": {
            "cite": "docs/tutorial.rst:1153-1185 (alt 80 deg, az 0 at HERA, 2015-03-01 00:00 UTC)",
            "stokes": [1.0, 0.2, 0.0, 0.0],
            "alt_deg": 80.0,
            "az_deg": 0.0,
            "coherency_re": [[5.99999999e-01, -1.20482720e-05], [-1.20482720e-05, 4.00000001e-01]],
            "note": "off-diagonal is sensitive at the percent level to IERS polar motion / UT1-UTC",
        },
        "tutorial_altaz_lmn": {
            "cite": "docs/tutorial.rst:1204-1246 (sources at alt [90, 80], az [0, 10] deg)",
            "alt_az_2dp": [[1.57, 1.4], [1.73, 0.17]],
            "pos_lmn_1sf": [[2.1e-13, 3.0e-02], [-3.4e-14, 1.7e-01], [1.0, 9.8e-01]],
        },
        "tutorial_horizon_mask": {
            "cite": "docs/tutorial.rst:1086-1132",
            "site_lat_deg": -26.7033194,
            "site_lon_deg": 116.67081524,
            "site_height_m": 377.83,
            "jd_utc": 2457086.5,
            "ra_center_deg_is_hera_lon": True,
            "above_horizon": [
                True, True, True, True, True, True, True, True, False, False, True, True,
                False, False, False, True, True, False, False, False, False, False, False, False,
                False,
            ],
        },
        "polarized_visibilities": {
            "cite": "tests/test_skymodel.py:1388-1508 (rtol 1e-5 via units.quantity.allclose)",
            "stokes": [1, -0.2, 0.3, 0.1],
            "time_offsets_hr": [-1 / 6.0, -1 / 6.0 + 0.1, 0.0, -1 / 6.0 + 0.3, -1 / 6.0 + 0.4],
            "expected_re": [
                [
                    [0.60572311, 0.60250361, 0.5999734, 0.59400581, 0.58875092],
                    [0.14530468, 0.14818987, 0.15001773, 0.15342311, 0.15574023],
                ],
                [
                    [0.14530468, 0.14818987, 0.15001773, 0.15342311, 0.15574023],
                    [0.39342384, 0.39736029, 0.4000266, 0.40545359, 0.40960028],
                ],
            ],
            "expected_im": [
                [
                    [0.0, 0.0, 0.0, 0.0, 0.0],
                    [4.99646383e-02, 4.99943414e-02, 5.00000000e-02, 4.99773672e-02, 4.99307016e-02],
                ],
                [
                    [-4.99646383e-02, -4.99943414e-02, -5.00000000e-02, -4.99773672e-02, -4.99307016e-02],
                    [0.0, 0.0, 0.0, 0.0, 0.0],
                ],
            ],
        },
        "jy_to_ksr_cgs": {
            "cite": "tests/test_utils.py:139-152 (np.allclose, default rtol 1e-5)",
            "c_cmps": 29979245800.0,
            "k_boltz_erg": 1.380658e-16,
            "freqs_mhz": [100.0, 200.0, 200],
        },
        "tutorial_unit_conversion": {
            "cite": "docs/tutorial.rst:1279-1297",
            "gleam_jy": [0.528997, -0.032702, 0.463359, 2.686571, 0.393777],
            "gleam_ksr": [0.00298095, -0.00018428, 0.00261107, 0.01513907, 0.00221897],
            "gleam_freq_hz": 76e6,
            "gsm_k": [4704.91299386, 3864.90157423, 3933.76949248, 4258.30083558, 6520.16612935],
            "gsm_jy_sr": [361379.47094723, 296859.06795353, 302148.74108755, 327075.65583124,
                          500807.17526256],
            "gsm_freq_hz": 50e6,
        },
        "at_frequencies_power_law": {
            "cite": "tests/test_skymodel.py:162-221 and :3542-3583 (10 comps, 30 freqs 100-130 MHz, alpha=-0.5)",
            "alpha": -0.5,
            "ncomp": 10,
            "freqs_in": [100e6, 130e6, 30],
            "freqs_fine": [100e6, 130e6, 50],
        },
        "healpix_pixel": {
            "cite": "tests/test_skymodel.py:2094 (ipix 357, nside 8, ring)",
            "nside": 8,
            "ipix": 357,
            "z": 1.0 / 12.0,
            "phi": 4.123340357836604,
        },
    }
    with open(os.path.join(HERE, "reference_known_answers.json"), "w") as fh:
        json.dump(ka, fh, indent=1)


def copy_skyh5_fixtures():
    """Three of the reference's own (small) SkyH5 test *data* fixtures, byte for byte (under names that say where
    they come from): files written by h5py / libhdf5 that the product's HDF5 reader (include/rsk_io.h) must read
    without libhdf5 -- the only h5py-written bytes available in an image that has neither h5py nor libhdf5.

    * ``gsm_icrs.skyh5`` (768-pixel HEALPix map, nside 8, 10 channels, K; chunked Stokes, old-style
      ``lon``/``lat``/``frame`` header) -- the tutorial prints its first five Stokes I values
      (docs/tutorial.rst:1288-1297, transcribed into ``skyh5_known_answers.json``).
    * ``old_skyh5_point_sources.skyh5`` (50 GLEAM sources x 20 subbands, ``ra``/``dec`` in degrees) -- read by
      tests/test_skymodel.py:4507-4522, which fixes the warnings and the subband -> full downgrade.
    * ``gsm_galactic.skyh5`` (the same map in the galactic frame): ``healpix_interp_transform("icrs")`` of it must
      equal ``gsm_icrs.skyh5`` (tests/test_skymodel.py:4423-4452) -- the golden pair that pins the bilinear
      HEALPix interpolation and the galactic -> ICRS rotation.
    """
    import shutil

    src = "/root/reference/src/pyradiosky/data"
    for name, fixture in (("gsm_icrs.skyh5", "ref_fixture_gsm_icrs_nside8.skyh5"),
                          ("old_skyh5_point_sources.skyh5", "ref_fixture_gleam50_old_header.skyh5"),
                          ("gsm_galactic.skyh5", "ref_fixture_gsm_galactic_nside8.skyh5"),
                          # VOTable test data (TABLEDATA text): the 50-source GLEAM extract every GLEAM test of the
                          # reference reads (tests/test_skymodel.py:2879-3077), its two-row file and its ID-less one
                          ("gleam_50srcs.vot", "ref_fixture_gleam_50srcs.vot"),
                          ("simple_test.vot", "ref_fixture_simple_test.vot"),
                          ("single_source_old.vot", "ref_fixture_single_source_old.vot"),
                          # text catalogues read by tests/test_skymodel.py:3214-3238 and :3436-3441
                          ("pointsource_catalog.txt", "ref_fixture_pointsource_catalog.txt"),
                          ("mock_hera_text_2458098.27471.txt", "ref_fixture_mock_hera_text.txt"),
                          # FHD IDL save files read by tests/test_skymodel.py:3079-3211 (the smaller four)
                          ("extended_source_test.sav", "ref_fixture_extended_source_test.sav"),
                          ("fhd_catalog_no_extend.sav", "ref_fixture_fhd_catalog_no_extend.sav"),
                          ("fhd_catalog_bad.sav", "ref_fixture_fhd_catalog_bad.sav"),
                          ("fhd_source_array.sav", "ref_fixture_fhd_source_array.sav")):
        shutil.copyfile(os.path.join(src, name), os.path.join(HERE, fixture))
    ka = {
        "gsm_icrs": {
            "cite": "docs/tutorial.rst:1288-1297",
            "Ncomponents": 768,
            "stokes_I_first5_K": [4704.91299386, 3864.90157423, 3933.76949248, 4258.30083558, 6520.16612935],
            "stokes_I_first5_Jy_per_sr": [361379.47094723, 296859.06795353, 302148.74108755, 327075.65583124,
                                          500807.17526256],
        },
        "old_skyh5_point_sources": {
            "cite": "tests/test_skymodel.py:4507-4522; docs/tutorial.rst:1279-1286 (same 50 GLEAM sources)",
            "warnings": ["Parameter skycoord not found in skyh5 file.",
                         "No freq_edge_array in this file and frequencies are not evenly spaced"],
            "spectral_type_after_read": "full",
            "stokes_I_first5_Jy": [0.528997, -0.032702, 0.463359, 2.686571, 0.393777],
            "stokes_I_first5_K_sr": [0.00298095, -0.00018428, 0.00261107, 0.01513907, 0.00221897],
        },
    }
    with open(os.path.join(HERE, "skyh5_known_answers.json"), "w") as fh:
        json.dump(ka, fh, indent=1)


if __name__ == "__main__":
    make_sct_golden()
    make_known_answers()
    copy_skyh5_fixtures()
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".json")))
