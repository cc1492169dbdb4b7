"""Oracle (test infrastructure): numpy/scipy restatement of the SkyModel hot path.

Mirrors, on plain float64 arrays (radians, Hz, Jy or K), these reference functions
(paths relative to the upstream tree):

  src/pyradiosky/skymodel.py:1037-1042   _polarized / _n_polarized
  src/pyradiosky/skymodel.py:1593-1635   jansky_to_kelvin (and :1541-1591 kelvin_to_jansky)
  src/pyradiosky/skymodel.py:1646-1673   get_lon_lat (healpix branch -> oracle.healpix_np)
  src/pyradiosky/skymodel.py:2084-2389   at_frequencies
  src/pyradiosky/skymodel.py:2413-2479   update_positions (astropy part -> oracle.erfa_np)
  src/pyradiosky/skymodel.py:2481-2497   calc_frame_coherency
  src/pyradiosky/skymodel.py:2499-2555   _calc_average_rotation_matrix
  src/pyradiosky/skymodel.py:2557-2601   _calc_rotation_matrix      (per-source Python loop kept)
  src/pyradiosky/skymodel.py:2603-2646   _calc_coherency_rotation   (per-source Python loop kept)
  src/pyradiosky/skymodel.py:2648-2712   coherency_calc
  src/pyradiosky/utils.py:17-61          stokes_to_coherency
  src/pyradiosky/utils.py:105-126        jy_to_ksr

The ``*_vec`` variants are vectorised re-expressions of the two per-source loops; they are used only
to make larger parity cases finish in seconds and are themselves tested against the loop versions.
"""

import warnings

import numpy as np
import scipy.interpolate
from scipy.linalg import orthogonal_procrustes as ortho_procr

from . import erfa_np, healpix_np
from . import sct_np as sct

C_MPS = 299792458.0
K_B = 1.380649e-23


def polarized_indices(stokes):
    """skymodel.py:1037-1042 -- NB the predicate is sum(Q,U,V) != 0 at any stored frequency."""
    return np.where(np.any(np.sum(stokes[1:, :, :], axis=0) != 0.0, axis=0))[0]


def jy_to_ksr(freqs_hz):
    """utils.py:105-126: astropy ``brightness_temperature(freqs, beam_area=1 sr)`` factor [K sr/Jy]."""
    freqs_hz = np.atleast_1d(np.asarray(freqs_hz, dtype=float))
    return 1e-26 * C_MPS**2 / (2.0 * K_B * freqs_hz**2)


def stokes_to_coherency(stokes_arr):
    """utils.py:17-61: C = 0.5 [[I+Q, U-iV], [U+iV, I-Q]]."""
    stokes_arr = np.asarray(stokes_arr)
    initial_shape = stokes_arr.shape
    if initial_shape[0] != 4:
        raise ValueError("First dimension of stokes_vector must be length 4.")
    if stokes_arr.size == 4 and len(initial_shape) == 1:
        stokes_arr = stokes_arr[:, np.newaxis, np.newaxis]
    coherency = 0.5 * np.array(
        [
            [stokes_arr[0] + stokes_arr[1], stokes_arr[2] - 1j * stokes_arr[3]],
            [stokes_arr[2] + 1j * stokes_arr[3], stokes_arr[0] - stokes_arr[1]],
        ]
    )
    if stokes_arr.size == 4 and len(initial_shape) == 1:
        coherency = coherency.squeeze()
    return coherency


def get_lon_lat_healpix(nside, hpx_inds, hpx_order="ring"):
    """skymodel.py:1656-1669."""
    return healpix_np.healpix_to_lonlat(nside, hpx_inds, nested=(hpx_order == "nested"))


def update_positions(lon, lat, astrom, frame_to_icrs=None):
    """skymodel.py:2446-2479: returns alt_az (2,N), pos_lmn (3,N), above_horizon (N,)."""
    alt, az = erfa_np.icrs_to_altaz(lon, lat, astrom, frame_to_icrs)
    alt_az = np.array([alt, az])
    pos_lmn = np.zeros((3, alt_az.shape[1]), dtype=float)
    pos_lmn[0, :] = np.sin(alt_az[1, :]) * np.cos(alt_az[0, :])
    pos_lmn[1, :] = np.cos(alt_az[1, :]) * np.cos(alt_az[0, :])
    pos_lmn[2, :] = np.sin(alt_az[0, :])
    above_horizon = alt_az[0, :] > 0.0
    return alt_az, pos_lmn, above_horizon


def calc_average_rotation_matrix(astrom):
    """skymodel.py:2499-2555: polar (orthogonal) factor of the AltAz images of the ICRS axes.

    The three ICRS unit axes go through the same astropy transform as the sources (:2518-2527;
    dimensionless cartesian data are treated as unit-spherical by astropy).
    """
    lon = np.array([0.0, np.pi / 2.0, 0.0])
    lat = np.array([0.0, 0.0, np.pi / 2.0])
    alt, az = erfa_np.icrs_to_altaz(lon, lat, astrom)
    # astropy AltAz cartesian: x = cos(alt) cos(az), y = cos(alt) sin(az), z = sin(alt)
    r_screwy = np.array([np.cos(alt) * np.cos(az), np.cos(alt) * np.sin(az), np.sin(alt)])
    r_really_orthogonal, _ = ortho_procr(r_screwy, np.eye(3))
    return np.array(r_really_orthogonal).T


def calc_rotation_matrix(lon, lat, alt_az, r_avg, inds=None):
    """skymodel.py:2557-2601, per-source loop as upstream."""
    if inds is None:
        inds = range(len(lon))
    inds = np.asarray(list(inds), dtype=int)
    n_inds = len(inds)
    theta_frame = np.pi / 2.0 - lat[inds]
    phi_frame = lon[inds]
    frame_vec = sct.r_hat(theta_frame, phi_frame)
    theta_altaz = np.pi / 2.0 - alt_az[0, inds]
    phi_altaz = alt_az[1, inds]
    altaz_vec = sct.r_hat(theta_altaz, phi_altaz)
    r_exact = np.zeros((3, 3, n_inds), dtype=np.float64)
    for src_i in range(n_inds):
        intermediate_vec = np.matmul(r_avg, frame_vec[:, src_i])
        r_perturb = sct.vecs2rot(r1=intermediate_vec, r2=altaz_vec[:, src_i])
        r_exact[:, :, src_i] = np.matmul(r_perturb, r_avg)
    return r_exact


def calc_coherency_rotation(lon, lat, alt_az, r_avg, inds=None):
    """skymodel.py:2603-2646, per-source loop as upstream. Returns (2, 2, n)."""
    if inds is None:
        inds = range(len(lon))
    inds = np.asarray(list(inds), dtype=int)
    n_inds = len(inds)
    basis_rotation_matrix = calc_rotation_matrix(lon, lat, alt_az, r_avg, inds)
    theta_frame = np.pi / 2.0 - lat[inds]
    phi_frame = lon[inds]
    theta_altaz = np.pi / 2.0 - alt_az[0, inds]
    phi_altaz = alt_az[1, inds]
    coherency_rot_matrix = np.zeros((2, 2, n_inds), dtype=np.float64)
    for src_i in range(n_inds):
        coherency_rot_matrix[:, :, src_i] = sct.spherical_basis_vector_rotation_matrix(
            theta_frame[src_i],
            phi_frame[src_i],
            basis_rotation_matrix[:, :, src_i],
            theta_altaz[src_i],
            phi_altaz[src_i],
        )
    return coherency_rot_matrix


def calc_coherency_rotation_vec(lon, lat, alt_az, r_avg, inds=None):
    """Vectorised equivalent of :func:`calc_coherency_rotation` (same formulas, no Python loop)."""
    if inds is None:
        inds = np.arange(len(lon))
    inds = np.asarray(inds, dtype=int)
    theta_frame = np.pi / 2.0 - lat[inds]
    phi_frame = lon[inds]
    theta_altaz = np.pi / 2.0 - alt_az[0, inds]
    phi_altaz = alt_az[1, inds]
    r1 = r_avg @ sct.r_hat(theta_frame, phi_frame)
    r2 = sct.r_hat(theta_altaz, phi_altaz)
    norm = np.cross(r1, r2, axis=0)
    sin_psi = np.sqrt(np.sum(norm * norm, axis=0))
    n_hat = norm / sin_psi
    cos_psi = np.sum(r1 * r2, axis=0)
    psi = np.arctan2(sin_psi, cos_psi)
    zero = np.zeros_like(psi)
    k = np.array([[zero, -n_hat[2], n_hat[1]], [n_hat[2], zero, -n_hat[0]], [-n_hat[1], n_hat[0], zero]])
    kk = np.einsum("abn,bcn->acn", k, k)
    r_perturb = np.eye(3)[:, :, None] + np.sin(psi) * k + (1.0 - np.cos(psi)) * kk
    r_exact = np.einsum("abn,bc->acn", r_perturb, r_avg)
    th = sct.theta_hat(theta_frame, phi_frame)
    ph = sct.phi_hat(theta_frame, phi_frame)
    bh = np.einsum("ban,bn->an", r_exact, sct.theta_hat(theta_altaz, phi_altaz))
    cos_x = np.sum(bh * th, axis=0)
    sin_x = np.sum(bh * ph, axis=0)
    return np.array([[cos_x, sin_x], [-sin_x, cos_x]])


def coherency_calc(stokes, lon, lat, alt_az, above_horizon, r_avg, polarized=None, vectorized=False):
    """skymodel.py:2648-2712. ``stokes`` is (4, Nfreqs, N); returns (2, 2, Nfreqs, N_above)."""
    ncomp = stokes.shape[2]
    if above_horizon is None:
        warnings.warn(
            "Horizon cutoff undefined. Assuming all source components are above the horizon."
        )
        above_horizon = np.ones(ncomp).astype(bool)
    if polarized is None:
        polarized = polarized_indices(stokes)
    frame_coherency = stokes_to_coherency(stokes)
    coherency_local = frame_coherency[..., above_horizon]
    if np.unique(polarized).size > 0:
        pol_over_hor = np.isin(np.arange(ncomp)[above_horizon], polarized)
        full_pol_over_hor = [pi for pi in polarized if above_horizon[pi]]
        if len(pol_over_hor) > 0:
            rot = calc_coherency_rotation_vec if vectorized else calc_coherency_rotation
            rotation_matrix = rot(lon, lat, alt_az, r_avg, full_pol_over_hor)
            rotation_matrix_t = np.swapaxes(rotation_matrix, 0, 1)
            coherency_local[:, :, :, pol_over_hor] = np.einsum(
                "aby,bcxy,cdy->adxy",
                rotation_matrix_t,
                frame_coherency[:, :, :, full_pol_over_hor],
                rotation_matrix,
            )
    return coherency_local


def at_frequencies(
    stokes,
    spectral_type,
    freqs,
    freq_array=None,
    reference_frequency=None,
    spectral_index=None,
    freq_interp_kind="cubic",
    nan_handling="clip",
    atol=1.0,
):
    """skymodel.py:2084-2389. Returns the new (4, len(freqs), N) Stokes array (Hz everywhere)."""
    freqs = np.asarray(freqs, dtype=float)
    if spectral_type == "spectral_index":
        if np.any(np.isnan(spectral_index)):
            raise ValueError("Some spectral index values are NaNs.")
        return stokes * (freqs[:, None] / reference_frequency[None, :]) ** spectral_index[None, :]
    if spectral_type == "full":
        matches = np.fromiter(
            (np.isclose(freq, freqs, atol=atol).any() for freq in freq_array), dtype=bool
        )
        if np.sum(matches) != freqs.size:
            raise ValueError(
                "Some requested frequencies are not present in the current SkyModel."
            )
        return stokes[:, matches, :]
    if spectral_type == "subband":
        if np.max(freqs) > np.max(freq_array):
            raise ValueError("A requested frequency is larger than the highest subband frequency.")
        if np.min(freqs) < np.min(freq_array):
            raise ValueError("A requested frequency is smaller than the lowest subband frequency.")
        if not np.any(np.isnan(stokes)):
            finterp = scipy.interpolate.interp1d(freq_array, stokes, axis=1, kind=freq_interp_kind)
            return finterp(freqs)
        if nan_handling not in ["propagate", "interp", "clip"]:
            raise ValueError("nan_handling must be one of ['propagate', 'interp', 'clip']")
        warnings.warn("Some Stokes values are NaNs.")
        ncomp = stokes.shape[2]
        wh_nan = np.nonzero(np.any(np.isnan(stokes), axis=(0, 1)))[0]
        wh_non_nan = np.nonzero(np.all(~np.isnan(stokes), axis=(0, 1)))[0]
        new_stokes = np.zeros((4, freqs.size, ncomp), dtype=stokes.dtype)
        if wh_non_nan.size > 0:
            finterp = scipy.interpolate.interp1d(
                freq_array, stokes[:, :, wh_non_nan], axis=1, kind=freq_interp_kind
            )
            new_stokes[:, :, wh_non_nan] = finterp(freqs)
        if nan_handling == "propagate":
            new_stokes[:, :, wh_nan] = np.nan
            return new_stokes
        for comp in wh_nan:
            use = np.nonzero(np.all(~np.isnan(stokes[:, :, comp]), axis=0))[0]
            if use.size == 0:
                new_stokes[:, :, comp] = np.nan
                continue
            at_use = np.arange(freqs.size)
            if np.max(freqs) > np.max(freq_array[use]):
                at_use = np.nonzero(freqs <= np.max(freq_array[use]))[0]
                large = np.nonzero(freqs > np.max(freq_array[use]))[0]
                if nan_handling == "interp":
                    new_stokes[:, large, comp] = np.nan
                else:
                    new_stokes[:, large, comp] = stokes[:, np.full(large.size, use[-1]), comp]
            if np.min(freqs) < np.min(freq_array[use]):
                at_use_low = np.nonzero(freqs >= np.min(freq_array[use]))[0]
                at_use = np.intersect1d(at_use, at_use_low)
                small = np.nonzero(freqs < np.min(freq_array[use]))[0]
                if nan_handling == "interp":
                    new_stokes[:, small, comp] = np.nan
                else:
                    new_stokes[:, small, comp] = stokes[:, np.full(small.size, use[0]), comp]
            if at_use.size > 0:
                try:
                    finterp = scipy.interpolate.interp1d(
                        freq_array[use], stokes[:, use, comp], axis=1, kind=freq_interp_kind
                    )
                except ValueError:
                    finterp = scipy.interpolate.interp1d(
                        freq_array[use], stokes[:, use, comp], axis=1, kind="linear"
                    )
                new_stokes[:, at_use, comp] = finterp(freqs[at_use])
        return new_stokes
    # flat spectrum, skymodel.py:2369-2372
    return np.repeat(stokes, len(freqs), axis=1)


def jansky_to_kelvin_factor(spectral_type, nfreqs, ncomp, freq_array=None, reference_frequency=None):
    """skymodel.py:1609-1628: the full (4, Nfreqs, N) multiplicative factor Jy -> K sr."""
    if spectral_type == "spectral_index" or (
        spectral_type == "flat" and reference_frequency is not None
    ):
        conv = jy_to_ksr(reference_frequency)
        return np.repeat(np.repeat(conv[np.newaxis, np.newaxis, :], 4, axis=0), nfreqs, axis=1)
    if freq_array is not None:
        conv = jy_to_ksr(freq_array)
        return np.repeat(np.repeat(conv[np.newaxis, :, np.newaxis], 4, axis=0), ncomp, axis=2)
    raise ValueError("Either reference_frequency or freq_array must be set to convert to K.")


def jansky_to_kelvin(stokes, spectral_type, freq_array=None, reference_frequency=None):
    """skymodel.py:1593-1635 (unit bookkeeping left to the caller)."""
    fac = jansky_to_kelvin_factor(
        spectral_type, stokes.shape[1], stokes.shape[2], freq_array, reference_frequency
    )
    return stokes * fac


def kelvin_to_jansky(stokes, spectral_type, freq_array=None, reference_frequency=None):
    """skymodel.py:1541-1591."""
    fac = jansky_to_kelvin_factor(
        spectral_type, stokes.shape[1], stokes.shape[2], freq_array, reference_frequency
    )
    return stokes * (1 / fac)


def short_dipole_jones(az, za):
    """Analytic short-dipole E-field Jones J[vec in (az, za), feed in (e, n), N].

    Restates ``pyuvdata.ShortDipoleBeam.efield_eval`` (third-party, absent here) as used by the
    reference's golden test, test_skymodel.py:1446-1466.
    """
    az = np.asarray(az, dtype=float)
    za = np.asarray(za, dtype=float)
    j = np.zeros((2, 2, az.size))
    j[0, 0] = -np.sin(az)
    j[0, 1] = np.cos(az)
    j[1, 0] = np.cos(za) * np.cos(az)
    j[1, 1] = np.cos(za) * np.sin(az)
    return j
