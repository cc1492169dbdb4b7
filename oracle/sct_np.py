"""Oracle (test infrastructure): restatement of the reference's spherical-basis rotation helpers.

Follows ``src/pyradiosky/spherical_coords_transforms.py`` of the reference, keeping the order of
floating-point operations so results are bit-comparable with the original (checked in this container
by ``tests/golden/make_golden.py``, which imports the original by file path and stores its outputs in
``tests/golden/sct_golden.json``).  theta = co-latitude, phi = azimuth from +x.
"""

import numpy as np


def _pair(theta, phi):
    theta, phi = np.array(theta), np.array(phi)
    if theta.shape != phi.shape:
        raise ValueError("theta and phi must have the same shape")
    return theta, phi


def r_hat(theta, phi):
    """Radial unit vectors (3, N).  Reference spherical_coords_transforms.py:8-31."""
    theta, phi = _pair(theta, phi)
    st = np.sin(theta)
    return np.stack((np.cos(phi) * st, np.sin(phi) * st, np.cos(theta)))


def theta_hat(theta, phi):
    """Co-latitude unit vectors (3, N).  Reference spherical_coords_transforms.py:34-57."""
    theta, phi = _pair(theta, phi)
    ct = np.cos(theta)
    return np.stack((np.cos(phi) * ct, np.sin(phi) * ct, -np.sin(theta)))


def phi_hat(theta, phi):
    """Azimuthal unit vectors (3, N).  Reference spherical_coords_transforms.py:60-83."""
    theta, phi = _pair(theta, phi)
    return np.stack((-np.sin(phi), np.cos(phi), np.zeros_like(phi)))


def rotate_points_3d(rot_matrix, theta, phi):
    """(beta, alpha) of R.q for the point q=(theta, phi).  Reference :86-137 (scalar only)."""
    rot_matrix = np.array(rot_matrix)
    if rot_matrix.shape != (3, 3):
        raise ValueError("rot_matrix must be a 3x3 array")
    q = np.stack((np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta), np.cos(theta)))
    p = np.einsum("ab...,b...->a...", rot_matrix, q)
    if np.isclose(p[2], 1.0, rtol=0.0, atol=1e-12):
        p[2] = 1.0
    beta = np.arccos(p[2])
    alpha = np.arctan2(p[1], p[0])
    if alpha < 0:
        alpha += 2.0 * np.pi
    return beta, alpha


def spherical_basis_vector_rotation_matrix(theta, phi, rot_matrix, beta=None, alpha=None):
    """2x2 rotation taking the (theta, phi) tangent basis to the (beta, alpha) one.

    Reference :140-186: bh = R^T . theta_hat(beta, alpha); cosX = bh.theta_hat; sinX = bh.phi_hat;
    returns [[cosX, sinX], [-sinX, cosX]].
    """
    if alpha is None or beta is None:
        beta, alpha = rotate_points_3d(rot_matrix, theta, phi)
    th = theta_hat(theta, phi)
    ph = phi_hat(theta, phi)
    bh = np.einsum("ab...,b...->a...", rot_matrix.T, theta_hat(beta, alpha))
    cos_x = np.einsum("a...,a...", bh, th)
    sin_x = np.einsum("a...,a...", bh, ph)
    return np.array([[cos_x, sin_x], [-sin_x, cos_x]])


def is_orthogonal(matrix, tol=1e-15):
    """Reference :225-239 (note: np.allclose keeps its default rtol of 1e-5)."""
    return np.allclose(np.matmul(matrix, matrix.T), np.eye(3), atol=tol)


def is_unit_vector(vec, tol=1e-15):
    """Reference :242-256."""
    return np.allclose(np.dot(vec, vec), 1, atol=tol)


def axis_angle_rotation_matrix(axis, angle):
    """Rodrigues rotation I + sin(a) K + (1-cos(a)) K.K.  Reference :189-222."""
    if axis.shape != (3,):
        raise ValueError("axis must be a must be length 3 vector")
    if not is_unit_vector(axis):
        raise ValueError("axis must be a unit vector")
    k = np.array([[0.0, -axis[2], axis[1]], [axis[2], 0.0, -axis[0]], [-axis[1], axis[0], 0.0]])
    return np.identity(3) + np.sin(angle) * k + (1.0 - np.cos(angle)) * np.dot(k, k)


def vecs2rot(r1=None, r2=None, theta1=None, phi1=None, theta2=None, phi2=None):
    """Rotation taking unit vector r1 onto r2 about r1 x r2.  Reference :259-310.

    NaN axis (r1 == r2) raises "axis must be a unit vector", as upstream does (:300-303).
    """
    if r1 is None or r2 is None:
        if theta1 is None or phi1 is None or theta2 is None or phi2 is None:
            raise ValueError(
                "Either r1 and r2 must be supplied or all of "
                "theta1, phi1, theta2 and phi2 must be supplied."
            )
        r1 = r_hat(theta1, phi1)
        r2 = r_hat(theta2, phi2)
        assert is_unit_vector(r1) and is_unit_vector(r2)
    else:
        r1, r2 = np.array(r1), np.array(r2)
        if r1.shape != (3,) or r2.shape != (3,):
            raise ValueError("r1 and r2 must be length 3 vectors")
        if not is_unit_vector(r1) or not is_unit_vector(r2):
            raise ValueError("r1 and r2 must be unit vectors")
    norm = np.cross(r1, r2)
    sin_psi = np.sqrt(np.dot(norm, norm))
    n_hat = norm / sin_psi
    cos_psi = np.dot(r1, r2)
    psi = np.arctan2(sin_psi, cos_psi)
    rotation = axis_angle_rotation_matrix(n_hat, psi)
    assert is_unit_vector(n_hat), "n_hat is not a unit vector: " + str(n_hat)
    assert is_orthogonal(rotation), "rotation matrix is not orthogonal: " + str(rotation)
    return rotation
