"""Oracle (test infrastructure): HEALPix pixel centres, ring and nested ordering.

Stands in for ``astropy_healpix.HEALPix(nside, order, frame).healpix_to_skycoord(hpx_inds)`` at the
reference call site ``src/pyradiosky/skymodel.py:1664-1667`` (``get_lon_lat`` healpix branch).
astropy-healpix (>=1.0.2, reference pyproject.toml:43) is a third-party dependency absent from this
image, so the published HEALPix pix2ang formulas (Gorski et al. 2005) are restated.  PARITY UNPINNED
against astropy-healpix itself; pinned by invariants (ring/nested give the same centre set, equal
areas => sum of unit vectors ~ 0, z non-increasing in ring order, pixel 357 @ nside 8 of
reference test_skymodel.py:2094).
"""

import numpy as np

_JRLL = np.array([2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4], dtype=np.int64)
_JPLL = np.array([1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7], dtype=np.int64)


def _isqrt(v):
    r = np.floor(np.sqrt(v.astype(np.float64))).astype(np.int64)
    r = np.where(r * r > v, r - 1, r)
    r = np.where((r + 1) * (r + 1) <= v, r + 1, r)
    return r


def _compact_bits(v):
    """Keep the even bits of v and pack them (inverse of bit interleaving)."""
    v = v & 0x5555555555555555
    v = (v | (v >> 1)) & 0x3333333333333333
    v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0F
    v = (v | (v >> 4)) & 0x00FF00FF00FF00FF
    v = (v | (v >> 8)) & 0x0000FFFF0000FFFF
    v = (v | (v >> 16)) & 0x00000000FFFFFFFF
    return v


def pix2zphi(nside, pix, nested=False):
    """Return (z, one_minus_abs_z, phi) of pixel centres; the exact 1-|z| is kept for the caps."""
    nside = int(nside)
    pix = np.asarray(pix, dtype=np.int64)
    npix = 12 * nside * nside
    ncap = 2 * nside * (nside - 1)
    nl4 = 4 * nside
    if nested:
        npface = nside * nside
        face = pix // npface
        ipf = pix % npface
        ix = _compact_bits(ipf)
        iy = _compact_bits(ipf >> 1)
        jr = _JRLL[face] * nside - ix - iy - 1
        north = jr < nside
        south = jr > 3 * nside
        nr = np.where(north, jr, np.where(south, nl4 - jr, nside))
        kshift = np.where(north | south, 0, (jr - nside) & 1)
        jp = (_JPLL[face] * nr + ix - iy + 1 + kshift) // 2
        jp = np.where(jp > nl4, jp - nl4, jp)
        jp = np.where(jp < 1, jp + nl4, jp)
        iring_eq = jr
    else:
        north = pix < ncap
        south = pix >= npix - ncap
        # north cap
        irn = (1 + _isqrt(1 + 2 * np.where(north, pix, 0))) >> 1
        ipn = pix + 1 - 2 * irn * (irn - 1)
        # south cap
        ips = npix - pix
        irs = (1 + _isqrt(2 * np.where(south, ips, 1) - 1)) >> 1
        ipsn = 4 * irs + 1 - (ips - 2 * irs * (irs - 1))
        # equatorial belt
        ipe = pix - ncap
        ire = ipe // nl4 + nside
        ipen = ipe % nl4 + 1
        nr = np.where(north, irn, np.where(south, irs, nside))
        jp = np.where(north, ipn, np.where(south, ipsn, ipen))
        kshift = np.where(north | south, 0, (ire - nside) & 1)
        iring_eq = ire
    nrf = nr.astype(np.float64)
    omz_cap = nrf * nrf / (3.0 * nside * nside)
    z_eq = (2 * nside - iring_eq).astype(np.float64) * (2.0 / (3.0 * nside))
    z = np.where(north, 1.0 - omz_cap, np.where(south, omz_cap - 1.0, z_eq))
    omz = np.where(north | south, omz_cap, 1.0 - np.abs(z_eq))
    phi = (jp.astype(np.float64) - (kshift + 1) * 0.5) * ((np.pi / 2.0) / nrf)
    return z, omz, phi


def healpix_to_lonlat(nside, pix, nested=False):
    """Pixel-centre (lon, lat) in radians; lon in [0, 2pi)."""
    z, omz, phi = pix2zphi(nside, pix, nested=nested)
    st = np.sqrt(omz * (2.0 - omz))  # sqrt((1-|z|)(1+|z|)) without cancellation near the poles
    return phi, np.arctan2(z, st)


def _spread_bits(v):
    x = np.asarray(v, dtype=np.uint64) & np.uint64(0xFFFFFFFF)
    for shift, mask in ((16, 0x0000FFFF0000FFFF), (8, 0x00FF00FF00FF00FF), (4, 0x0F0F0F0F0F0F0F0F),
                        (2, 0x3333333333333333), (1, 0x5555555555555555)):
        x = (x | (x << np.uint64(shift))) & np.uint64(mask)
    return x.astype(np.int64)


def lonlat_to_healpix(nside, lon, lat, nested=False):
    """Pixel index containing each (lon, lat) [radians]: the published HEALPix ang2pix (Gorski et al. 2005),
    which is what ``astropy_healpix.HEALPix.skycoord_to_healpix`` computes for
    ``SkyModel.assign_to_healpix`` (src/pyradiosky/skymodel.py:1916).  astropy-healpix is absent from this
    image, so parity with it on exact pixel boundaries is unpinned; interior points (every test here) are
    unambiguous."""
    lon = np.atleast_1d(np.asarray(lon, dtype=np.float64))
    lat = np.atleast_1d(np.asarray(lat, dtype=np.float64))
    nside = int(nside)
    z = np.sin(lat)
    za = np.abs(z)
    sth = np.abs(np.cos(lat))
    tt = np.mod(lon, 2 * np.pi) / (np.pi / 2)
    tt = np.where(tt >= 4.0, 0.0, tt)
    nl4 = 4 * nside
    pix = np.zeros(lon.shape, dtype=np.int64)
    eq = za <= 2.0 / 3.0
    # equatorial belt
    temp1 = nside * (0.5 + tt)
    temp2 = nside * (z * 0.75)
    jp = np.floor(temp1 - temp2).astype(np.int64)
    jm = np.floor(temp1 + temp2).astype(np.int64)
    if nested:
        ifp, ifm = jp // nside, jm // nside
        face = np.where(ifp == ifm, ifp | 4, np.where(ifp < ifm, ifp, ifm + 8))
        ix = jm & (nside - 1)
        iy = nside - (jp & (nside - 1)) - 1
        pix_eq = (face & 15) * nside * nside + _spread_bits(ix) + 2 * _spread_bits(iy)
    else:
        ir = nside + 1 + jp - jm
        kshift = 1 - (ir & 1)
        ip = ((jp + jm - nside + kshift + 1 + 2 * nl4) >> 1) % nl4
        pix_eq = 2 * nside * (nside - 1) + (ir - 1) * nl4 + ip
    # polar caps
    ntt = np.minimum(tt.astype(np.int64), 3)
    tp = tt - ntt
    tmp = np.where(za > 0.99, nside * sth / np.sqrt((1.0 + za) / 3.0), nside * np.sqrt(3.0 * np.maximum(1.0 - za, 0.0)))
    jpc = np.floor(tp * tmp).astype(np.int64)
    jmc = np.floor((1.0 - tp) * tmp).astype(np.int64)
    if nested:
        jpn = np.minimum(jpc, nside - 1)
        jmn = np.minimum(jmc, nside - 1)
        north = z >= 0
        face = np.where(north, ntt, ntt + 8)
        ix = np.where(north, nside - jmn - 1, jpn)
        iy = np.where(north, nside - jpn - 1, jmn)
        pix_cap = face * nside * nside + _spread_bits(ix) + 2 * _spread_bits(iy)
    else:
        ir = jpc + jmc + 1
        ip = np.floor(tt * ir).astype(np.int64) % (4 * ir)
        pix_cap = np.where(z > 0, 2 * ir * (ir - 1) + ip, 12 * nside * nside - 2 * ir * (ir + 1) + ip)
    return np.where(eq, pix_eq, pix_cap)


def nside_to_pixel_area(nside):
    """astropy_healpix.nside_to_pixel_area: 4 pi / (12 nside^2) steradian."""
    import math

    return 4 * math.pi / (12 * int(nside) * int(nside))
