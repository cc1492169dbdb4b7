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


def ring2nest(nside, pix):
    """Ring-ordered pixel number -> nested (through the pixel centre, which both orderings share)."""
    lon, lat = healpix_to_lonlat(nside, pix, nested=False)
    return lonlat_to_healpix(nside, lon, lat, nested=True)


def _ring_info(nside, ring):
    """(first pixel, pixels in ring, colatitude, shifted) of ring 1 .. 4 nside - 1 (HEALPix get_ring_info2)."""
    nside = int(nside)
    npix = 12 * nside * nside
    ncap = 2 * nside * (nside - 1)
    north = np.where(ring > 2 * nside, 4 * nside - ring, ring)
    cap = north < nside
    tmp = (north * north).astype(np.float64) * (4.0 / npix)
    theta_cap = np.arctan2(np.sqrt(tmp * (2.0 - tmp)), 1.0 - tmp)
    theta_eq = np.arccos(np.clip((2 * nside - north).astype(np.float64) * (2.0 / (3.0 * nside)), -1.0, 1.0))
    theta = np.where(cap, theta_cap, theta_eq)
    ringpix = np.where(cap, 4 * north, 4 * nside)
    shifted = np.where(cap, True, ((north - nside) & 1) == 0)
    start = np.where(cap, 2 * north * (north - 1), ncap + (north - nside) * 4 * nside)
    south = north != ring
    theta = np.where(south, np.pi - theta, theta)
    start = np.where(south, npix - start - ringpix, start)
    return start, ringpix, theta, shifted


def interp_weights(nside, lon, lat, nested=False):
    """The four neighbours and bilinear weights of ``astropy_healpix.HEALPix.interpolate_bilinear_lonlat`` -- the
    interpolation SkyModel.healpix_interp_transform relies on (skymodel.py:1507-1509).

    astropy-healpix is not on this box; its interpolation (two pixels that bracket the longitude on each of the two
    rings that bracket the latitude, linear in longitude along a ring and in latitude between rings, a fictitious
    polar pixel equal to the mean of the four cap pixels inside the first / last ring) is the classic HEALPix
    ``get_interpol`` scheme, restated here.  Returns (pix [4, n] int64, wgt [4, n]).  Pinned by the reference's own
    golden pair gsm_galactic.skyh5 -> gsm_icrs.skyh5 (tests/test_skymodel.py:4423-4452) in tests/test_oracle.py.
    """
    nside = int(nside)
    lon = np.mod(np.atleast_1d(np.asarray(lon, dtype=np.float64)), 2 * np.pi)
    lat = np.atleast_1d(np.asarray(lat, dtype=np.float64))
    theta = np.pi / 2 - lat
    z = np.sin(lat)
    az = np.abs(z)
    ir_eq = np.floor(nside * (2.0 - 1.5 * z)).astype(np.int64)
    ir_cap = np.floor(nside * np.sqrt(3.0 * (1.0 - az))).astype(np.int64)
    ir1 = np.where(az <= 2.0 / 3.0, ir_eq, np.where(z > 0, ir_cap, 4 * nside - ir_cap - 1))  # ring above the point
    ir2 = ir1 + 1
    npix = 12 * nside * nside
    pix = np.zeros((4, lon.size), dtype=np.int64)
    wgt = np.zeros((4, lon.size))
    thetas = []
    for slot, ring in ((0, ir1), (2, ir2)):
        ok = (ring > 0) & (ring < 4 * nside)
        start, nr, th, shifted = _ring_info(nside, np.where(ok, ring, 1))
        dphi = 2 * np.pi / nr
        tmp = lon / dphi - 0.5 * shifted
        i1 = np.where(tmp < 0, tmp - 1, tmp).astype(np.int64)
        w1 = (lon - (i1 + 0.5 * shifted) * dphi) / dphi
        i2 = i1 + 1
        i1 = np.where(i1 < 0, i1 + nr, i1)
        i2 = np.where(i2 >= nr, i2 - nr, i2)
        pix[slot], pix[slot + 1] = start + i1, start + i2
        wgt[slot], wgt[slot + 1] = 1.0 - w1, w1
        thetas.append(th)
    theta1, theta2 = thetas
    northpole = ir1 == 0
    southpole = ir2 == 4 * nside
    mid = ~(northpole | southpole)
    with np.errstate(divide="ignore", invalid="ignore"):
        wt_mid = (theta - theta1) / (theta2 - theta1)
        wt_n = theta / theta2
        wt_s = (theta - theta1) / (np.pi - theta1)
    # between two rings
    w = wgt.copy()
    w[0], w[1] = wgt[0] * (1 - wt_mid), wgt[1] * (1 - wt_mid)
    w[2], w[3] = wgt[2] * wt_mid, wgt[3] * wt_mid
    # north of the first ring: the polar value is the mean of the four pixels of ring 1
    fac = (1 - wt_n) * 0.25
    wn = np.array([fac, fac, wgt[2] * wt_n + fac, wgt[3] * wt_n + fac])
    pn = np.array([(pix[2] + 2) & 3, (pix[3] + 2) & 3, pix[2], pix[3]])
    # south of the last ring
    fac = wt_s * 0.25
    ws = np.array([wgt[0] * (1 - wt_s) + fac, wgt[1] * (1 - wt_s) + fac, fac, fac])
    ps = np.array([pix[0], pix[1], ((pix[0] + 2) & 3) + npix - 4, ((pix[1] + 2) & 3) + npix - 4])
    out_w = np.where(mid, w, np.where(northpole, wn, ws))
    out_p = np.where(mid, pix, np.where(northpole, pn, ps))
    if nested:
        out_p = ring2nest(nside, out_p.ravel()).reshape(out_p.shape)
    return out_p, out_w


def interp_transform(nside, nested, hpx_inds, values, rot_new_to_old, full_sky=False):
    """SkyModel.healpix_interp_transform (skymodel.py:1448-1538) for one map row: ``values`` on pixels ``hpx_inds`` of
    the old frame -> the map at the pixel centres of the new frame, ``rot_new_to_old`` being the rotation of unit vectors
    from the new frame to the old one.  Returns (new_hpx_inds, new_values); a new pixel whose four neighbours are not
    all present is dropped (masked-array arithmetic upstream), or kept as 0 when ``full_sky``."""
    npix = 12 * int(nside) ** 2
    lon, lat = healpix_to_lonlat(nside, np.arange(npix), nested=nested)
    v = rot_new_to_old @ np.array([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)])
    lon_o, lat_o = np.arctan2(v[1], v[0]), np.arctan2(v[2], np.hypot(v[0], v[1]))
    pix, wgt = interp_weights(nside, lon_o, lat_o, nested=nested)
    lookup = np.full(npix, -1, dtype=np.int64)
    lookup[np.asarray(hpx_inds)] = np.arange(len(hpx_inds))
    pos = lookup[pix]
    valid = np.all(pos >= 0, axis=0)
    vals = np.asarray(values, dtype=np.float64)
    out = np.sum(wgt * vals[np.where(pos >= 0, pos, 0)], axis=0)
    out = np.where(valid, out, 0.0)
    if full_sky:
        return np.arange(npix), out
    keep = np.nonzero(valid)[0]
    return keep, out[keep]
