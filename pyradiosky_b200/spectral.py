"""Host-side preparation of the subband interpolation operator.

``SkyModel.at_frequencies`` interpolates subband spectra with
``scipy.interpolate.interp1d(freq_array, stokes, axis=1, kind=freq_interp_kind)``
(src/pyradiosky/skymodel.py:2362-2368).  For every supported ``kind`` that interpolant is (a) linear in
the data and (b) built on knots shared by all components, so it is a piecewise polynomial whose
coefficients are a fixed linear map of the stored channel values.  The map is computed once here, with
scipy itself (the same ``make_interp_spline`` call ``interp1d`` makes), and handed to the kernel as a
small dense operator; the kernel then needs ``(degree+1) * Nfreqs_in`` FMAs per component per break
interval and a Horner evaluation per output channel.
"""

import numpy as np
import scipy.interpolate

_SPLINE_ORDER = {"slinear": 1, "quadratic": 2, "cubic": 3}
# interp1d kinds that pick one stored channel per request frequency (piecewise constant in f)
STEP_KINDS = ("nearest", "nearest-up", "zero", "previous", "next")


def is_step_kind(kind):
    return (isinstance(kind, str) and kind in STEP_KINDS) or (
        isinstance(kind, (int, np.integer)) and not isinstance(kind, bool) and int(kind) == 0)


def interp_degree(kind):
    """Polynomial degree of ``interp1d(kind=...)`` or raise for kinds the kernel does not cover."""
    if is_step_kind(kind):
        return 0
    if isinstance(kind, (int, np.integer)):
        order = int(kind)
    elif kind == "linear":
        order = 1
    elif kind in _SPLINE_ORDER:
        order = _SPLINE_ORDER[kind]
    else:
        raise NotImplementedError(
            f"freq_interp_kind={kind!r} is not supported on the GPU path; use 'linear', 'nearest', 'nearest-up', "
            "'zero', 'slinear', 'quadratic', 'cubic', 'previous', 'next' or a spline order 0-3."
        )
    if order < 0 or order > 3:
        raise NotImplementedError("Only spline orders 0 to 3 are supported on the GPU path.")
    return order


def build_step_operator(freq_array, kind):
    """Piecewise-constant form of ``interp1d(freq_array, ., kind)`` for the channel-picking kinds ('nearest',
    'nearest-up', 'zero', 'previous', 'next', order 0), in the layout of :func:`build_ppoly_operator` with degree 0.

    Which side of a knot (or of a midpoint between knots) belongs to which channel differs between the kinds and
    is decided by scipy itself: ``interp1d`` of an identity matrix is evaluated at every possible jump location
    ``p`` and at the next double after it, so each returned interval ``[breaks[b], breaks[b+1])`` -- closed on the
    left, as the kernel looks intervals up -- carries exactly scipy's choice, ties included."""
    x = np.asarray(freq_array, dtype=np.float64)
    n = x.size
    order = np.argsort(x, kind="stable")
    xs = x[order]
    if np.any(np.diff(xs) <= 0):
        raise ValueError("Expect x to not have duplicates")
    big = 1e300
    if n == 1:
        raise ValueError("x and y arrays must have at least 2 entries")
    kind_use = 0 if isinstance(kind, (int, np.integer)) else kind
    finterp = scipy.interpolate.interp1d(xs, np.eye(n)[order], axis=0, kind=kind_use)
    mids = 0.5 * (xs[1:] + xs[:-1])
    jumps = np.unique(np.concatenate((xs, mids)))
    cand = np.unique(np.concatenate((jumps, np.nextafter(jumps, np.inf))))
    cand = cand[(cand >= xs[0]) & (cand <= xs[-1])]
    vals = finterp(cand)  # (ncand, n): one-hot rows
    keep = np.ones(cand.size, dtype=bool)
    keep[1:] = np.any(vals[1:] != vals[:-1], axis=1)
    breaks = np.concatenate((cand[keep], [big]))
    ppoly = np.ascontiguousarray(vals[keep][:, None, :])
    return 0, breaks, ppoly


def build_ppoly_operator(freq_array, kind="cubic"):
    """Piecewise-polynomial form of ``interp1d(freq_array, ., kind)``.

    Returns ``(degree, breaks, ppoly)`` with ``breaks`` of shape (nb+1,) and ``ppoly`` of shape
    (nb, degree+1, Nfreqs_in): on interval b, with u = (f - breaks[b]) / (breaks[b+1] - breaks[b]),
    ``value = sum_p u**p * (ppoly[b, p, :] @ y)`` for stored channel values ``y``.
    """
    if is_step_kind(kind):
        return build_step_operator(freq_array, kind)
    x = np.asarray(freq_array, dtype=np.float64)
    n = x.size
    degree = interp_degree(kind)
    if n <= degree:
        # same condition interp1d raises on (and which makes the reference fall back to linear
        # for NaN-depleted components, skymodel.py:2302-2309)
        raise ValueError(f"The number of derivatives at boundaries does not match: expected "
                         f"at least {degree + 1} points for kind={kind!r}, got {n}")
    order = np.argsort(x, kind="stable")
    xs = x[order]
    if np.any(np.diff(xs) <= 0):
        raise ValueError("Expect x to not have duplicates")
    eye = np.eye(n)[order]  # data rows follow the sorted knots; columns = original channel index
    if degree == 1:
        breaks = xs
        ppoly = np.zeros((n - 1, 2, n))
        ppoly[:, 0, :] = eye[:-1]
        ppoly[:, 1, :] = eye[1:] - eye[:-1]
        return degree, breaks, ppoly
    spl = scipy.interpolate.make_interp_spline(xs, eye, k=degree, check_finite=False)
    # B-spline -> piecewise polynomial: Taylor coefficients at the left end of every non-empty knot
    # interval (BSpline evaluation is right-continuous at interior knots).
    t = spl.t
    breaks = np.unique(t[degree : t.size - degree])
    h = np.diff(breaks)
    ppoly = np.zeros((h.size, degree + 1, n))
    fact = 1.0
    for p in range(degree + 1):
        if p > 0:
            fact *= p
        deriv = spl.derivative(p) if p > 0 else spl
        ppoly[:, p, :] = deriv(breaks[:-1]) / fact * (h**p)[:, None]
    return degree, breaks, ppoly


def evaluate_operator(degree, breaks, ppoly, y, freqs):
    """numpy evaluation of the operator (host-side self-check helper; not used on the hot path)."""
    freqs = np.asarray(freqs, dtype=np.float64)
    b = np.clip(np.searchsorted(breaks, freqs, side="right") - 1, 0, len(breaks) - 2)
    u = (freqs - breaks[b]) / (breaks[b + 1] - breaks[b])
    coef = np.einsum("bpj,j...->bp...", ppoly, y)  # (nb, degree+1, ...)
    out = 0.0
    for p in range(degree, -1, -1):
        out = out * u.reshape((-1,) + (1,) * (coef.ndim - 2)) + coef[b, p]
    return out


def build_bspline_operator(freq_array, kind="cubic"):
    """Banded (B-spline) form of ``interp1d(freq_array, ., kind)``.

    The interpolant is a spline ``sum_i c_i B_i(f)`` whose coefficients are a fixed linear map of the
    stored channel values, ``c = cmat @ y`` (dense, time independent: computed once per catalogue on
    the GPU by ``rsk_spline_coefficients``), and on every break interval only ``degree + 1`` basis
    functions are non-zero.  Returns ``(degree, breaks, cmat, first, pband)``:

    * ``cmat``  (nc, Nfreqs_in)  -- data -> spline coefficients,
    * ``first`` (nb,) int32      -- index of the first coefficient used on interval b,
    * ``pband`` (nb, degree+1, degree+1) -- with u = (f - breaks[b]) / (breaks[b+1] - breaks[b]),
      ``value = sum_p u**p * sum_m pband[b, p, m] * c[first[b] + m]``.
    """
    x = np.asarray(freq_array, dtype=np.float64)
    n = x.size
    degree = interp_degree(kind)
    if n <= degree:
        raise ValueError(f"The number of derivatives at boundaries does not match: expected "
                         f"at least {degree + 1} points for kind={kind!r}, got {n}")
    order = np.argsort(x, kind="stable")
    xs = x[order]
    if np.any(np.diff(xs) <= 0):
        raise ValueError("Expect x to not have duplicates")
    spl = scipy.interpolate.make_interp_spline(xs, np.eye(n)[order], k=degree, check_finite=False)
    t = spl.t
    cmat = np.ascontiguousarray(spl.c)  # (nc, n): columns are original channel indices
    nc = cmat.shape[0]
    breaks = np.unique(t[degree: t.size - degree])
    h = np.diff(breaks)
    nb = h.size
    basis = scipy.interpolate.BSpline(t, np.eye(nc), degree)
    span = np.searchsorted(t, breaks[:-1], side="right") - 1  # t[span] <= break < t[span + 1]
    first = (span - degree).astype(np.int32)
    pband = np.zeros((nb, degree + 1, degree + 1))
    fact = 1.0
    for p in range(degree + 1):
        if p > 0:
            fact *= p
        dv = (basis.derivative(p) if p > 0 else basis)(breaks[:-1]) / fact * (h**p)[:, None]  # (nb, nc)
        for b in range(nb):
            pband[b, p, :] = dv[b, first[b]: first[b] + degree + 1]
    return degree, breaks, cmat, first, pband


def evaluate_bspline_operator(degree, breaks, cmat, first, pband, y, freqs):
    """numpy evaluation of the banded operator (host-side self-check helper)."""
    freqs = np.asarray(freqs, dtype=np.float64)
    c = np.tensordot(cmat, y, axes=(1, 0))  # (nc, ...)
    b = np.clip(np.searchsorted(breaks, freqs, side="right") - 1, 0, len(breaks) - 2)
    u = (freqs - breaks[b]) / (breaks[b + 1] - breaks[b])
    out = np.zeros((freqs.size,) + c.shape[1:])
    for k in range(freqs.size):
        loc = c[first[b[k]]: first[b[k]] + degree + 1]
        coef = np.tensordot(pband[b[k]], loc, axes=(1, 0))  # (degree+1, ...)
        acc = 0.0
        for p in range(degree, -1, -1):
            acc = acc * u[k] + coef[p]
        out[k] = acc
    return out


def build_nan_group_operator(freq_array, use, kind="cubic", nan_handling="clip"):
    """Dense operator for components whose only non-NaN stored channels are ``use`` (storage indices).

    Restates the per-component branch of skymodel.py:2245-2312 as a piecewise polynomial over the whole
    frequency axis: inside [min, max] of the valid channel frequencies the ``interp1d`` of the valid
    subset (falling back to linear when there are too few points for ``kind``, :2302-2309); outside,
    the value at the first / last valid channel (``clip``) or NaN (``interp``).  ``propagate`` and
    components without any valid channel give NaN everywhere.

    Returns ``(degree, breaks, ppoly, chan_map, fallback_linear)`` with ``ppoly`` of shape
    (nb, degree+1, len(chan_map)).
    """
    if nan_handling not in ("propagate", "interp", "clip"):
        raise ValueError("nan_handling must be one of ['propagate', 'interp', 'clip']")
    x = np.asarray(freq_array, dtype=np.float64)
    use = np.asarray(use, dtype=np.int64)
    big = 1e300
    if nan_handling == "propagate" or use.size == 0:
        return 0, np.array([0.0, big]), np.full((1, 1, 1), np.nan), np.zeros(1, dtype=np.int32), False
    xs = x[use]
    fmin, fmax = xs.min(), xs.max()
    fallback = False
    if use.size == 1:
        # One valid channel: upstream falls back to a linear interp1d on a single point (:2302-2309),
        # whose 0/0 weights give NaN at that exact frequency; outside it clips / NaNs as usual.
        degree = 0
        core_breaks = np.array([fmin, np.nextafter(fmin, np.inf)])
        core = np.full((1, 1, 1), np.nan)
        fallback = kind != "linear"  # interp1d(kind="linear") accepts one point silently
    else:
        degree = interp_degree(kind)
        eff_kind = kind
        if use.size <= degree:
            eff_kind, degree, fallback = "linear", 1, True
        degree, core_breaks, core = build_ppoly_operator(xs, eff_kind)
    nuse = use.size
    edge = np.nan if nan_handling == "interp" else 1.0
    low = np.zeros((1, degree + 1, nuse))
    high = np.zeros((1, degree + 1, nuse))
    # the reference clips to stokes[:, freq_inds_use[0]] / [-1]: first / last valid index in storage order
    low[0, 0, 0] = edge
    high[0, 0, nuse - 1] = edge
    if nan_handling == "interp":
        low[0, 0, :] = np.nan
        high[0, 0, :] = np.nan
    # f == fmax must still be interpolated (reference: at_freq_arr <= max), so the upper edge interval
    # starts one ulp above it
    breaks = np.concatenate(([min(0.0, fmin - 1.0)], core_breaks[:-1], [np.nextafter(fmax, np.inf), big]))
    last_core = core[-1:].copy()
    ppoly = np.concatenate((low, core[:-1], last_core, high), axis=0) if core.shape[0] > 1 else np.concatenate((low, last_core, high), axis=0)
    return degree, breaks, ppoly, use.astype(np.int32), fallback
