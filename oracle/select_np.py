"""CPU restatement of the component down-selection of pyradiosky's SkyModel -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
(pyradiosky_b200) never does.

What is restated (reference: src/pyradiosky/skymodel.py), on plain numpy arrays in radians / stokes units:
    select_indices()      SkyModel.select :3127-3283, with _select_lat :3012-3026, _select_lon :3028-3045 and
                          _select_brightness :3047-3124 -- the component index list the reference ends up with
    nonrising_keep()      SkyModel.cut_nonrising :3369-3390
    rise_set_lsts()       SkyModel.calculate_rise_set_lsts :3284-3329

Parity status: pinned by the reference's own known answers -- the six component counts of
tests/test_skymodel.py:2174-2223 (46, 49, 32, 31, 29, 4 on the 50-source GLEAM fixture, checked in
tests/test_oracle.py) and the circumpolar / never-rising identities of :2661-2740.
"""

import warnings

import numpy as np


def _lat_cut(inds, lat, lat_range):
    # :3012-3026
    if np.asarray(lat_range).size != 2 or lat_range[0] >= lat_range[1]:
        raise ValueError("lat_range must be 2 element range with the second component larger than the first.")
    vals = lat[inds]
    return inds[np.nonzero((vals >= lat_range[0]) & (vals <= lat_range[1]))[0]]


def _lon_cut(inds, lon, lon_range):
    # :3028-3045; a range that ends below its start wraps through lon = 0 (and comes back sorted, np.union1d)
    if np.asarray(lon_range).size != 2:
        raise ValueError("lon_range must be 2 element range.")
    vals = lon[inds]
    if lon_range[1] < lon_range[0]:
        a = inds[np.nonzero(vals >= lon_range[0])[0]]
        b = inds[np.nonzero(vals <= lon_range[1])[0]]
        return np.union1d(a, b)
    return inds[np.nonzero((vals >= lon_range[0]) & (vals <= lon_range[1]))[0]]


def _brightness_cut(inds, stokes, spectral_type, freq_array, min_b, max_b, freq_range):
    # :3047-3124
    if spectral_type == "spectral_index":
        raise NotImplementedError("Flux cuts with spectral index type objects is not supported yet.")
    if freq_range is not None and np.atleast_1d(freq_range).size != 2:
        raise ValueError("brightness_freq_range must have 2 elements.")
    use = None
    if freq_array is not None:
        if freq_range is not None:
            use = np.where((freq_array >= np.min(freq_range)) & (freq_array <= np.max(freq_range)))[0]
            if use.size == 0:
                raise ValueError("No object frequencies in specified range for flux cuts.")
        else:
            use = np.arange(stokes.shape[1])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        if min_b is not None:
            s = stokes[0][:, inds] if use is None else stokes[0][use, :][:, inds]
            inds = inds[np.nonzero(np.min(s, axis=0) >= min_b)[0]]
        if max_b is not None:
            s = stokes[0][:, inds] if use is None else stokes[0][use, :][:, inds]
            inds = inds[np.nonzero(np.max(s, axis=0) <= max_b)[0]]
    return inds


def select_indices(lon, lat, stokes, spectral_type="flat", freq_array=None, component_inds=None, lat_range=None,
                   lon_range=None, min_brightness=None, max_brightness=None, brightness_freq_range=None, non_nan=None,
                   non_negative=False):
    """The component indices SkyModel.select keeps, in the order the reference leaves them (:3195-3252).

    Returns None when no criterion is given (the reference returns before doing anything)."""
    n = stokes.shape[2]
    if (component_inds is None and lat_range is None and lon_range is None and min_brightness is None
            and max_brightness is None and non_nan is None and non_negative is False):
        return None
    inds = None if component_inds is None else np.asarray(component_inds)
    if non_nan is not None:
        if non_nan not in ("any", "all"):
            raise ValueError("If set, non_nan can only be set to one of: ['any', 'all']")
        if non_nan == "any":
            good = np.nonzero(~np.any(np.isnan(stokes), axis=(0, 1)))[0]
        else:
            good = np.nonzero(~np.all(np.any(np.isnan(stokes), axis=0), axis=0))[0]
        inds = good if inds is None else np.intersect1d(inds, good)
    if non_negative:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            good = np.nonzero(~np.any(stokes[0] < 0, axis=0))[0]
        inds = good if inds is None else np.intersect1d(inds, good)
    if inds is None:
        inds = np.arange(n)
    if lat_range is not None:
        inds = _lat_cut(inds, lat, lat_range)
    if lon_range is not None:
        inds = _lon_cut(inds, lon, lon_range)
    if min_brightness is not None or max_brightness is not None:
        inds = _brightness_cut(inds, stokes, spectral_type, freq_array, min_brightness, max_brightness,
                               brightness_freq_range)
    if np.asarray(inds).size == 0:
        raise ValueError("Select would result in an empty object.")
    return np.asarray(inds)


def nonrising_keep(lat, telescope_latitude):
    """Indices cut_nonrising keeps: everything but tan(lat_tel) tan(lat) < -1 (:3378-3383)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        tans = np.tan(telescope_latitude) * np.tan(lat)
    return np.nonzero(~(tans < -1))[0]


def rise_set_lsts(lon, lat, telescope_latitude, horizon_buffer=0.04364):
    """(_rise_lst, _set_lst) of calculate_rise_set_lsts (:3309-3329); NaN where a component never rises or sets."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        tans = np.tan(telescope_latitude) * np.tan(lat)
        ha = np.arccos((-1) * tans)
        rise = lon - ha - horizon_buffer
        sett = lon + ha + horizon_buffer
        rise[rise < 0] += 2 * np.pi
        sett[sett < 0] += 2 * np.pi
        rise[rise > 2 * np.pi] -= 2 * np.pi
        sett[sett > 2 * np.pi] -= 2 * np.pi
    return rise, sett
