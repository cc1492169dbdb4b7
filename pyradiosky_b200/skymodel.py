"""``SkyModel``: the host-side mirror of ``pyradiosky.SkyModel`` for the per-component hot path.

Same method names, argument meaning and error behaviour as the reference class
(src/pyradiosky/skymodel.py:335) for

    update_positions   :2413-2479        coherency_calc        :2648-2712
    at_frequencies     :2084-2389        calc_frame_coherency  :2481-2497
    jansky_to_kelvin   :1593-1635        kelvin_to_jansky      :1541-1591
    get_lon_lat        :1646-1673        clear_time_position_specific_params

but every array lives in HBM as a torch CUDA tensor (torch is used only as the device-buffer
allocator and stream provider) and all arithmetic runs in librsk's sm_100a kernels through ctypes.
There is no CPU fallback: constructing a SkyModel without a CUDA device or without the built
library raises.

Differences from the reference that a caller can observe (all documented in DESIGN.md):
  * arrays are torch CUDA tensors (``.cpu().numpy()`` gives the reference's numpy layout); units are
    carried as a plain string ``stokes_unit`` instead of astropy Quantities;
  * ``coherency_calc(freqs=...)`` is an extension that fuses ``at_frequencies`` into the coherency
    kernel so the (4, Nfreqs, Ncomponents) Stokes cube is never materialised;
  * ``coherency_calc`` never needs ``frame_coherency`` (it is computed on the fly), so
    ``store_frame_coherency=False`` works here (it raises TypeError upstream, skymodel.py:2685);
  * a source whose R_avg-rotated direction coincides exactly with its AltAz direction gets the
    identity perturbation instead of upstream's "axis must be a unit vector" ValueError.
"""

import copy as _copy
import ctypes
import math
import os
import queue as _queue
import threading as _threading
import warnings

import numpy as np
import torch

from . import _lib, astrometry, spectral
from .utils import _as_device_f64

# running estimates used by coherency_calc_to_host(packed="auto"): PCIe device->host rate and the host cost of
# expanding one packed evaluation (measured on every call; process-wide because they describe the machine)
_E2E_STATE = {"link_bytes_per_s": 50e9, "unpack_s_per_eval": 1.0e-9}

_JY_UNITS = {"point": "Jy", "healpix": "Jy / sr"}
_K_UNITS = {"point": "K sr", "healpix": "K"}


def _norm_unit(unit):
    return " ".join(str(unit).replace("/", " / ").split())


def _to_hz_host(freqs):
    """1-D float64 numpy array in Hz from Quantity-like / torch / array input."""
    if freqs is None:
        return None
    if hasattr(freqs, "to_value"):
        freqs = freqs.to_value("Hz")
    if isinstance(freqs, torch.Tensor):
        freqs = freqs.detach().cpu().numpy()
    return np.atleast_1d(np.asarray(freqs, dtype=np.float64))


def _to_rad(angle):
    if hasattr(angle, "rad"):
        angle = angle.rad
    elif hasattr(angle, "to_value"):
        angle = angle.to_value("rad")
    return angle


class SkyModel:
    """Catalogue of point sources or HEALPix pixels held on one GPU.

    Parameters
    ----------
    name : sequence of str, optional
        Component names (kept on the host, unused by the hot path).
    ra, dec / lon, lat : array_like, radians (or objects with ``.rad``)
        Component coordinates in ``frame`` (point catalogues).
    frame : str
        'icrs' (default), 'fk5' (``equinox`` as a Julian epoch, default 2000.0) or 'galactic'.
    stokes : array_like, shape (4, Nfreqs, Ncomponents)
        [I, Q, U, V]; numpy, torch or Quantity-like.
    spectral_type : {'flat', 'spectral_index', 'subband', 'full'}
    freq_array : array_like, Hz -- required for 'subband' and 'full'.
    reference_frequency, spectral_index : array_like (Ncomponents,) -- required for 'spectral_index'.
    component_type : {'point', 'healpix'}
    nside, hpx_inds, hpx_order, hpx_frame : HEALPix description; ``hpx_inds=None`` with
        ``component_type='healpix'`` means "all 12 nside^2 pixels" (or the contiguous range starting at
        ``hpx_first`` when ``stokes`` holds fewer), without materialising an index array.
    stokes_unit : str
        'Jy' / 'K sr' (point) or 'Jy / sr' / 'K' (healpix).  Defaults to the Jy flavour, or to the
        unit of ``stokes`` if it is Quantity-like.
    device : torch device, default current CUDA device.
    freq_edge_array, history, filename, extended_model_group, stokes_error, beam_amp, extra_columns :
        Catalogue metadata carried through file I/O (skymodel.py:578-700); not used by the hot path.
    """

    def __init__(
        self,
        name=None,
        ra=None,
        dec=None,
        lon=None,
        lat=None,
        frame="icrs",
        equinox=2000.0,
        stokes=None,
        spectral_type=None,
        freq_array=None,
        reference_frequency=None,
        spectral_index=None,
        component_type=None,
        nside=None,
        hpx_inds=None,
        hpx_order=None,
        hpx_frame=None,
        hpx_first=0,
        stokes_unit=None,
        device=None,
        freq_edge_array=None,
        history=None,
        filename=None,
        extended_model_group=None,
        stokes_error=None,
        beam_amp=None,
        extra_columns=None,
    ):
        _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("pyradiosky_b200 needs a CUDA device; there is no CPU fallback.")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if spectral_type not in _lib.SPECTRAL_MODES:
            raise ValueError("spectral_type must be one of 'flat', 'spectral_index', 'subband', 'full'")
        self.spectral_type = spectral_type
        self.name = None if name is None else np.atleast_1d(np.asarray(name))

        if component_type is None:
            component_type = "healpix" if nside is not None else "point"
        if component_type not in ("point", "healpix"):
            raise ValueError("component_type must be 'point' or 'healpix'")
        self.component_type = component_type

        unit = getattr(stokes, "unit", None)
        if stokes_unit is None:
            stokes_unit = str(unit) if unit is not None else _JY_UNITS[component_type]
        self.stokes_unit = _norm_unit(stokes_unit)
        st = _as_device_f64(stokes, self.device)
        if st.dim() == 1:
            st = st.reshape(4, 1, 1)
        if st.dim() != 3 or st.shape[0] != 4:
            raise ValueError("stokes must have shape (4, Nfreqs, Ncomponents)")
        self.stokes = st
        self.Nfreqs = int(st.shape[1])
        self.Ncomponents = int(st.shape[2])

        if component_type == "point":
            lon = ra if lon is None else lon
            lat = dec if lat is None else lat
            if lon is None or lat is None:
                raise ValueError("ra/dec (or lon/lat) are required for point components")
            self.lon = _as_device_f64(np.atleast_1d(_to_rad(lon)) if not isinstance(lon, torch.Tensor) else lon, self.device)
            self.lat = _as_device_f64(np.atleast_1d(_to_rad(lat)) if not isinstance(lat, torch.Tensor) else lat, self.device)
            if self.lon.numel() != self.Ncomponents or self.lat.numel() != self.Ncomponents:
                raise ValueError("coordinate arrays must have length Ncomponents")
            self.frame = str(frame).lower()
            # a point object made by healpix_to_point keeps its pixel description (skymodel.py:1716-1727), so that
            # _point_to_healpix can undo the conversion -- also after a write / read round trip
            self.nside = None if nside is None else int(nside)
            self.hpx_order = None if hpx_order is None else str(hpx_order).lower()
            self.hpx_inds = None
            if hpx_inds is not None:
                if isinstance(hpx_inds, torch.Tensor):
                    self.hpx_inds = hpx_inds.to(self.device, dtype=torch.int64).contiguous()
                else:
                    self.hpx_inds = torch.as_tensor(np.ascontiguousarray(hpx_inds, dtype=np.int64)).to(self.device)
            self.hpx_first = 0
        else:
            if nside is None:
                raise ValueError("nside is required for healpix components")
            self.nside = int(nside)
            self.hpx_order = "ring" if hpx_order is None else str(hpx_order).lower()
            if self.hpx_order not in ("ring", "nested"):
                raise ValueError("hpx_order must be 'ring' or 'nested'")
            self.frame = str(hpx_frame if hpx_frame is not None else frame).lower()
            self.hpx_first = int(hpx_first)
            if hpx_inds is None:
                self.hpx_inds = None
                if self.hpx_first + self.Ncomponents > 12 * self.nside**2:
                    raise ValueError("pixel range exceeds 12 * nside**2")
            else:
                if isinstance(hpx_inds, torch.Tensor):
                    self.hpx_inds = hpx_inds.to(self.device, dtype=torch.int64).contiguous()
                else:
                    self.hpx_inds = torch.as_tensor(np.ascontiguousarray(hpx_inds, dtype=np.int64)).to(self.device)
                if self.hpx_inds.numel() != self.Ncomponents:
                    raise ValueError("hpx_inds must have length Ncomponents")
            self.lon = None
            self.lat = None
        self.hpx_frame = self.frame if component_type == "healpix" else None
        self.equinox = float(equinox)
        self._frame_to_icrs = astrometry.frame_to_icrs_matrix(self.frame, self.equinox)

        self.freq_array = _to_hz_host(freq_array)
        if spectral_type in ("subband", "full"):
            if self.freq_array is None:
                raise ValueError("freq_array must be provided for spectral_type 'subband' and 'full'")
            if self.freq_array.size != self.Nfreqs:
                raise ValueError("freq_array must have length Nfreqs")
        elif self.Nfreqs != 1:
            raise ValueError("stokes must have Nfreqs == 1 for 'flat' and 'spectral_index' spectral types")
        self.reference_frequency = None
        self.spectral_index = None
        if reference_frequency is not None:
            rf = reference_frequency.to_value("Hz") if hasattr(reference_frequency, "to_value") else reference_frequency
            self.reference_frequency = _as_device_f64(np.atleast_1d(rf) if not isinstance(rf, torch.Tensor) else rf, self.device)
        if spectral_index is not None:
            self.spectral_index = _as_device_f64(
                np.atleast_1d(spectral_index) if not isinstance(spectral_index, torch.Tensor) else spectral_index, self.device)
        if spectral_type == "spectral_index":
            if self.reference_frequency is None or self.spectral_index is None:
                raise ValueError("reference_frequency and spectral_index are required for spectral_type 'spectral_index'")
            if self.reference_frequency.numel() != self.Ncomponents or self.spectral_index.numel() != self.Ncomponents:
                raise ValueError("reference_frequency and spectral_index must have length Ncomponents")

        # skymodel.py:1037-1042 -- frozen at construction, as upstream
        self._polarized_mask = torch.empty(self.Ncomponents, dtype=torch.uint8, device=self.device)
        self._call("rsk_polarized_mask", _lib.ptr(self.stokes), self.Nfreqs, self.Ncomponents,
                   _lib.ptr(self._polarized_mask))
        self._n_polarized = self._count(self._polarized_mask, _lib.COUNT_NONZERO_U8)

        self.freq_edge_array = None if freq_edge_array is None else np.asarray(_to_hz_host(freq_edge_array)).reshape(2, -1)
        self.history = history
        self.filename = filename
        self.extended_model_group = None if extended_model_group is None else np.atleast_1d(np.asarray(extended_model_group))
        self.stokes_error = None if stokes_error is None else _as_device_f64(stokes_error, self.device)
        self.beam_amp = None if beam_amp is None else _as_device_f64(beam_amp, self.device)
        self.extra_columns = extra_columns
        self.freq_tol = 1.0  # Hz, skymodel.py:491
        self.time = None
        self.telescope_location = None
        self.alt_az = None
        self.pos_lmn = None
        self.above_horizon = None
        self.frame_coherency = None
        self._astrom = None
        self._freq_dev_cache = {}

    # ------------------------------------------------------------------------------------------
    # plumbing
    # ------------------------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _call(self, fname, *args):
        with torch.cuda.device(self.device):
            code = getattr(_lib.load(), fname)(*args, self._stream())
        _lib.check(fname, code)
        _lib.count_launches(fname)

    def _count(self, tensor, kind):
        """Number of NaNs (COUNT_NAN_F64) / non-zero bytes (COUNT_NONZERO_U8) of a device array (rsk_count)."""
        out = torch.empty(1, dtype=torch.int64, device=self.device)
        self._call("rsk_count", _lib.ptr(tensor), tensor.numel(), kind, _lib.ptr(out))
        return int(out.item())

    @property
    def _polarized(self):
        """Indices of polarised components (skymodel.py:1039-1041) as a CUDA int64 tensor."""
        return torch.nonzero(self._polarized_mask, as_tuple=False).flatten()

    def _coords_struct(self):
        c = _lib.RskCoords()
        if self.component_type == "healpix":
            c.lon = None
            c.lat = None
            c.nside = self.nside
            c.nested = 1 if self.hpx_order == "nested" else 0
            c.hpx_inds = None if self.hpx_inds is None else self.hpx_inds.data_ptr()
            c.hpx_first = self.hpx_first
        else:
            c.lon = self.lon.data_ptr()
            c.lat = self.lat.data_ptr()
            c.nside = 0
        return c

    def _freqs_dev(self, freqs_host):
        key = freqs_host.tobytes()
        t = self._freq_dev_cache.get(key)
        if t is None:
            if len(self._freq_dev_cache) > 8:
                self._freq_dev_cache.clear()
            t = torch.as_tensor(freqs_host).to(self.device)
            self._freq_dev_cache[key] = t
        return t

    def _spectrum_struct(self, mode, keep, freq_interp_kind="cubic", freq_map=None, polarized=True, want_cm=True):
        """RskSpectrum for the stored model; ``keep`` collects tensors that must outlive the call."""
        s = _lib.RskSpectrum()
        s.mode = mode
        s.stokes = self.stokes.data_ptr()
        s.nf_in = self.Nfreqs
        s.n = self.Ncomponents
        if mode == _lib.SPECTRAL_INDEX:
            s.spectral_index = self.spectral_index.data_ptr()
            s.ref_freq = self.reference_frequency.data_ptr()
        if mode == _lib.SUBBAND and spectral.is_step_kind(freq_interp_kind):
            # channel-picking kinds ('nearest', 'previous', ...): a degree-0 dense operator, nothing to precompute
            op = self._step_operator(freq_interp_kind)
            keep.append(op)
            s.degree = 0
            s.breaks = op["breaks"].data_ptr()
            s.nb = op["breaks"].numel() - 1
            s.ppoly = op["ppoly"].data_ptr()
        elif mode == _lib.SUBBAND:
            op = self._subband_operator(freq_interp_kind)
            keep.append(op)
            s.degree = op["degree"]
            s.breaks = op["breaks"].data_ptr()
            s.nb = op["breaks"].numel() - 1
            s.bcoef = op["bcoef"].data_ptr()
            s.nc = op["nc"]
            s.first = op["first"].data_ptr()
            s.pband = op["pband"].data_ptr()
        if mode == _lib.FULL and polarized and want_cm:
            cm = self._stokes_component_major()
            if cm is not None:
                keep.append(cm)
                s.stokes_cm = cm.data_ptr()
        if mode == _lib.FULL and freq_map is not None:
            fm = torch.as_tensor(np.ascontiguousarray(freq_map, dtype=np.int32)).to(self.device)
            keep.append(fm)
            s.freq_map = fm.data_ptr()
        if polarized and self._n_polarized > 0:
            s.polarized = self._polarized_mask.data_ptr()
        return s

    def _stokes_component_major(self):
        """Component-major copy [N][Nfreqs][4] of a stored ('full' / single-channel) Stokes cube for the coherency
        kernels (RskSpectrum.stokes_cm): made once per catalogue, dropped when the Stokes values change.  None when
        the copy would not fit comfortably in free HBM, for single-channel cubes whose (4, 1, N) rows are read once
        anyway, or with RSK_FULL_CM=0."""
        if os.environ.get("RSK_FULL_CM", "1") == "0" or self.Nfreqs < 2 or self.Ncomponents == 0:
            return None
        key = (self.stokes.data_ptr(), self.stokes._version, self.__dict__.get("_stokes_gen", 0), tuple(self.stokes.shape))
        hit = self.__dict__.get("_cm_cache")
        if hit is not None and hit[0] == key:
            return hit[1]
        nbytes = self.stokes.numel() * 8
        free, _ = torch.cuda.mem_get_info(self.device)
        if hit is None and nbytes > 0.35 * free:
            self._cm_cache = (key, None)
            return None
        self.__dict__.pop("_cm_cache", None)
        cm = torch.empty((self.Ncomponents, self.Nfreqs, 4), dtype=torch.float64, device=self.device)
        self._call("rsk_stokes_component_major", _lib.ptr(self.stokes), self.Nfreqs, self.Ncomponents, _lib.ptr(cm))
        self._cm_cache = (key, cm)
        return cm

    def _step_operator(self, freq_interp_kind):
        """Device copy of ``spectral.build_step_operator`` (depends on freq_array and kind only)."""
        key = (str(freq_interp_kind), self.freq_array.tobytes())
        op = self.__dict__.get("_step_cache")
        if op is None or op["key"] != key:
            _, breaks, ppoly = spectral.build_step_operator(self.freq_array, freq_interp_kind)
            op = dict(key=key, breaks=torch.as_tensor(breaks).to(self.device),
                      ppoly=torch.as_tensor(np.ascontiguousarray(ppoly)).to(self.device))
            self._step_cache = op
        return op

    def _subband_operator(self, freq_interp_kind):
        """Device-resident banded spline operator + the spline coefficients of every stored spectrum.

        Both are time independent, so they are built once per (stokes contents, freq_array, kind) and
        reused by every ``coherency_calc`` / ``at_frequencies`` call (the reference rebuilds its
        ``interp1d`` on every call, skymodel.py:2362-2368)."""
        key = (self.stokes.data_ptr(), self.stokes._version, self.__dict__.get("_stokes_gen", 0),
               tuple(self.stokes.shape), str(freq_interp_kind), self.freq_array.tobytes())
        op = self.__dict__.get("_subband_cache")
        if op is not None and op["key"] == key:
            return op
        degree, breaks, cmat, first, pband = spectral.build_bspline_operator(self.freq_array, freq_interp_kind)
        dev = self.device
        cm = torch.as_tensor(cmat).to(dev)
        nc = int(cmat.shape[0])
        bcoef = torch.empty((4, nc, self.Ncomponents), dtype=torch.float64, device=dev)
        self._call("rsk_spline_coefficients", _lib.ptr(self.stokes), self.Nfreqs, self.Ncomponents, _lib.ptr(cm), nc,
                   _lib.ptr(bcoef))
        op = dict(key=key, degree=degree, nc=nc, bcoef=bcoef, breaks=torch.as_tensor(breaks).to(dev),
                  first=torch.as_tensor(np.ascontiguousarray(first, dtype=np.int32)).to(dev),
                  pband=torch.as_tensor(np.ascontiguousarray(pband)).to(dev))
        self._subband_cache = op
        return op

    _FRAME_COMPONENT_NAMES = {"icrs": ("ra", "dec"), "fk5": ("ra", "dec"), "galactic": ("l", "b")}

    def __getattr__(self, name):
        """Handle references to frame coordinates, ``ra`` / ``dec`` or ``l`` / ``b`` depending on the frame
        (skymodel.py:1065-1090): radians, CUDA tensors.  Other unknown names raise AttributeError as usual."""
        if not name.startswith("_"):
            frame = self.__dict__.get("frame")
            comp_names = self._FRAME_COMPONENT_NAMES.get(frame, ())
            if name in comp_names:
                if self.__dict__.get("component_type") == "healpix":
                    warnings.warn("It is more efficient to use the `get_lon_lat` method to get "
                                  "longitudinal and latitudinal coordinates for HEALPix maps.")
                    return self.get_lon_lat()[comp_names.index(name)]
                return self.__dict__["lon" if comp_names.index(name) == 0 else "lat"]
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")

    @property
    def ncomponent_length_params(self):
        """Names of the parameters that have an Ncomponents-long axis and are set (skymodel.py:1267-1277)."""
        names = ["name", "stokes", "lon", "lat", "hpx_inds", "spectral_index", "reference_frequency", "stokes_error",
                 "beam_amp", "extended_model_group", "extra_columns", "alt_az", "pos_lmn", "above_horizon",
                 "frame_coherency"]
        return [n for n in names if self.__dict__.get(n) is not None]

    def check(self, check_extra=True, run_check_acceptability=True):
        """Check that the parameters are set reasonably (skymodel.py:1184-1277): the SkyModel-specific consistency rules
        and, with ``run_check_acceptability``, the reference's three warnings about NaN spectral indices, NaN Stokes
        values and negative Stokes I.  (The generic UVBase shape / type validation is not reproduced: shapes are
        enforced by the constructor here.)  Not called implicitly -- the hot path never pays for it."""
        if self.freq_array is not None and self.reference_frequency is not None:
            raise ValueError("Only one of freq_array and reference_frequency can be specified, not both.")
        allowed_units = ("Jy", "K sr") if self.component_type == "point" else ("Jy / sr", "K")
        if self.stokes_unit not in allowed_units:
            raise ValueError(
                f"For {self.component_type} component types, the stokes parameter must have a unit that can be "
                f"converted to {allowed_units}. Currently units are {self.stokes_unit}"
            )
        for attr in ("stokes_error", "beam_amp"):
            t = getattr(self, attr)
            if check_extra and t is not None and tuple(t.shape) != tuple(self.stokes.shape):
                raise ValueError(f"{attr} must have the shape of stokes, {tuple(self.stokes.shape)}; it is {tuple(t.shape)}")
        if run_check_acceptability:
            if self.spectral_type == "spectral_index" and self._has_nan("spectral_index"):
                warnings.warn(
                    "Some spectral index values are NaNs. If this is a GLEAM-based "
                    "sky model, consider using the 'subband' spectral type to"
                    "avoid this error (GLEAM assigns NaN spectral indices for "
                    "sources that are not well-fit by a power law, this can include "
                    "bright sources)."
                )
            if self._has_nan("stokes"):
                warnings.warn(
                    "Some Stokes values are NaNs. Use the select method with the"
                    "'non_nan' parameter to remove sources with NaN values at "
                    "any or all frequencies."
                )
            if self._count(self.stokes[0], _lib.COUNT_NEGATIVE_F64) > 0:
                warnings.warn(
                    "Some Stokes I values are negative. Use the select method "
                    "with the 'non_negative' parameter to remove sources with "
                    "negative Stokes I values."
                )
        return True

    def __eq__(self, other):
        """Equality of every parameter, as UVBase.__eq__ gives the reference's object (arrays to the UVParameter default
        tolerances rtol 1e-5 / atol 1e-8, angles to pyuvdata's 1 mas (RADIAN_TOL), NaNs equal to NaNs, histories up to
        white space)."""
        if not isinstance(other, SkyModel):
            return False
        for attr in ("component_type", "spectral_type", "Ncomponents", "Nfreqs", "stokes_unit", "frame", "equinox", "nside",
                     "hpx_order", "time", "telescope_location"):
            if getattr(self, attr) != getattr(other, attr):
                return False
        if not self._histories_match(self.history, other.history):
            return False
        f1 = None if self.filename is None else list(np.atleast_1d(self.filename))
        f2 = None if other.filename is None else list(np.atleast_1d(other.filename))
        if f1 != f2:
            return False
        for attr in ("freq_array", "freq_edge_array", "name", "extended_model_group"):
            x, y = getattr(self, attr), getattr(other, attr)
            if (x is None) != (y is None):
                return False
            if x is not None:
                x, y = np.asarray(x), np.asarray(y)
                if x.shape != y.shape:
                    return False
                if x.dtype.kind in "fc":
                    if not np.allclose(x, y, rtol=1e-5, atol=1e-8, equal_nan=True):
                        return False
                elif not np.array_equal(x, y):
                    return False
        if (self.extra_columns is None) != (other.extra_columns is None):
            return False
        if self.extra_columns is not None:
            if set(self.extra_columns) != set(other.extra_columns):
                return False
            for k in self.extra_columns:
                if not np.array_equal(np.asarray(self.extra_columns[k]), np.asarray(other.extra_columns[k])):
                    return False
        angles = {"alt_az"}
        lon1, lat1 = self.get_lon_lat()
        lon2, lat2 = other.get_lon_lat()
        pairs = [("lon", lon1, lon2), ("lat", lat1, lat2)]
        for attr in ("stokes", "spectral_index", "reference_frequency", "stokes_error", "beam_amp", "hpx_inds", "alt_az",
                     "pos_lmn", "above_horizon", "frame_coherency"):
            pairs.append((attr, getattr(self, attr), getattr(other, attr)))
        if self.component_type == "healpix" and (self.hpx_inds is None) != (other.hpx_inds is None):
            n = self.Ncomponents  # an implicit pixel range equals the same pixels listed explicitly
            pairs = [p for p in pairs if p[0] != "hpx_inds"]
            a = self.hpx_inds if self.hpx_inds is not None else torch.arange(self.hpx_first, self.hpx_first + n, device=self.device)
            b = other.hpx_inds if other.hpx_inds is not None else torch.arange(other.hpx_first, other.hpx_first + n, device=other.device)
            pairs.append(("hpx_inds", a, b))
        elif self.component_type == "healpix" and self.hpx_inds is None and self.hpx_first != other.hpx_first:
            return False
        for attr, x, y in pairs:
            if (x is None) != (y is None):
                return False
            if x is None:
                continue
            if x.shape != y.shape:
                return False
            y = y.to(x.device)
            if x.dtype in (torch.bool, torch.int64, torch.uint8):
                if not torch.equal(x, y):
                    return False
            elif attr in angles or attr in ("lon", "lat"):
                if not torch.allclose(x, y, rtol=0.0, atol=4.848136811095360e-09, equal_nan=True):
                    return False
            elif not torch.allclose(x, y, rtol=1e-5, atol=1e-8, equal_nan=True):
                return False
        return True

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = object.__hash__

    def copy(self):
        """Deep copy (device tensors are cloned)."""
        new = _copy.copy(self)
        for k, v in self.__dict__.items():
            if isinstance(v, torch.Tensor):
                setattr(new, k, v.clone())
            elif isinstance(v, np.ndarray):
                setattr(new, k, v.copy())
        new._freq_dev_cache = {}
        for name in ("_subband_cache", "_nan_group_cache", "_nan_cache", "_cm_cache", "_staging", "_graph",
                     "_copy_stream", "_timing_events", "_overlap", "_compacted"):
            new.__dict__.pop(name, None)
        if new.__dict__.get("_rot_key") is not None and new.alt_az is not None:
            new._rot_key = (new.alt_az.data_ptr(), new.alt_az._version)  # the cloned coefficients go with the cloned alt_az
        return new

    # ------------------------------------------------------------------------------------------
    # file I/O (SkyH5): skymodel.py:3491-3774, :4629-4868, :4871-5029
    # ------------------------------------------------------------------------------------------
    @classmethod
    def from_skyh5(cls, filename, skip_params=False, device=None, nthreads=0, **_checks):
        """Create a :class:`SkyModel` from a SkyH5 file (skymodel.py:3768-3774).

        The Stokes cube is streamed from the file into HBM by librsk (include/rsk_io.h); h5py is not
        used.  ``run_check`` / ``check_extra`` / ``run_check_acceptability`` are accepted and ignored
        (the UVBase validation framework is out of scope)."""
        from . import skyh5

        _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("pyradiosky_b200 needs a CUDA device; there is no CPU fallback.")
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        return cls(**skyh5.read_skyh5_params(filename, dev, skip_params=skip_params, nthreads=nthreads), device=dev)

    def read_skyh5(self, filename, skip_params=False, nthreads=0, **_checks):
        """Read a SkyH5 file into this object (skymodel.py:3491-3766)."""
        from . import skyh5

        self.__init__(**skyh5.read_skyh5_params(filename, self.device, skip_params=skip_params, nthreads=nthreads),
                      device=self.device)

    # ------------------------------------------------------------------------------------------
    # VOTable catalogues (skymodel.py:3779-4203): host-side parsing, arrays to the GPU through __init__
    # ------------------------------------------------------------------------------------------
    def read_votable_catalog(self, votable_file, table_name, id_column, lon_column, lat_column, flux_columns, frame,
                             reference_frequency=None, freq_array=None, freq_edge_array=None, spectral_index_column=None,
                             flux_error_columns=None, history="", run_check=True, check_extra=True,
                             run_check_acceptability=True):
        """Read a votable catalog file into this object (skymodel.py:3779-4025).

        Same column matching (a column is named by a part of its name that matches exactly one column), spectral-type
        rules, unit checks, warnings and errors as the reference; the file is parsed by ``pyradiosky_b200.votable``
        (TABLEDATA serialisation, no astropy).  ``frame``: 'icrs', 'fk5' or 'galactic'; frequencies in Hz or
        Quantity-like."""
        from . import votable as vot
        from .skyh5 import get_freq_edges_from_centers

        if isinstance(flux_columns, str):
            flux_columns = [flux_columns]
        if isinstance(flux_error_columns, str):
            flux_error_columns = [flux_error_columns]
        wanted = [id_column, lon_column, lat_column, *flux_columns, *(flux_error_columns or [])]
        if spectral_index_column is not None:
            wanted.append(spectral_index_column)
        wanted = [w.casefold() for w in wanted]
        # only the columns that can match are kept while the file streams by (the full GLEAM table has ~300)
        tables = vot.parse(votable_file, columns=lambda name: any(w in name.casefold() for w in wanted))
        if any(t.ID is None for t in tables):
            raise ValueError(f"File {votable_file} contains tables with no name or ID.")
        try:
            use = vot.get_matching_fields(table_name, [t.ID for t in tables])
            table = [t for t in tables if t.ID == use][0]
        except ValueError:
            use = vot.get_matching_fields(table_name, [t.name for t in tables])
            table = [t for t in tables if t.name == use][0]
        cols = table.colnames
        id_col = vot.get_matching_fields(id_column, cols)
        lon_col = vot.get_matching_fields(lon_column, cols, exclude_start_pattern="_")
        lat_col = vot.get_matching_fields(lat_column, cols, exclude_start_pattern="_")
        flux_cols = [vot.get_matching_fields(c, cols) for c in flux_columns]
        freq_array = _to_hz_host(freq_array)
        freq_edge_array = None if freq_edge_array is None else np.asarray(_to_hz_host(freq_edge_array)).reshape(2, -1)
        if len(flux_columns) > 1 and freq_array is None and freq_edge_array is None:
            raise ValueError(
                "Frequency information must be provided with multiple flux columns. "
                "Must provide either freq_edge_array or freq_array (if the "
                "frequencies are evenly spaced), both can be provided."
            )
        if len(flux_columns) > 1:
            if freq_edge_array is None:
                try:
                    freq_edge_array = get_freq_edges_from_centers(freq_array, atol=1.0)
                    warnings.warn("freq_edge_array not set, calculating it from the freq_array.")
                except ValueError as ve:
                    raise ValueError("freq_edge_array must be provided for multiple flux columns if "
                                     "freq_array is not regularly spaced.") from ve
            elif freq_array is None:
                warnings.warn("freq_array not set, calculating it from the freq_edge_array.")
                freq_array = np.mean(freq_edge_array, axis=0)
        n = len(table)
        spectral_index = None
        if reference_frequency is not None or len(flux_cols) == 1:
            if reference_frequency is not None:
                rf = reference_frequency.to_value("Hz") if hasattr(reference_frequency, "to_value") else reference_frequency
                reference_frequency = np.full(n, float(np.asarray(rf).ravel()[0]))
            if spectral_index_column is not None:
                spectral_type = "spectral_index"
                spectral_index = table.columns[vot.get_matching_fields(spectral_index_column, cols)]
            else:
                spectral_type = "flat"
        else:
            spectral_type = "subband"

        def flux_unit(col):
            return vot.FLUX_UNITS.get((table.units[col] or "").strip(), (None, None))

        kinds = {flux_unit(c)[0] for c in flux_cols}
        if len(kinds) != 1 or None in kinds:
            raise ValueError("All flux columns must have compatible units and must be compatible "
                             "with one of ['Jy', 'Jy/sr', 'K', 'K sr'].")
        unit_use = kinds.pop()
        stokes = np.zeros((4, len(flux_cols), n))
        for k, c in enumerate(flux_cols):
            stokes[0, k] = table.columns[c] * flux_unit(c)[1]
        stokes_error = None
        if flux_error_columns is not None:
            err_cols = [vot.get_matching_fields(c, cols) for c in flux_error_columns]
            if any(flux_unit(c)[0] != unit_use for c in err_cols):
                raise ValueError("All flux error columns must have units compatible with the units "
                                 "of the flux columns.")
            stokes_error = np.zeros((4, len(err_cols), n))
            for k, c in enumerate(err_cols):
                stokes_error[0, k] = table.columns[c] * flux_unit(c)[1]
        angle = {}
        for what, c in (("lon", lon_col), ("lat", lat_col)):
            u = (table.units[c] or "deg").strip()
            if u not in vot.ANGLE_UNITS:
                raise ValueError(f"column {c} has unit {u!r}, which is not an angle")
            angle[what] = table.columns[c] * vot.ANGLE_UNITS[u]
        self.__init__(name=table.columns[id_col].astype(str), lon=np.mod(angle["lon"], 2 * np.pi), lat=angle["lat"],
                      frame=str(getattr(frame, "name", frame)).lower(), stokes=stokes, stokes_unit=unit_use,
                      spectral_type=spectral_type, freq_array=freq_array if spectral_type == "subband" else None,
                      freq_edge_array=freq_edge_array if spectral_type == "subband" else None,
                      reference_frequency=reference_frequency, spectral_index=spectral_index, stokes_error=stokes_error,
                      history=history, filename=[os.path.basename(os.fspath(votable_file))],
                      device=self.__dict__.get("device"))

    @classmethod
    def from_votable_catalog(cls, votable_file, *args, device=None, **kwargs):
        """Create a new :class:`SkyModel` from a votable catalog (skymodel.py:4027-4031)."""
        sm = cls.__new__(cls)
        sm.device = device
        sm.read_votable_catalog(votable_file, *args, **kwargs)
        return sm

    def read_gleam_catalog(self, gleam_file, spectral_type="subband", with_error=False, use_paper_freqs=False,
                           run_check=True, check_extra=True, run_check_acceptability=True):
        """Read the GLEAM votable catalog file into this object (skymodel.py:4033-4195; GLEAM EGC catalog, version 2).

        ``spectral_type``: 'flat' (wide-band integrated flux, 200 MHz), 'spectral_index' (fitted flux at 200 MHz and
        ``alpha``; NaN where the fit failed) or 'subband' (the 20 sub-band fluxes with the catalogue's band centres and
        edges, or the 7.68 MHz-spaced ones with ``use_paper_freqs``); ``with_error`` adds the fitting errors as
        ``stokes_error``."""
        spec_type_list = ["flat", "spectral_index", "subband"]
        if spectral_type not in spec_type_list:
            raise ValueError(f"spectral_type {spectral_type} is not an allowed type. Allowed types are: {spec_type_list}")
        freq_array = freq_edge_array = reference_frequency = spectral_index_column = None
        if spectral_type == "flat":
            flux_columns, flux_error_columns, reference_frequency = "Fintwide", "e_Fintwide", 200e6
        elif spectral_type == "spectral_index":
            flux_columns, flux_error_columns, reference_frequency = "Fintfit200", "e_Fintfit200", 200e6
            spectral_index_column = "alpha"
        else:
            bands = ["076", "084", "092", "099", "107", "115", "122", "130", "143", "151", "158", "166", "174", "181",
                     "189", "197", "204", "212", "220", "227"]
            flux_columns = ["Fint" + b for b in bands]
            flux_error_columns = ["e_Fint" + b for b in bands]
            if use_paper_freqs:
                # 24 fine channels of 1.28 MHz in each of 5 coarse bands, four sub-bands of six channels each (:4129-4139)
                starts = np.array([57, 81, 109, 133, 157])
                chans = np.arange(24)[np.newaxis, :] + starts[:, np.newaxis]
                temp = np.reshape(chans, (5, 4, 6)) * 1.28 * 1e6
                freq_array = np.reshape(np.mean(temp, axis=2), 20)
                freq_lower = np.reshape(np.min(temp, axis=2), 20)
                freq_upper = np.reshape(np.max(temp, axis=2), 20)
            else:
                freq_array = np.array([76, 84, 92, 99, 107, 115, 122, 130, 143, 151, 158, 166, 174, 181, 189, 197, 204,
                                       212, 220, 227]) * 1e6
                freq_lower = np.array([72, 80, 88, 95, 103, 111, 118, 126, 139, 147, 154, 162, 170, 177, 185, 193, 200,
                                       208, 216, 223]) * 1e6
                freq_upper = np.array([80, 88, 95, 103, 111, 118, 126, 134, 147, 154, 162, 170, 177, 185, 193, 200, 208,
                                       216, 223, 231]) * 1e6
            freq_edge_array = np.stack([freq_lower, freq_upper])
        if not with_error:
            flux_error_columns = None
        self.read_votable_catalog(gleam_file, "GLEAM", "GLEAM", "RAJ2000", "DEJ2000", frame="fk5",
                                  flux_columns=flux_columns, freq_array=freq_array, freq_edge_array=freq_edge_array,
                                  reference_frequency=reference_frequency, spectral_index_column=spectral_index_column,
                                  flux_error_columns=flux_error_columns)

    @classmethod
    def from_gleam_catalog(cls, gleam_file, device=None, **kwargs):
        """Create a :class:`SkyModel` from a GLEAM catalog (skymodel.py:4197-4203)."""
        sm = cls.__new__(cls)
        sm.device = device
        sm.read_gleam_catalog(gleam_file, **kwargs)
        return sm

    # ------------------------------------------------------------------------------------------
    # text catalogues (skymodel.py:4204-4413, :5031-5182)
    # ------------------------------------------------------------------------------------------
    def read_text_catalog(self, catalog_csv, run_check=True, check_extra=True, run_check_acceptability=True):
        """Read a text file of sources into this object (skymodel.py:4204-4407): columns ``source_id``,
        ``<lon>_<frame>``, ``<lat>_<frame>`` in degrees, then ``Flux [Jy]`` (+ ``Frequency`` [Hz], ``Spectral_Index``) or one
        ``Flux_<f>_<unit>`` column per frequency, each optionally followed by its ``Flux_error`` column."""
        from . import textcat

        kw = textcat.read(catalog_csv)
        self.__init__(**kw, filename=[os.path.basename(os.fspath(catalog_csv))], history="",
                      device=self.__dict__.get("device"))

    @classmethod
    def from_text_catalog(cls, catalog_csv, device=None, **kwargs):
        """Create a :class:`SkyModel` from a text catalog (skymodel.py:4409-4413)."""
        sm = cls.__new__(cls)
        sm.device = device
        sm.read_text_catalog(catalog_csv, **kwargs)
        return sm

    def write_text_catalog(self, filename):
        """Write out this object to a text file (skymodel.py:5031-5182).  Text files hold point sources in Jy without
        subbands or extended-model groups; readable with :meth:`read_text_catalog`."""
        from . import textcat

        if self.component_type != "point":
            raise ValueError("component_type must be 'point' to use this method.")
        if self.stokes_unit != "Jy":
            raise ValueError("Stokes units must be equivalent to Jy to use this method.")
        if self.spectral_type == "subband":
            raise ValueError(
                "Text files do not support subband types, use write_skyh5. If you "
                "really need to get this into a text file, you could convert this "
                "to a 'full' spectral type (losing the frequency edge array "
                "information)."
            )
        if self.extended_model_group is not None:
            raise ValueError(
                "Text files do not support catalogs with extended_model_group, "
                "use write_skyh5. If you really need to get this into a text file, "
                "you could remove the extended_model_group information."
            )
        if self.name is None:
            raise ValueError("component names are required to write a text catalog.")
        textcat.write(filename, self)

    def read_fhd_catalog(self, filename_sav, expand_extended=True, extra_columns=None, run_check=True, check_extra=True,
                         run_check_acceptability=True):
        """Read in an FHD style catalog file (IDL save file, skymodel.py:4415-4621).  ``expand_extended``: replace
        extended sources by their components (names ``<id>_<k>``, ``extended_model_group`` = the parent's id);
        ``extra_columns``: {catalogue column: attribute name} to carry along."""
        from . import fhdcat

        self.__init__(**fhdcat.read(filename_sav, expand_extended=expand_extended, extra_columns=extra_columns), history="",
                      device=self.__dict__.get("device"))

    @classmethod
    def from_fhd_catalog(cls, filename_sav, device=None, **kwargs):
        """Create a :class:`SkyModel` from an FHD catalog (skymodel.py:4623-4627)."""
        sm = cls.__new__(cls)
        sm.device = device
        sm.read_fhd_catalog(filename_sav, **kwargs)
        return sm

    @classmethod
    def from_file(cls, filename, filetype=None, spectral_type=None, with_error=False, use_paper_freqs=False,
                  table_name=None, id_column=None, lon_column=None, lat_column=None, frame=None, flux_columns=None,
                  reference_frequency=None, freq_array=None, freq_edge_array=None, spectral_index_column=None,
                  flux_error_columns=None, history="", expand_extended=True, extra_columns=None, device=None, **kwargs):
        """``SkyModel.from_file`` (skymodel.py:4629-4868) for every format of the reference: 'skyh5', 'gleam', 'vot', 'text', 'fhd'."""
        allowed_filetypes = ["skyh5", "gleam", "vot", "text", "fhd"]
        name = os.fspath(filename)
        if filetype is not None:
            if filetype not in allowed_filetypes:
                raise ValueError(f"Invalid filetype. Filetype options are: {allowed_filetypes}")
        else:
            ext = os.path.splitext(name)[1]
            if ext == ".vot":
                filetype = "gleam" if "gleam" in name.casefold() else "vot"
            else:
                filetype = {".skyh5": "skyh5", ".txt": "text", ".sav": "fhd"}.get(ext)
            if filetype is None:
                raise ValueError("Cannot determine the file type.")
        if filetype == "skyh5":
            return cls.from_skyh5(filename, device=device, **kwargs)
        if filetype == "gleam":
            return cls.from_gleam_catalog(filename, device=device, spectral_type=spectral_type or "subband",
                                          with_error=with_error, use_paper_freqs=use_paper_freqs)
        if filetype == "vot":
            required = {"table_name": table_name, "id_column": id_column, "lon_column": lon_column,
                        "lat_column": lat_column, "flux_columns": flux_columns, "frame": frame}
            for key, val in required.items():
                if val is None:
                    raise ValueError(f"{key} is required when reading vot files.")
            return cls.from_votable_catalog(filename, table_name, id_column, lon_column, lat_column, flux_columns, frame,
                                            device=device, reference_frequency=reference_frequency,
                                            freq_array=freq_array, freq_edge_array=freq_edge_array,
                                            spectral_index_column=spectral_index_column,
                                            flux_error_columns=flux_error_columns, history=history)
        if filetype == "text":
            return cls.from_text_catalog(filename, device=device)
        return cls.from_fhd_catalog(filename, device=device, expand_extended=expand_extended, extra_columns=extra_columns)

    def read(self, filename, **kwargs):
        """Read a file of any supported type into this object (skymodel.py:4629-4862); same arguments as
        :meth:`from_file`."""
        kwargs.setdefault("device", self.__dict__.get("device"))
        new = type(self).from_file(filename, **kwargs)
        self.__dict__.clear()
        self.__dict__.update(new.__dict__)

    def write_skyh5(self, filename, clobber=False, data_compression=None, **_checks):
        """Write this object to a SkyH5 file (skymodel.py:4871-5029)."""
        from . import skyh5

        skyh5.write_skyh5(self, filename, clobber=clobber, data_compression=data_compression)

    # ------------------------------------------------------------------------------------------
    # HEALPix <-> point conversions: skymodel.py:1675-2081
    # ------------------------------------------------------------------------------------------
    def _pixel_area(self, nside):
        """astropy_healpix.nside_to_pixel_area(nside) in steradian."""
        return 4 * math.pi / (12 * int(nside) * int(nside))

    def _stokes_changed(self):
        """The Stokes values were modified in place by a librsk kernel (which torch's ``_version`` counter does not
        see): drop everything derived from them -- spline coefficients, NaN-pattern groups, NaN flags."""
        self._stokes_gen = self.__dict__.get("_stokes_gen", 0) + 1
        for name in ("_subband_cache", "_nan_group_cache", "_nan_cache", "_cm_cache"):
            self.__dict__.pop(name, None)

    def _scale(self, tensor, factor, divide):
        self._call("rsk_scale", _lib.ptr(tensor), tensor.numel(), ctypes.c_double(factor), 1 if divide else 0)
        if tensor is self.stokes:
            self._stokes_changed()

    def _unit_times_sr(self, unit, divide):
        table = {("K", False): "K sr", ("Jy / sr", False): "Jy", ("K sr", True): "K", ("Jy", True): "Jy / sr"}
        try:
            return table[(unit, divide)]
        except KeyError:
            raise ValueError(f"cannot {'divide' if divide else 'multiply'} Stokes in {unit!r} by a solid angle") from None

    def healpix_to_point(self, to_jy=True, **_checks):
        """Convert a healpix object to point components: multiply by the pixel area, optionally convert
        to Jy (skymodel.py:1675-1736).  ``nside``, ``hpx_order`` and ``hpx_inds`` stay set, as upstream,
        so :meth:`_point_to_healpix` can undo it."""
        if self.component_type != "healpix":
            raise ValueError("This method can only be called if component_type is 'healpix'.")
        lon, lat = self.get_lon_lat()
        if self.hpx_inds is None:
            self.hpx_inds = torch.arange(self.hpx_first, self.hpx_first + self.Ncomponents, dtype=torch.int64,
                                         device=self.device)
            self.hpx_first = 0
        self.lon, self.lat = lon, lat
        self.hpx_frame = None
        self.component_type = "point"
        area = self._pixel_area(self.nside)
        self._scale(self.stokes, area, divide=False)
        self.stokes_unit = self._unit_times_sr(self.stokes_unit, divide=False)
        if self.frame_coherency is not None:
            self.frame_coherency = self.frame_coherency * area
        inds = self.hpx_inds.cpu().numpy()
        self.name = np.array(["nside" + str(self.nside) + "_" + self.hpx_order + "_" + str(ind) for ind in inds])
        if to_jy:
            self.kelvin_to_jansky()

    def _point_to_healpix(self, to_k=True, **_checks):
        """Undo :meth:`healpix_to_point` (skymodel.py:1738-1806): divide by the pixel area, optionally
        convert to K.  Does not assign general point components to a grid (see :meth:`assign_to_healpix`)."""
        if self.component_type != "point" or self.nside is None or self.hpx_inds is None or self.hpx_order is None:
            raise ValueError(
                "This method can only be called if component_type is 'point' and "
                "the nside, hpx_order and hpx_inds parameters are set."
            )
        self.component_type = "healpix"
        area = self._pixel_area(self.nside)
        self._scale(self.stokes, area, divide=True)
        self.stokes_unit = self._unit_times_sr(self.stokes_unit, divide=True)
        if self.frame_coherency is not None:
            self.frame_coherency = self.frame_coherency / area
        self.hpx_frame = self.frame
        self.name = None
        self.lon = None
        self.lat = None
        if to_k:
            self.jansky_to_kelvin()

    def assign_to_healpix(self, nside, order="ring", to_k=True, full_sky=False, sort=True, inplace=False, **_checks):
        """Assign point components to their nearest pixel of a HEALPix grid, summing the flux of components
        that share a pixel, dividing by the pixel area and optionally converting to K
        (skymodel.py:1808-2081).  Pixel lookup, per-pixel sums and the area scaling run in librsk kernels
        (``rsk_lonlat_to_healpix``, ``rsk_segment_sum``, ``rsk_scale``); torch only sorts the pixel indices.

        Differences from upstream: fluxes of a pixel are accumulated strictly in catalogue order (numpy
        switches to pairwise summation beyond 8 terms); the polarised-component bookkeeping is rebuilt for
        the new pixel set (upstream keeps the indices of the old catalogue)."""
        if self.component_type != "point":
            raise ValueError("This method can only be called if component_type is 'point'.")
        order = str(order).lower()
        if order not in ("ring", "nested"):
            raise ValueError("order must be 'ring' or 'nested'")
        nside = int(nside)
        n = self.Ncomponents
        pix = torch.empty(n, dtype=torch.int64, device=self.device)
        self._call("rsk_lonlat_to_healpix", nside, 1 if order == "nested" else 0, _lib.ptr(self.lon), _lib.ptr(self.lat),
                   n, _lib.ptr(pix))
        spix, perm = torch.sort(pix, stable=True)  # within a pixel the catalogue order is kept
        upix, counts = torch.unique_consecutive(spix, return_counts=True)
        nseg = int(upix.numel())
        start = torch.zeros(nseg + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(counts, 0, out=start[1:])
        first = perm[start[:-1]]  # first catalogue member of every pixel
        if nseg < n:
            # parameters that must not vary among the components of one pixel (skymodel.py:1929-1960)
            seg_of = torch.repeat_interleave(torch.arange(nseg, device=self.device), counts)
            for attr, label in (("reference_frequency", "reference_frequency"), ("spectral_index", "spectral_index")):
                vals = getattr(self, attr)
                if vals is not None and bool((vals[perm] != vals[first][seg_of]).any()):
                    raise ValueError(
                        "Multiple components map to a single healpix pixel "
                        f"and the {label} varies among them. Consider "
                        "using the `at_frequencies` method first or a "
                        "larger nside."
                    )
            if self.extended_model_group is not None:
                grp = np.asarray(self.extended_model_group)
                ph, fh, sh = perm.cpu().numpy(), first.cpu().numpy(), seg_of.cpu().numpy()
                if np.any(grp[ph] != grp[fh][sh]):
                    raise ValueError(
                        "Multiple components map to a single healpix pixel "
                        "and the extended_model_group varies among them."
                        "Consider using a larger nside."
                    )
            if self.beam_amp is not None and bool((self.beam_amp[:, :, perm] != self.beam_amp[:, :, first][:, :, seg_of]).any()):
                raise ValueError(
                    "Multiple components map to a single healpix pixel and "
                    "the beam_amp varies among them. "
                    "Consider using a larger nside."
                )
        nrows = 4 * self.Nfreqs
        area = self._pixel_area(nside)

        def per_pixel(values, quadrature):
            out = torch.empty((4, self.Nfreqs, nseg), dtype=torch.float64, device=self.device)
            self._call("rsk_segment_sum", _lib.ptr(values), nrows, n, _lib.ptr(perm), _lib.ptr(start), nseg,
                       1 if quadrature else 0, _lib.ptr(out))
            self._scale(out, area, divide=True)
            return out

        new_stokes = per_pixel(self.stokes, False)
        new_err = per_pixel(self.stokes_error, True) if self.stokes_error is not None else None
        ref_freq = None if self.reference_frequency is None else self.reference_frequency[first]
        sp_index = None if self.spectral_index is None else self.spectral_index[first]
        beam_amp = None if self.beam_amp is None else self.beam_amp[:, :, first].contiguous()
        ext = None if self.extended_model_group is None else np.asarray(self.extended_model_group)[first.cpu().numpy()]
        extra = None
        if self.extra_columns:
            fh = first.cpu().numpy()
            extra = {k: np.asarray(v)[fh] for k, v in self.extra_columns.items()}
        hpx_inds = upix
        if not sort:
            # upstream keeps the order of first appearance (dict insertion order, skymodel.py:1921-1962)
            appear = torch.argsort(first)
            hpx_inds, new_stokes = upix[appear], new_stokes[:, :, appear].contiguous()
            new_err = None if new_err is None else new_err[:, :, appear].contiguous()
            ref_freq = None if ref_freq is None else ref_freq[appear]
            sp_index = None if sp_index is None else sp_index[appear]
            beam_amp = None if beam_amp is None else beam_amp[:, :, appear].contiguous()
            ah = appear.cpu().numpy()
            ext = None if ext is None else ext[ah]
            extra = None if extra is None else {k: v[ah] for k, v in extra.items()}
        npix = 12 * nside * nside
        if full_sky and nseg < npix:
            # zero-flux pixels; spectral parameters take the median of the assigned ones (skymodel.py:2014-2066)
            def widen(t, fill=0.0):
                if t is None:
                    return None
                shape = tuple(t.shape[:-1]) + (npix,)
                full = torch.full(shape, fill, dtype=t.dtype, device=self.device)
                full[..., hpx_inds] = t
                return full

            new_stokes, new_err, beam_amp = widen(new_stokes), widen(new_err), widen(beam_amp)
            if ref_freq is not None:
                ref_freq = widen(ref_freq, float(np.median(ref_freq.cpu().numpy())))
            if sp_index is not None:
                sp_index = widen(sp_index, float(np.median(sp_index.cpu().numpy())))
            hh = hpx_inds.cpu().numpy()
            if ext is not None:
                full_ext = np.full(npix, "", dtype=ext.dtype)
                full_ext[hh] = ext
                ext = full_ext
            if extra is not None:
                wide = {}
                for k, v in extra.items():
                    w = np.zeros(npix, dtype=v.dtype)
                    w[hh] = v
                    wide[k] = w
                extra = wide
            if sort:
                hpx_inds = torch.arange(npix, dtype=torch.int64, device=self.device)
            else:  # assigned pixels first (in their order), then the empty ones
                member = torch.zeros(npix, dtype=torch.bool, device=self.device)
                member[hpx_inds] = True
                rest = torch.nonzero(~member, as_tuple=False).flatten()
                order_all = torch.cat([hpx_inds, rest])
                oh = order_all.cpu().numpy()
                new_stokes = new_stokes[:, :, order_all].contiguous()
                new_err = None if new_err is None else new_err[:, :, order_all].contiguous()
                beam_amp = None if beam_amp is None else beam_amp[:, :, order_all].contiguous()
                ref_freq = None if ref_freq is None else ref_freq[order_all]
                sp_index = None if sp_index is None else sp_index[order_all]
                ext = None if ext is None else ext[oh]
                extra = None if extra is None else {k: v[oh] for k, v in extra.items()}
                hpx_inds = order_all
        new = SkyModel(component_type="healpix", nside=nside, hpx_order=order, hpx_inds=hpx_inds, hpx_frame=self.frame,
                       equinox=self.equinox, stokes=new_stokes, spectral_type=self.spectral_type,
                       freq_array=self.freq_array, freq_edge_array=self.freq_edge_array, reference_frequency=ref_freq,
                       spectral_index=sp_index, stokes_unit=self._unit_times_sr(self.stokes_unit, divide=True),
                       device=self.device, history=self.history, filename=self.filename, extended_model_group=ext,
                       stokes_error=new_err, beam_amp=beam_amp, extra_columns=extra)
        if self.frame_coherency is not None:
            new.calc_frame_coherency()
        if to_k:
            new.jansky_to_kelvin()
        if inplace:
            self.__dict__.clear()
            self.__dict__.update(new.__dict__)
            return None
        return new

    # ------------------------------------------------------------------------------------------
    # coordinates
    # ------------------------------------------------------------------------------------------
    def get_lon_lat(self):
        """Longitudes and latitudes (radians, CUDA tensors) of the components in ``frame``.

        For healpix models the pixel centres are generated on the GPU (reference: astropy_healpix,
        skymodel.py:1656-1669)."""
        if self.component_type != "healpix":
            return self.lon, self.lat
        lon = torch.empty(self.Ncomponents, dtype=torch.float64, device=self.device)
        lat = torch.empty_like(lon)
        self._call("rsk_healpix_lonlat", self.nside, 1 if self.hpx_order == "nested" else 0,
                   _lib.ptr(self.hpx_inds), self.hpx_first, self.Ncomponents, _lib.ptr(lon), _lib.ptr(lat))
        return lon, lat

    def transform_to(self, frame, equinox=2000.0):
        """Transform to a different coordinate frame (skymodel.py:1419-1446; there a thin wrapper on
        ``SkyCoord.transform_to``): 'icrs', 'fk5' (with ``equinox``) or 'galactic'.  The coordinates are rotated on the
        GPU (rsk_rotate_lonlat); as in the reference the Stokes parameters are left as they are."""
        if self.component_type == "healpix":
            raise ValueError(
                "Direct coordinate transformation between frames is not valid "
                "for `healpix` type catalogs. Please use the "
                "`healpix_interp_transform` to transform to a new frame and "
                "interpolate to the new pixel centers. Alternatively, you can "
                "call `healpix_to_point` to convert the healpix map to a point "
                "source catalog before calling this function."
            )
        frame = getattr(frame, "name", frame)
        new_m = astrometry.frame_to_icrs_matrix(str(frame).lower(), float(equinox))
        rot = np.ascontiguousarray(new_m.T @ self._frame_to_icrs)
        lon = torch.empty_like(self.lon)
        lat = torch.empty_like(self.lat)
        self._call("rsk_rotate_lonlat", _lib.ptr(self.lon), _lib.ptr(self.lat), self.Ncomponents,
                   _lib.host_array(rot.ravel()), _lib.ptr(lon), _lib.ptr(lat))
        self.lon, self.lat = lon, lat
        self.frame = str(frame).lower()
        self.equinox = float(equinox)
        self._frame_to_icrs = new_m
        self._rot_key = None  # the cached polarisation-rotation coefficients were for the old frame's basis

    def healpix_interp_transform(self, frame, full_sky=False, inplace=True, run_check=True, check_extra=True,
                                 run_check_acceptability=True, equinox=2000.0):
        """Transform a HEALPix map to a new frame and interpolate to new pixel centers (skymodel.py:1448-1538).

        The pixel centres of the new frame's map are rotated into the old frame and the old map is interpolated
        bilinearly there (rsk_healpix_interp: what ``astropy_healpix.HEALPix.interpolate_bilinear_skycoord`` does at
        :1507-1509), for every frequency in one launch.  As upstream, only Stokes I is carried (polarised maps raise),
        a new pixel is kept when all four of its neighbours exist in the old map, and ``full_sky=True`` returns all
        12 nside^2 pixels (zeros where the interpolation had no support).  Frames: 'icrs', 'fk5' (``equinox``),
        'galactic'.  Per-pixel extras that cannot follow the resampling (``stokes_error``, ``beam_amp``, names, extra
        columns) and the time-dependent attributes are cleared."""
        this = self if inplace else self.copy()
        if this.component_type != "healpix":
            raise ValueError("Healpix frame interpolation is not valid for point source catalogs.")
        if bool((this.stokes[1:] != 0).any()):
            raise NotImplementedError(
                "Healpix map transformations are currently not implemented for "
                "catalogs with polarization information."
            )
        frame = str(getattr(frame, "name", frame)).lower()
        new_m = astrometry.frame_to_icrs_matrix(frame, float(equinox))
        rot = np.ascontiguousarray(this._frame_to_icrs.T @ new_m)  # new frame -> old frame
        npix = 12 * this.nside * this.nside
        n_in = this.Ncomponents
        lookup = None
        if this.hpx_inds is not None:
            lookup = torch.full((npix,), -1, dtype=torch.int64, device=this.device)
            lookup[this.hpx_inds] = torch.arange(n_in, dtype=torch.int64, device=this.device)
        out = torch.zeros((4, this.Nfreqs, npix), dtype=torch.float64, device=this.device)
        valid = torch.empty(npix, dtype=torch.uint8, device=this.device)
        this._call("rsk_healpix_interp", this.nside, 1 if this.hpx_order == "nested" else 0, _lib.host_array(rot.ravel()),
                   _lib.ptr(lookup), this.hpx_first, n_in, _lib.ptr(this.stokes[0].contiguous()), this.Nfreqs,
                   _lib.ptr(out[0]), _lib.ptr(valid))
        had_fc = this.frame_coherency is not None
        if full_sky:
            this.hpx_inds, this.hpx_first, m = None, 0, npix
            this.stokes = out
        else:
            idx, m = this._compact_mask(valid)
            if m == npix:
                this.hpx_inds, this.hpx_first = None, 0
                this.stokes = out
            else:
                idx = idx[:m].contiguous()
                this.hpx_inds, this.hpx_first = idx, 0
                this.stokes = this._gather(out, idx, m)
        this.Ncomponents = m
        this.frame = frame
        this.hpx_frame = frame
        this.equinox = float(equinox)
        this._frame_to_icrs = new_m
        this.name = None
        this.stokes_error = this.beam_amp = this.extended_model_group = this.extra_columns = None
        this.clear_time_position_specific_params()
        this._polarized_mask = torch.zeros(m, dtype=torch.uint8, device=this.device)
        this._n_polarized = 0
        for cache in ("_rot", "_subband_cache", "_nan_group_cache", "_nan_cache", "_cm_cache", "_compacted", "_staging", "_graph",
                      "_rise_lst", "_set_lst"):
            this.__dict__.pop(cache, None)
        this.frame_coherency = None
        if had_fc:
            this.calc_frame_coherency()  # recalculate the coherency now that we are in the new frame (:1528-1530)
        return None if inplace else this

    def _check_tel_location(self, telescope_location):
        # skymodel.py:2391-2411 (the lunarsky/MoonLocation branch is out of scope)
        if not astrometry.is_earth_location(telescope_location):
            raise ValueError(
                "telescope_location must be an :class:`astropy.EarthLocation` "
                f"object. value was: {str(telescope_location)}"
            )

    def add_extra_columns(self, *, names, values, dtype=None):
        """Add one or more length Ncomponent attributes to the object (skymodel.py:1279-1355).  ``extra_columns`` is a
        dict of 1-D host arrays here (a numpy recarray upstream); they follow select / concat and are stored in SkyH5."""
        if isinstance(names, str):
            names = [names]
        if isinstance(values, np.ndarray):
            values = [values]
        if len(names) != len(values):
            raise ValueError("Must provide the same number of names and values.")
        if dtype is not None:
            if isinstance(dtype, (list, tuple, np.ndarray)):
                if len(dtype) != len(names):
                    raise ValueError("If dtype is set, it must be the same length as `name`.")
            else:
                dtype = [dtype]
        if self.extra_columns is not None:
            for name in names:
                if name in self.extra_columns:
                    raise ValueError(f"{name} column already exists, use a different name.")
        values = [np.asarray(v) for v in values]
        for index, val in enumerate(values):
            if val.shape != (self.Ncomponents,):
                raise ValueError("value array(s) must be 1D, Ncomponents length array(s). "
                                 f"The value array in index {index} is not the right shape.")
        new = {name: (val if dtype is None else val.astype(dtype[k])) for k, (name, val) in enumerate(zip(names, values))}
        self.extra_columns = {**(self.extra_columns or {}), **new}

    def remove_extra_columns(self, names):
        """Remove one or more length Ncomponent attributes from the object (skymodel.py:1357-1376)."""
        if isinstance(names, str):
            names = [names]
        kept = {k: v for k, v in (self.extra_columns or {}).items() if k not in names}
        self.extra_columns = kept or None

    def clear_time_position_specific_params(self):
        """Set time/position dependent attributes to None (skymodel.py:1385-ff)."""
        self.time = None
        self.telescope_location = None
        self.alt_az = None
        self.pos_lmn = None
        self.above_horizon = None
        self._astrom = None
        self._rot_key = None

    def update_positions(self, time, telescope_location, astrom=None):
        """Calculate the altitude/azimuth positions for source components (skymodel.py:2413-2479).

        Updates ``alt_az`` (2, N) [alt; az] radians, ``pos_lmn`` (3, N), ``above_horizon`` (N,) bool,
        ``time`` and ``telescope_location`` in place; all arrays are CUDA tensors.

        ``astrom`` (extension): a ready 31-double eraASTROM for this (time, location), e.g. exported
        from a machine that has astropy; by default the context is built from ``time`` and
        ``telescope_location``.
        """
        if not astrometry.is_time(time):
            raise ValueError(f"time must be an astropy Time object. value was: {time}")
        self._check_tel_location(telescope_location)
        # Don't repeat calculations (skymodel.py:2439-2441)
        if self.time is not None and self.telescope_location is not None and astrom is None:
            if self.time == time and self.telescope_location == telescope_location:
                return
        self.time = time
        self.telescope_location = telescope_location
        if astrom is None:
            astrom = astrometry.astrom_context(time, telescope_location)
        astrom = np.ascontiguousarray(astrom, dtype=np.float64).ravel()
        if astrom.size != _lib.ASTROM_LEN:
            raise ValueError(f"astrom must hold {_lib.ASTROM_LEN} doubles (eraASTROM)")
        self._astrom = astrom
        n = self.Ncomponents
        if self.__dict__.get("_overlap") is not None:
            return self._update_positions_overlapped(astrom)
        if self.alt_az is None or self.alt_az.shape[1] != n:
            self.alt_az = torch.empty((2, n), dtype=torch.float64, device=self.device)
        if self.pos_lmn is None or self.pos_lmn.shape[1] != n:
            self.pos_lmn = torch.empty((3, n), dtype=torch.float64, device=self.device)
        above = torch.empty(n, dtype=torch.uint8, device=self.device)
        coords = self._coords_struct()
        ravg = None
        if self._n_polarized > 0:
            # K1 already holds every sine and cosine the polarisation rotation needs, so it also writes the
            # per-component rotation coefficients that coherency_calc will apply at this time step.
            if self.__dict__.get("_rot") is None or self._rot.shape[1] != n:
                self._rot = torch.empty((3, n), dtype=torch.float64, device=self.device)
            coords.rot = self._rot.data_ptr()
            ravg = _lib.host_array(astrometry.average_rotation_matrix(astrom).ravel())
        self._call("rsk_positions", ctypes.byref(coords), n, _lib.host_array(self._frame_to_icrs.ravel()),
                   _lib.host_array(astrom), ravg, _lib.ptr(self.alt_az), _lib.ptr(self.pos_lmn), _lib.ptr(above))
        self._rot_key = (self.alt_az.data_ptr(), self.alt_az._version) if ravg is not None else None
        self.above_horizon = above.view(torch.bool)

    def enable_position_overlap(self, enable=True):
        """Let :meth:`update_positions` run beside the previous time step's coherency kernel (extension).

        A time step is K1 (positions) -> K1b (compaction) -> K2 (coherency); K2 is >= 97 % of it and bound by HBM, the
        other four kernels are latency-sized.  With the overlap on, K1 and K1b of step k + 1 run on a side stream into
        the *other* of two buffer sets while K2 of step k is still streaming on the caller's stream; the caller's
        stream then waits (on the device) for the side stream, so anything queued after ``update_positions`` returns
        -- ``coherency_calc``, a read of ``alt_az`` -- sees the new positions.  Results are bit-identical; what
        changes is that ``alt_az`` / ``pos_lmn`` / ``above_horizon`` of step k are overwritten by step k + 2, not
        k + 1.  Worth 5-8 % of a step for shards of ~100k components (strong scaling over 8 GPUs), nothing for
        large ones."""
        if not enable:
            self.__dict__.pop("_overlap", None)
            return
        n, dev = self.Ncomponents, self.device
        nbytes = _lib.load().rsk_workspace_bytes(_lib.OP_COMPACT, n, 0)
        sets = []
        for _ in range(2):
            sets.append(dict(
                alt_az=torch.empty((2, n), dtype=torch.float64, device=dev),
                pos_lmn=torch.empty((3, n), dtype=torch.float64, device=dev),
                above=torch.empty(n, dtype=torch.uint8, device=dev),
                rot=torch.empty((3, n), dtype=torch.float64, device=dev) if self._n_polarized > 0 else None,
                idx=torch.empty(max(n, 1), dtype=torch.int64, device=dev),
                cnt=torch.empty(1, dtype=torch.int64, device=dev),
                work=torch.empty(max(nbytes, 8), dtype=torch.uint8, device=dev), work_bytes=nbytes, mark=None))
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))  # the model's own construction kernels
        for st in sets:
            for t in st.values():
                if isinstance(t, torch.Tensor):
                    t.record_stream(side)  # written on the side stream, read on the caller's
        self._overlap = dict(sets=sets, side=side, k=0)

    def _update_positions_overlapped(self, astrom):
        st = self._overlap
        n = self.Ncomponents
        if st["sets"][0]["alt_az"].shape[1] != n or (st["sets"][0]["rot"] is None) != (self._n_polarized == 0):
            self.enable_position_overlap()  # the component list changed (select / concat): new buffer sets
            st = self._overlap
        main = torch.cuda.current_stream(self.device)
        side = st["side"]
        cur, other = st["sets"][st["k"] % 2], st["sets"][(st["k"] + 1) % 2]
        st["k"] += 1
        # everything the caller queued up to here may still read the set that is current now ("other" after this
        # call): mark it, and make the kernels that will overwrite that set next time wait for the mark.  The set
        # written now was marked one call ago, i.e. before the previous coherency kernel was queued, so K1 below
        # does not wait for that kernel.
        mark = torch.cuda.Event()
        mark.record(main)
        other["mark"] = mark
        if cur["mark"] is not None:
            side.wait_event(cur["mark"])
        coords = self._coords_struct()
        ravg = None
        if cur["rot"] is not None:
            coords.rot = cur["rot"].data_ptr()
            ravg = _lib.host_array(astrometry.average_rotation_matrix(astrom).ravel())
        with torch.cuda.stream(side):
            self._call("rsk_positions", ctypes.byref(coords), n, _lib.host_array(self._frame_to_icrs.ravel()),
                       _lib.host_array(astrom), ravg, _lib.ptr(cur["alt_az"]), _lib.ptr(cur["pos_lmn"]),
                       _lib.ptr(cur["above"]))
            self._call("rsk_compact_index", _lib.ptr(cur["above"]), n, _lib.ptr(cur["idx"]), _lib.ptr(cur["cnt"]),
                       _lib.ptr(cur["work"]), ctypes.c_size_t(cur["work_bytes"]))
            ready = torch.cuda.Event()
            ready.record(side)
        main.wait_event(ready)  # device-side: later work on the caller's stream sees the new positions
        self.alt_az, self.pos_lmn = cur["alt_az"], cur["pos_lmn"]
        self.above_horizon = cur["above"].view(torch.bool)
        if cur["rot"] is not None:
            self._rot = cur["rot"]
        self._rot_key = (self.alt_az.data_ptr(), self.alt_az._version) if ravg is not None else None
        # the compaction that belongs to this mask (reused by _plan_coherency while the mask is this very tensor)
        self._compacted = (self.above_horizon, self.above_horizon._version, cur["idx"], cur["cnt"])

    def _positions_from_step(self, step_dev):
        """update_positions with the time-dependent record (astrometry.pack_step) read from device memory."""
        n = self.Ncomponents
        coords = self._coords_struct()
        if self._n_polarized > 0:
            coords.rot = self._rot.data_ptr()
        self._call("rsk_positions_step", ctypes.byref(coords), n, _lib.ptr(step_dev), _lib.ptr(self.alt_az),
                   _lib.ptr(self.pos_lmn), _lib.ptr(self.above_horizon.view(torch.uint8)))

    def capture_step(self, freqs=None, freq_interp_kind="cubic"):
        """Capture one whole time step -- positions, horizon compaction, fused ``coherency_calc(freqs)`` -- in a CUDA
        graph (extension).  Returns a :class:`StepGraph`; ``graph.run(time, telescope_location)`` replays it for a
        new time with one small host->device copy and one graph launch instead of five kernel launches and their
        Python set-up, which is what a step costs for small catalogues."""
        return StepGraph(self, freqs, freq_interp_kind)

    def sweep_contexts(self, times, telescope_location, astroms=None):
        """The device-ready contexts of :meth:`positions_sweep` for a list of times ((T, RSK_CTX_LEN) CUDA tensor):
        built and uploaded once, then passed as ``contexts=`` to any number of sweeps (e.g. one per catalogue shard)."""
        self._check_tel_location(telescope_location)
        if astroms is None:
            for t in times:
                if not astrometry.is_time(t):
                    raise ValueError(f"time must be an astropy Time object. value was: {t}")
            astroms = astrometry.astrom_contexts(times, telescope_location)
        astroms = np.ascontiguousarray(astroms, dtype=np.float64).reshape(-1, _lib.ASTROM_LEN)
        return torch.as_tensor(astrometry.pack_contexts(astroms, self._frame_to_icrs)).to(self.device)

    def positions_sweep(self, times, telescope_location, astroms=None, want_lmn=False, want_mask=True, contexts=None):
        """Alt/Az of every component at many times in one launch (extension; BASELINE config 5).

        Equivalent to calling :meth:`update_positions` for each element of ``times`` (as the
        reference's LST-sweep test does, tests/test_skymodel.py:2661-2740) without touching the
        object's state.  Returns ``(alt_az [T, 2, N], pos_lmn [T, 3, N] or None, above [T, N] bool or
        None)`` as CUDA tensors.  ``astroms`` optionally supplies ready (T, 31) eraASTROM rows; ``contexts`` the
        already uploaded result of :meth:`sweep_contexts` (then ``times`` / ``astroms`` are not looked at).
        """
        if contexts is not None:
            ctx_dev = contexts
            if (not isinstance(ctx_dev, torch.Tensor) or not ctx_dev.is_cuda or ctx_dev.dtype != torch.float64
                    or ctx_dev.dim() != 2 or ctx_dev.shape[1] != _lib.CTX_LEN or not ctx_dev.is_contiguous()):
                raise ValueError(f"contexts must be the (T, {_lib.CTX_LEN}) float64 CUDA tensor of sweep_contexts()")
        else:
            ctx_dev = self.sweep_contexts(times, telescope_location, astroms)
        nt, n = int(ctx_dev.shape[0]), self.Ncomponents
        alt_az = torch.empty((nt, 2, n), dtype=torch.float64, device=self.device)
        lmn = torch.empty((nt, 3, n), dtype=torch.float64, device=self.device) if want_lmn else None
        above = torch.empty((nt, n), dtype=torch.uint8, device=self.device) if want_mask else None
        coords = self._coords_struct()
        events = getattr(self, "_timing_events", None)
        if events is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(torch.cuda.current_stream(self.device))
        self._call("rsk_positions_sweep", ctypes.byref(coords), n, _lib.ptr(ctx_dev), nt, _lib.ptr(alt_az),
                   _lib.ptr(lmn), _lib.ptr(above))
        if events is not None:
            ev1.record(torch.cuda.current_stream(self.device))
            events.append((ev0, ev1))
        return alt_az, lmn, (None if above is None else above.view(torch.bool))

    # ------------------------------------------------------------------------------------------
    # component down-selection: select / cut_nonrising / calculate_rise_set_lsts
    # ------------------------------------------------------------------------------------------
    def _gather(self, t, idx, m):
        """t[..., idx] for a CUDA tensor whose last axis is the component axis (rsk_gather_rows)."""
        t = t.contiguous()
        n_in = t.shape[-1]
        nrows = t.numel() // max(n_in, 1)
        out = torch.empty(t.shape[:-1] + (m,), dtype=t.dtype, device=self.device)
        esize = t.element_size()
        if esize not in (1, 8):
            raise TypeError(f"cannot gather tensors of dtype {t.dtype}")
        self._call("rsk_gather_rows", _lib.ptr(t), nrows, n_in, _lib.ptr(idx), m, esize, _lib.ptr(out))
        return out

    def _take_components(self, idx, m):
        """Keep the components ``idx[:m]`` (CUDA int64), in that order, on every Ncomponents-shaped attribute:
        what ``_select_along_param_axis({"Ncomponents": inds})`` does in the reference (skymodel.py:3258)."""
        idx = idx[:m].contiguous()
        self.stokes = self._gather(self.stokes, idx, m)
        if self.component_type == "point":
            self.lon = self._gather(self.lon, idx, m)
            self.lat = self._gather(self.lat, idx, m)
        elif self.hpx_inds is not None:
            self.hpx_inds = self._gather(self.hpx_inds, idx, m)
        else:  # an implicit pixel range becomes an explicit list
            self.hpx_inds = idx + self.hpx_first
            self.hpx_first = 0
        for attr in ("spectral_index", "reference_frequency", "stokes_error", "beam_amp", "alt_az", "pos_lmn",
                     "_polarized_mask", "_rise_lst", "_set_lst"):
            t = self.__dict__.get(attr)
            if t is not None:
                setattr(self, attr, self._gather(t, idx, m))
        if self.above_horizon is not None:
            self.above_horizon = self._gather(self.above_horizon.view(torch.uint8), idx, m).view(torch.bool)
        host_idx = None
        if self.name is not None or self.extended_model_group is not None or self.extra_columns:
            host_idx = idx.cpu().numpy()
        if self.name is not None:
            self.name = self.name[host_idx]
        if self.extended_model_group is not None:
            self.extended_model_group = self.extended_model_group[host_idx]
        if self.extra_columns:
            self.extra_columns = {k: np.asarray(v)[host_idx] for k, v in self.extra_columns.items()}
        self.Ncomponents = m
        self._n_polarized = self._count(self._polarized_mask, _lib.COUNT_NONZERO_U8)
        self._rot_key = None  # the cached rotation coefficients belong to the old component list
        self.__dict__.pop("_rot", None)
        for cache in ("_subband_cache", "_nan_group_cache", "_nan_cache", "_cm_cache", "_compacted", "_staging", "_graph"):
            self.__dict__.pop(cache, None)
        if self.frame_coherency is not None:
            self.calc_frame_coherency()

    @staticmethod
    def _range_rad(value, what, other):
        """A (lo, hi) range in radians from an astropy Latitude / Longitude or plain radians."""
        if type(value).__name__ == other:
            raise TypeError(f"{what}_range must be an astropy {'Latitude' if what == 'lat' else 'Longitude'} object.")
        return np.atleast_1d(np.asarray(_to_rad(value), dtype=np.float64))

    def _brightness_value(self, value, what):
        if value is None:
            return None
        if hasattr(value, "to_value"):
            try:
                return float(value.to_value(self.stokes_unit))
            except Exception as err:  # astropy UnitConversionError
                raise TypeError(
                    f"{what} must be a Quantity object with units that can be converted to {self.stokes_unit}"
                ) from err
        return float(value)  # plain numbers are taken to be in the unit of ``stokes``

    def select(self, component_inds=None, lat_range=None, lon_range=None, min_brightness=None, max_brightness=None,
               brightness_freq_range=None, non_nan=None, non_negative=False, inplace=True, run_check=True,
               check_extra=True, run_check_acceptability=True):
        """Downselect sources based on various criteria (skymodel.py:3127-3283).

        ``component_inds``: indices to keep; ``lat_range`` / ``lon_range``: (2,) ranges in the catalogue frame
        (astropy Latitude / Longitude or radians; a lon range whose end is below its start wraps through 0);
        ``min_brightness`` / ``max_brightness``: Stokes I cuts (>=, <=; Quantity or a number in the unit of
        ``stokes``) over ``brightness_freq_range`` (Hz, 2 elements; default all stored frequencies); ``non_nan``:
        "any" / "all" / None; ``non_negative``: drop components with any negative Stokes I.  The predicates are
        evaluated in one kernel (rsk_select_mask), the surviving indices come from the horizon-compaction kernels and
        every per-component array is gathered on the GPU (rsk_gather_rows).  The ``run_check*`` arguments are accepted
        and ignored (the UVBase validation framework is out of scope)."""
        skyobj = self if inplace else self.copy()
        if (component_inds is None and lat_range is None and lon_range is None and min_brightness is None
                and max_brightness is None and non_nan is None and non_negative is False):
            return None if inplace else skyobj
        sel = _lib.RskSelect()
        if non_nan is not None:
            allowed_vals = ["any", "all"]
            if non_nan not in allowed_vals:
                raise ValueError(f"If set, non_nan can only be set to one of: {allowed_vals}")
            sel.non_nan = 1 if non_nan == "any" else 2
        sel.non_negative = 1 if non_negative else 0
        if lat_range is not None:
            r = skyobj._range_rad(lat_range, "lat", "Longitude")
            if r.size != 2 or r[0] >= r[1]:
                raise ValueError("lat_range must be 2 element range with the second component larger than the first.")
            sel.use_lat, sel.lat_lo, sel.lat_hi = 1, float(r[0]), float(r[1])
        if lon_range is not None:
            r = skyobj._range_rad(lon_range, "lon", "Latitude")
            if r.size != 2:
                raise ValueError("lon_range must be 2 element range.")
            sel.use_lon, sel.lon_lo, sel.lon_hi = 1, float(r[0]), float(r[1])
        freq_use = None
        if min_brightness is not None or max_brightness is not None:
            # skymodel.py:3047-3124
            if skyobj.spectral_type == "spectral_index":
                raise NotImplementedError("Flux cuts with spectral index type objects is not supported yet.")
            fr = brightness_freq_range
            if fr is not None:
                fr = np.atleast_1d(np.asarray(fr.to_value("Hz") if hasattr(fr, "to_value") else fr, dtype=np.float64))
                if fr.size != 2:
                    raise ValueError("brightness_freq_range must have 2 elements.")
                if skyobj.freq_array is not None:
                    use = (skyobj.freq_array >= fr.min()) & (skyobj.freq_array <= fr.max())
                    if not use.any():
                        raise ValueError("No object frequencies in specified range for flux cuts.")
                    freq_use = torch.as_tensor(use.astype(np.uint8)).to(skyobj.device)
                    sel.freq_use = freq_use.data_ptr()
            mn = skyobj._brightness_value(min_brightness, "min_brightness")
            mx = skyobj._brightness_value(max_brightness, "max_brightness")
            if mn is not None:
                sel.use_min, sel.min_brightness = 1, mn
            if mx is not None:
                sel.use_max, sel.max_brightness = 1, mx
        n = skyobj.Ncomponents
        inds = None
        if component_inds is not None:
            if isinstance(component_inds, torch.Tensor):
                inds = component_inds.to(skyobj.device, dtype=torch.int64).flatten().contiguous()
            else:
                inds = torch.as_tensor(np.atleast_1d(np.asarray(component_inds)).astype(np.int64)).to(skyobj.device)
            if inds.numel() > 0:
                lo, hi = int(inds.min().item()), int(inds.max().item())
                if lo < -n or hi >= n:
                    raise IndexError(f"index {hi if hi >= n else lo} is out of bounds for axis 0 with size {n}")
                if lo < 0:
                    inds = torch.where(inds < 0, inds + n, inds)
            # np.intersect1d (:3226, :3234) and np.union1d (:3041) leave the list sorted and free of repeats
            if sel.non_nan or sel.non_negative or (sel.use_lon and sel.lon_hi < sel.lon_lo):
                inds = torch.unique(inds, sorted=True)
        m = n if inds is None else int(inds.numel())
        predicates = (sel.use_lat or sel.use_lon or sel.use_min or sel.use_max or sel.non_nan or sel.non_negative
                      or sel.cut_nonrising)
        if predicates and m > 0:
            keep = torch.empty(m, dtype=torch.uint8, device=skyobj.device)
            coords = skyobj._coords_struct()
            skyobj._call("rsk_select_mask", ctypes.byref(coords), _lib.ptr(skyobj.stokes), skyobj.Nfreqs, n,
                         _lib.ptr(inds), m, ctypes.byref(sel), _lib.ptr(keep))
            pos, m_out = skyobj._compact_mask(keep)
            idx = pos[:m_out] if inds is None else skyobj._gather(inds, pos[:m_out].contiguous(), m_out)
        else:
            m_out = m
            idx = inds if inds is not None else torch.arange(n, dtype=torch.int64, device=skyobj.device)
        if m_out == 0:
            raise ValueError("Select would result in an empty object.")
        if m_out == n:  # nothing removed (skymodel.py:3250-3254)
            return None if inplace else skyobj
        skyobj.history = (skyobj.history or "") + "  Downselected to specific components using pyradiosky."
        skyobj._take_components(idx, m_out)
        # if all extended models were removed in select, set extended_model_group to None (:3260-3265)
        if skyobj.extended_model_group is not None and not np.any(skyobj.extended_model_group != ""):
            skyobj.extended_model_group = None
        return None if inplace else skyobj

    # ------------------------------------------------------------------------------------------
    # concat (skymodel.py:2714-3010)
    # ------------------------------------------------------------------------------------------
    def _cat(self, a, na, b, nb):
        """np.concatenate((a, b), axis=-1) for CUDA tensors; a missing side is filled with NaN (rsk_concat_rows)."""
        ref = a if a is not None else b
        out = torch.empty(ref.shape[:-1] + (na + nb,), dtype=ref.dtype, device=self.device)
        nrows = ref.numel() // max(ref.shape[-1], 1)
        a = None if a is None else a.contiguous()
        b = None if b is None else b.contiguous()
        self._call("rsk_concat_rows", _lib.ptr(a), na, _lib.ptr(b), nb, nrows, ref.element_size(), _lib.ptr(out))
        return out

    @staticmethod
    def _histories_match(h1, h2):
        # pyuvdata history_utils._check_histories: equal up to white space
        return h1 == h2 or "".join((h1 or "").split()) == "".join((h2 or "").split())

    @staticmethod
    def _history_addition(h1, h2):
        """The words of ``h2`` that do not occur in ``h1``, in order (pyuvdata's _combine_history_addition)."""
        have = " " + (h1 or "") + " "
        words = [w for w in (h2 or "").split(" ") if " " + w + " " not in have]
        return " ".join(words) if words else None

    def concat(self, other, clear_time_position=True, verbose_history=False, inplace=True, run_check=True,
               check_extra=True, run_check_acceptability=True):
        """Combine two SkyModel objects along source axis (skymodel.py:2714-3010).

        Same compatibility rules, warnings, NaN / empty-string fills and history handling as the reference; every
        per-component array is concatenated on the GPU.  Both objects must live on the same device and hold Stokes
        values in the same unit and coordinates in the same frame (the reference converts through astropy there).
        The ``run_check*`` arguments are accepted and ignored."""
        this = self if inplace else self.copy()
        if not isinstance(other, SkyModel):
            raise ValueError("Only SkyModel (or subclass) objects can be added to a SkyModel (or subclass) object")
        if other.device != this.device:
            raise ValueError("Both objects must be on the same device to be combined.")

        def mismatch(name, time_pos=False):
            msg = "UVParameter " + name + " does not match. Cannot combine objects."
            if time_pos:
                msg += (" Set the clear_time_position keyword to True to set this and other time and position "
                        "dependent metadata to None to allow the concatenation to proceed. Time and position dependent "
                        "metadata can be set afterwards using the update_positions method.")
            raise ValueError(msg)

        def same_array(x, y):
            if x is None or y is None:
                return x is None and y is None
            x, y = np.asarray(x), np.asarray(y)
            return x.shape == y.shape and bool(np.allclose(x, y, rtol=0, atol=this.freq_tol))

        if this.component_type != other.component_type:
            mismatch("component_type")
        if this.spectral_type != other.spectral_type:
            mismatch("spectral_type")
        if this.spectral_type in ("subband", "full") and not same_array(this.freq_array, other.freq_array):
            mismatch("freq_array")
        if this.spectral_type == "subband" and not same_array(this.freq_edge_array, other.freq_edge_array):
            mismatch("freq_edge_array")
        if this.component_type == "healpix":
            if this.nside != other.nside:
                mismatch("nside")
            if this.hpx_order != other.hpx_order:
                mismatch("hpx_order")
        if this.stokes_unit != other.stokes_unit:
            mismatch("stokes")
        if this.frame != other.frame or this.equinox != other.equinox:
            mismatch("skycoord" if this.component_type == "point" else "hpx_frame")
        if clear_time_position:
            this.clear_time_position_specific_params()
            other.clear_time_position_specific_params()
        else:
            if this.time != other.time:
                mismatch("time", True)
            if this.telescope_location != other.telescope_location:
                mismatch("telescope_location", True)
            for name in ("alt_az", "pos_lmn", "above_horizon"):
                x, y = getattr(this, name), getattr(other, name)
                if (x is None) != (y is None) or (x is not None and (x.shape != y.shape or not torch.equal(x, y))):
                    mismatch(name, True)
        if this.extra_columns is not None or other.extra_columns is not None:
            if this.extra_columns is None or other.extra_columns is None:
                raise ValueError("One object has extra_columns and the other does not. Cannot combine objects.")
            left, right = tuple(this.extra_columns), tuple(other.extra_columns)
            if set(left) != set(right):
                raise ValueError(
                    "Both objects have extra_columns but the column names do not match. Cannot combine objects. "
                    f"Left object columns are: {left}. Right object columns are: {right}. Unmatched columns are "
                    f"{set(left) - set(right)}")
            for name in left:
                if np.asarray(this.extra_columns[name]).dtype.type != np.asarray(other.extra_columns[name]).dtype.type:
                    raise ValueError(f"Both objects have extra_columns but the dtypes for column {name} do not match. "
                                     "Cannot combine objects.")
            this.extra_columns = {k: np.concatenate((np.asarray(this.extra_columns[k]), np.asarray(other.extra_columns[k])))
                                  for k in left}
        na, nb = this.Ncomponents, other.Ncomponents
        # check for non-overlapping names or healpix inds (:2861-2875)
        if this.component_type == "healpix":
            pa = this.hpx_inds if this.hpx_inds is not None else torch.arange(
                this.hpx_first, this.hpx_first + na, dtype=torch.int64, device=this.device)
            pb = other.hpx_inds if other.hpx_inds is not None else torch.arange(
                other.hpx_first, other.hpx_first + nb, dtype=torch.int64, device=this.device)
            if bool(torch.isin(pa, pb).any()):
                raise ValueError("The two SkyModel objects contain overlapping Healpix pixels.")
            this.hpx_inds = torch.cat((pa, pb))
            this.hpx_first = 0
            if this.name is not None and other.name is not None:
                this.name = np.concatenate((this.name, other.name))
            elif this.name is not None or other.name is not None:
                warnings.warn("Only one object has name values, setting name to None on final object.")
                this.name = None
        else:
            if this.name is not None and other.name is not None:
                if np.intersect1d(this.name, other.name).size > 0:
                    raise ValueError("The two SkyModel objects contain components with the same name.")
                this.name = np.concatenate((this.name, other.name))
            elif this.name is not None or other.name is not None:
                warnings.warn("Only one object has name values, setting name to None on final object.")
                this.name = None
            this.lon = this._cat(this.lon, na, other.lon, nb)
            this.lat = this._cat(this.lat, na, other.lat, nb)
        this.stokes = this._cat(this.stokes, na, other.stokes, nb)
        had_fc = this.frame_coherency is not None, other.frame_coherency is not None
        this.frame_coherency = None
        if had_fc[0] != had_fc[1]:
            warnings.warn("Only one object has frame_coherency values, setting frame_coherency to None on final object. "
                          "Use `calc_frame_coherency` to recalculate them.")
        numeric = ["stokes_error", "beam_amp", "reference_frequency", "spectral_index"]
        for param in numeric:
            x, y = getattr(this, param), getattr(other, param)
            if x is None and y is None:
                continue
            if x is None or y is None:
                warnings.warn(f"Only one object has {param} values. Filling missing values with NaNs.")
            setattr(this, param, this._cat(x, na, y, nb))
        x, y = this.extended_model_group, other.extended_model_group
        if x is not None or y is not None:
            if x is None or y is None:
                warnings.warn("Only one object has extended_model_group values. Filling missing values with empty strings.")
                ref = x if x is not None else y
                x = np.full(na, "", dtype=ref.dtype) if x is None else x
                y = np.full(nb, "", dtype=ref.dtype) if y is None else y
            this.extended_model_group = np.concatenate((x, y))
        this._polarized_mask = this._cat(this._polarized_mask, na, other._polarized_mask, nb)
        this._n_polarized = this._n_polarized + other._n_polarized
        if not clear_time_position:
            for name in ("alt_az", "pos_lmn"):
                if getattr(this, name) is not None:
                    setattr(this, name, this._cat(getattr(this, name), na, getattr(other, name), nb))
            if this.above_horizon is not None:
                this.above_horizon = this._cat(this.above_horizon.view(torch.uint8), na,
                                               other.above_horizon.view(torch.uint8), nb).view(torch.bool)
        this.Ncomponents = na + nb
        this._rot_key = None
        for cache in ("_rot", "_subband_cache", "_nan_group_cache", "_nan_cache", "_cm_cache", "_compacted", "_staging",
                      "_graph", "_rise_lst", "_set_lst"):
            this.__dict__.pop(cache, None)
        if all(had_fc):
            this.calc_frame_coherency()
        # filename (pyuvdata utils._combine_filenames) and history (:3001-3021)
        f1 = this.filename if this.filename is None or isinstance(this.filename, (list, tuple)) else [this.filename]
        f2 = other.filename if other.filename is None or isinstance(other.filename, (list, tuple)) else [other.filename]
        if f1 is not None and f2 is not None:
            this.filename = sorted(set(f1).union(set(f2)))
        elif f2 is not None:
            this.filename = list(f2)
        histories_match = this._histories_match(this.history, other.history)
        this.history = (this.history or "") + " Combined skymodels along the component axis using pyradiosky."
        if not histories_match:
            if verbose_history:
                this.history += " Next object history follows. " + (other.history or "")
            else:
                extra = this._history_addition(this.history, other.history)
                if extra is not None:
                    this.history += " Unique part of next object history follows. " + extra
        return None if inplace else this

    def calculate_rise_set_lsts(self, telescope_latitude, horizon_buffer=0.04364):
        """Calculate the rise & set LSTs given a telescope latitude (skymodel.py:3284-3329).

        Sets ``_rise_lst`` and ``_set_lst`` (radians, CUDA tensors; NaN for sources that never rise or never set).
        ``telescope_latitude``: astropy Latitude or radians; ``horizon_buffer``: angle buffer in radians."""
        lat_rad = float(np.asarray(_to_rad(telescope_latitude), dtype=np.float64))
        n = self.Ncomponents
        self._rise_lst = torch.empty(n, dtype=torch.float64, device=self.device)
        self._set_lst = torch.empty(n, dtype=torch.float64, device=self.device)
        coords = self._coords_struct()
        self._call("rsk_rise_set_lsts", ctypes.byref(coords), n, lat_rad, float(horizon_buffer),
                   _lib.ptr(self._rise_lst), _lib.ptr(self._set_lst))

    def cut_nonrising(self, telescope_latitude, inplace=True, run_check=True, check_extra=True,
                      run_check_acceptability=True):
        """Remove sources that will never rise (skymodel.py:3331-3393): tan(lat_tel) tan(lat) < -1."""
        if type(telescope_latitude).__name__ in ("Longitude", "Angle", "Quantity") or isinstance(telescope_latitude, str):
            raise TypeError("telescope_latitude must be an astropy Latitude object.")
        skyobj = self if inplace else self.copy()
        lat_rad = float(np.asarray(_to_rad(telescope_latitude), dtype=np.float64))
        n = skyobj.Ncomponents
        sel = _lib.RskSelect()
        sel.cut_nonrising, sel.tan_tel_lat = 1, math.tan(lat_rad)
        keep = torch.empty(n, dtype=torch.uint8, device=skyobj.device)
        coords = skyobj._coords_struct()
        skyobj._call("rsk_select_mask", ctypes.byref(coords), _lib.ptr(skyobj.stokes), skyobj.Nfreqs, n, None, n,
                     ctypes.byref(sel), _lib.ptr(keep))
        idx, cnt = skyobj._compact_mask(keep)
        skyobj.select(component_inds=idx[:cnt])
        return None if inplace else skyobj

    def _compact_mask(self, mask_u8):
        """(idx, count) of the non-zero entries of a CUDA uint8 mask, in order (rsk_compact_index)."""
        m = int(mask_u8.numel())
        idx = torch.empty(max(m, 1), dtype=torch.int64, device=self.device)
        cnt = torch.empty(1, dtype=torch.int64, device=self.device)
        nbytes = _lib.load().rsk_workspace_bytes(_lib.OP_COMPACT, m, 0)
        work = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=self.device)
        self._call("rsk_compact_index", _lib.ptr(mask_u8), m, _lib.ptr(idx), _lib.ptr(cnt), _lib.ptr(work),
                   ctypes.c_size_t(nbytes))
        return idx, int(cnt.item())

    # ------------------------------------------------------------------------------------------
    # coherency
    # ------------------------------------------------------------------------------------------
    def calc_frame_coherency(self, store=True):
        """Coherency in the catalogue frame (skymodel.py:2481-2497): (2, 2, Nfreqs, N) complex128."""
        out = torch.empty((2, 2, self.Nfreqs, self.Ncomponents), dtype=torch.complex128, device=self.device)
        self._call("rsk_stokes_to_coherency", _lib.ptr(self.stokes), self.Nfreqs, self.Ncomponents, _lib.ptr(out))
        if store:
            self.frame_coherency = out
        else:
            return out

    def _calc_average_rotation_matrix(self):
        """skymodel.py:2499-2555, from the current astrometry context: (3, 3) numpy."""
        if self._astrom is None:
            raise ValueError("update_positions must be called before the rotation can be computed")
        return astrometry.average_rotation_matrix(self._astrom)

    def _compact_async(self):
        """Queue the horizon compaction; returns (idx, count) CUDA tensors without synchronising."""
        n = self.Ncomponents
        idx = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        cnt = torch.empty(1, dtype=torch.int64, device=self.device)
        nbytes = _lib.load().rsk_workspace_bytes(_lib.OP_COMPACT, n, 0)
        work = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=self.device)
        mask = self.above_horizon.view(torch.uint8)
        self._call("rsk_compact_index", _lib.ptr(mask), n, _lib.ptr(idx), _lib.ptr(cnt), _lib.ptr(work),
                   ctypes.c_size_t(nbytes))
        return idx, cnt

    def compact_above_horizon(self):
        """(idx, n_above): CUDA int64 indices of above-horizon components in catalogue order.

        This is the boolean fancy index of skymodel.py:2685 made explicit; the count comes back to
        the host because the output array of ``coherency_calc`` is sized by it."""
        idx, cnt = self._compact_async()
        return idx, int(cnt.item())

    def coherency_calc(self, store_frame_coherency=True, freqs=None, freq_interp_kind="cubic", out=None,
                       ld_out=None, col0=0, pad=True, sync=True, nan_handling="clip"):
        """Calculate the local coherency in the alt/az basis (skymodel.py:2648-2712).

        :meth:`update_positions` must be run prior to this method.

        Parameters
        ----------
        store_frame_coherency : bool
            Accepted for signature compatibility; the fused kernel never needs the stored frame
            coherency, so nothing is stored and no extra memory is used.
        freqs : array_like of Hz, optional  (extension)
            Evaluate the spectral model at these frequencies inside the same kernel, i.e. the result of
            ``at_frequencies(freqs)`` followed by ``coherency_calc()`` upstream, without materialising
            the (4, Nfreqs, Ncomponents) Stokes cube.  Default: the stored Stokes as they are.
        out, ld_out, col0 : optional  (extension, multi-GPU assembly / buffer reuse)
            Write into columns [col0, col0 + N_above) of a preallocated complex128 buffer whose last
            axis has length ``ld_out``.
        pad : bool  (extension)
            When the output is allocated here, give its rows a stride that is a multiple of 16 elements
            (256 bytes) -- B200 HBM write throughput is ~25 % higher for 256-byte aligned rows -- and
            return the (2, 2, Nfreqs, N_above) view of it.  ``pad=False`` returns a C-contiguous tensor.
        sync : bool  (extension)
            ``sync=False`` queues the whole computation without reading the above-horizon count back:
            returns ``(buffer, count)`` where ``buffer`` has room for every component
            (2, 2, Nfreqs, >= Ncomponents) and ``count`` is a CUDA int64 tensor; columns
            ``[col0, col0 + count)`` are valid.  Nothing in the call then waits for the GPU.

        Returns
        -------
        torch.Tensor, complex128, shape (2, 2, Nfreqs, N_above) -- component axis fastest, only
        components above the horizon, in catalogue order.
        """
        plan = self._plan_coherency(freqs, freq_interp_kind, need_count=sync, nan_handling=nan_handling)
        nf, ncols = plan["nf"], plan["n_above"]
        if out is None:
            if pad:
                # rows on 256-byte boundaries; the stride is also bucketed (~N/64) so that consecutive
                # time steps, whose N_above differ slightly, reuse the same cached allocation
                gran = max(16, -(-(self.Ncomponents // 64) // 16) * 16)
                ld_out = -(-ncols // gran) * gran
            else:
                ld_out = ncols
            try:
                out = torch.empty((2, 2, nf, max(ld_out, 0)), dtype=torch.complex128, device=self.device)
            except torch.OutOfMemoryError:
                # a result that takes more than half of HBM: the cached block of the previous time step (a slightly
                # different size) must go back to the driver before this one fits
                torch.cuda.empty_cache()
                out = torch.empty((2, 2, nf, max(ld_out, 0)), dtype=torch.complex128, device=self.device)
            col0 = 0
        elif ld_out is None:
            ld_out = int(out.shape[-1])
        self._launch_coherency(plan, 0, nf, out, ld_out, col0)
        if not sync:
            return out, plan["count"]
        if out.shape[-1] == ncols and col0 == 0:
            return out
        return out[..., col0:col0 + ncols]

    def apparent_coherency_calc(self, beam="short_dipole", jones=None, jones2=None, freqs=None,
                                freq_interp_kind="cubic", pad=True, nan_handling="clip"):
        """Instrumental coherency ``J1 C J2^H`` of every above-horizon component (extension).

        What a simulator does with the result of :meth:`coherency_calc` -- the contraction
        ``einsum("ab...,bc...,dc...->ad...", J, C, conj(J))`` of tests/test_skymodel.py:1464-1466 -- fused
        into the same kernel, so the Az/El coherency itself is never written to HBM.

        Parameters
        ----------
        beam : {'short_dipole', 'uniform'} -- analytic responses evaluated in the kernel from ``alt_az``
            ('short_dipole': feeds (north, east), as that test arranges pyuvdata's ShortDipoleBeam).
            Ignored when ``jones`` is given.
        jones, jones2 : complex128 CUDA tensors (2, 2, Nfreqs, N_above), optional
            Caller-evaluated Jones matrices [feed][basis] (basis 0 = zenith-angle direction, 1 = azimuth
            direction) for the above-horizon components in catalogue order; ``jones2`` is the second
            antenna of the baseline (default: the same antenna).
        freqs, freq_interp_kind : as :meth:`coherency_calc`.

        Returns
        -------
        torch.Tensor complex128 (2, 2, Nfreqs, N_above).
        """
        plan = self._plan_coherency(freqs, freq_interp_kind, need_count=True, nan_handling=nan_handling)
        nf, ncols = plan["nf"], plan["n_above"]
        b = _lib.RskBeam()
        if jones is not None:
            for name, j in (("jones", jones), ("jones2", jones2)):
                if j is None:
                    continue
                if (not isinstance(j, torch.Tensor) or j.dtype != torch.complex128 or not j.is_cuda
                        or tuple(j.shape) != (2, 2, nf, ncols) or not j.is_contiguous()):
                    raise ValueError(f"{name} must be a contiguous complex128 CUDA tensor of shape (2, 2, {nf}, {ncols})")
            b.type = _lib.BEAM_JONES
            b.jones1 = jones.data_ptr()
            b.jones2 = None if jones2 is None else jones2.data_ptr()
            b.ld_jones = max(ncols, 1)
        else:
            kinds = {"short_dipole": _lib.BEAM_SHORT_DIPOLE, "uniform": _lib.BEAM_UNIFORM}
            if beam not in kinds:
                raise ValueError("beam must be 'short_dipole' or 'uniform' (or pass jones=)")
            b.type = kinds[beam]
            if b.type == _lib.BEAM_SHORT_DIPOLE:
                if self.alt_az is None:
                    raise ValueError("update_positions must be called before apparent_coherency_calc")
                plan["altaz"] = self.alt_az
        plan["beam"] = b
        ld = -(-ncols // 16) * 16 if pad else ncols
        out = torch.empty((2, 2, nf, max(ld, 0)), dtype=torch.complex128, device=self.device)
        self._launch_coherency(plan, 0, nf, out, ld, 0)
        return out if ld == ncols else out[..., :ncols]

    def visibility_calc(self, beam="short_dipole", jones=None, jones2=None, freqs=None, uvw=None,
                        freq_interp_kind="cubic", nan_handling="clip"):
        """Source-summed instrumental coherency (extension; the last step of SURVEY.md 8f N4):

            vis[b, a, d, f] = sum_s (J1 C J2^H)[a, d, f, s] * exp(2j pi (f / c) uvw[b] . pos_lmn[:, s])

        over the above-horizon components, i.e. what a visibility simulator (pyuvsim) computes from
        :meth:`coherency_calc`, the beam Jones matrices and ``pos_lmn``.  Nothing per-component is written to
        HBM.  ``uvw``: (Nbls, 3) baseline vectors in metres (array_like or CUDA tensor); ``None`` gives the
        zero-baseline sum.  ``beam`` / ``jones`` / ``jones2`` / ``freqs`` as :meth:`apparent_coherency_calc`.
        Every spectral type (subband spectra are interpolated to ``freqs`` first, as ``at_frequencies`` does).

        Returns a complex128 CUDA tensor of shape (max(Nbls, 1), 2, 2, Nfreqs).
        """
        if self.spectral_type == "subband" and freqs is not None:
            # the source-sum kernel evaluates flat, power-law and stored spectra; subband spectra (NaN patterns
            # included) are first interpolated to the request (K3, 32 B per evaluation written and read back once)
            fhost = _to_hz_host(freqs)
            cube, _ = self._evaluate_stokes(fhost, freq_interp_kind, nan_handling)
            plan = self._plan_coherency(None, freq_interp_kind, need_count=True, want_cm=False)
            spec = plan["spec"]
            spec.stokes, spec.nf_in = cube.data_ptr(), int(cube.shape[1])
            plan["keep"].append(cube)
            plan["nf"], plan["fdev"] = fhost.size, self._freqs_dev(fhost)
        else:
            # (the source-sum kernel reads the (4, Nf, N) cube itself: no component-major copy is made for it)
            plan = self._plan_coherency(freqs, freq_interp_kind, need_count=True, want_cm=False)
        nf, ncols = plan["nf"], plan["n_above"]
        b = _lib.RskBeam()
        if jones is not None:
            for name, j in (("jones", jones), ("jones2", jones2)):
                if j is None:
                    continue
                if (not isinstance(j, torch.Tensor) or j.dtype != torch.complex128 or not j.is_cuda
                        or tuple(j.shape) != (2, 2, nf, ncols) or not j.is_contiguous()):
                    raise ValueError(f"{name} must be a contiguous complex128 CUDA tensor of shape (2, 2, {nf}, {ncols})")
            b.type = _lib.BEAM_JONES
            b.jones1 = jones.data_ptr()
            b.jones2 = None if jones2 is None else jones2.data_ptr()
            b.ld_jones = max(ncols, 1)
        else:
            kinds = {"short_dipole": _lib.BEAM_SHORT_DIPOLE, "uniform": _lib.BEAM_UNIFORM}
            if beam not in kinds:
                raise ValueError("beam must be 'short_dipole' or 'uniform' (or pass jones=)")
            b.type = kinds[beam]
        altaz = plan["altaz"]
        if b.type == _lib.BEAM_SHORT_DIPOLE:
            if self.alt_az is None:
                raise ValueError("update_positions must be called before visibility_calc")
            altaz = self.alt_az
        nbl = 0
        uvw_dev = None
        if uvw is not None:
            if self.pos_lmn is None:
                raise ValueError("update_positions must be called before visibility_calc with baselines")
            uvw_dev = _as_device_f64(uvw, self.device).reshape(-1, 3).contiguous()
            nbl = int(uvw_dev.shape[0])
        lib = _lib.load()
        wbytes = int(lib.rsk_visibility_workspace_bytes(ncols, nf, nbl))
        work = torch.empty(max(wbytes, 8), dtype=torch.uint8, device=self.device)
        vis = torch.empty((max(nbl, 1), 2, 2, nf), dtype=torch.complex128, device=self.device)
        fptr = None if plan["fdev"] is None else _lib.ptr(plan["fdev"])
        if fptr is None and nbl > 0:  # stored channels: the fringe needs their frequencies
            if self.freq_array is None:
                raise ValueError("freqs must be given to form baseline fringes for a model without a freq_array")
            fptr = _lib.ptr(self._freqs_dev(self.freq_array))
        self._call("rsk_visibility", ctypes.byref(plan["spec"]), ctypes.byref(plan["coords"]), _lib.ptr(altaz), plan["ravg"],
                   None if plan["idx"] is None else _lib.ptr(plan["idx"]), ncols, None, fptr, nf, ctypes.byref(b),
                   _lib.ptr(self.pos_lmn) if nbl > 0 else None, _lib.ptr(uvw_dev), nbl, _lib.ptr(vis), _lib.ptr(work),
                   ctypes.c_size_t(wbytes))
        return vis

    def _plan_coherency(self, freqs, freq_interp_kind, need_count=True, nan_handling="clip", want_cm=True):
        """Everything coherency_calc needs before the launch: horizon compaction, spectral description,
        R_avg.  With ``need_count`` one device->host read (the above-horizon count) is made, after all
        other host-side preparation so the GPU is idle for as short a time as possible."""
        if self.above_horizon is None:
            warnings.warn("Horizon cutoff undefined. Assuming all source components are above the horizon.")
            idx, count = None, None
        else:
            done = self.__dict__.get("_compacted")
            if (done is not None and done[0] is self.above_horizon and done[1] == self.above_horizon._version
                    and self.__dict__.get("_overlap") is not None):
                idx, count = done[2], done[3]  # compacted beside the previous step's kernel (enable_position_overlap)
            else:
                idx, count = self._compact_async()
        if self.telescope_location is not None:
            self._check_tel_location(self.telescope_location)
        keep = []
        fmap = None
        if freqs is None:
            mode = _lib.FULL
            nf = self.Nfreqs
            fdev = None
        else:
            fhost = self._check_request(_to_hz_host(freqs), freq_interp_kind)
            mode = _lib.SPECTRAL_MODES[self.spectral_type]
            fmap = self._full_freq_map(fhost) if mode == _lib.FULL else None
            nf = fhost.size
            fdev = self._freqs_dev(fhost)
        spec = self._spectrum_struct(mode, keep, freq_interp_kind, fmap, want_cm=want_cm)
        if mode == _lib.SPECTRAL_INDEX:
            spec.pivot_hz = float(fhost[fhost.size // 2])  # one pivot for every slab of the request
        # subband spectra with NaNs (skymodel.py:2196-2360): one more launch per NaN pattern, restricted to the
        # pattern's components by a device-side filter -- no host round trip, so it also works with sync=False,
        # coherency_calc_to_host and captured steps
        gspecs = []
        if freqs is not None and mode == _lib.SUBBAND:
            groups = self._nan_groups(fhost, freq_interp_kind, nan_handling)
            gspecs = [self._group_spectrum(g, polarized=True) for g in groups or ()]
        ravg = None
        altaz = None
        if spec.polarized:
            if self.alt_az is None:
                raise ValueError("update_positions must be called before coherency_calc for polarized sources")
            ravg = _lib.host_array(self._calc_average_rotation_matrix().ravel())
            altaz = self.alt_az
        coords = self._coords_struct()
        if count is None:
            n_above, count_dev = self.Ncomponents, None
        elif need_count:
            n_above, count_dev = int(count.item()), None
        else:
            n_above, count_dev = self.Ncomponents, count
        if spec.polarized and self.__dict__.get("_rot_key") == (self.alt_az.data_ptr(), self.alt_az._version):
            coords.rot = self._rot.data_ptr()  # still valid: alt_az untouched since update_positions
        self._last_n_above = n_above
        return dict(spec=spec, keep=keep, mode=mode, nf=nf, fdev=fdev, idx=idx, n_above=n_above, ravg=ravg,
                    altaz=altaz, coords=coords, count=count, count_dev=count_dev, group_specs=gspecs)

    def _launch_coherency(self, plan, f0, f1, out, ld_out, col0, packed=False):
        """Launch K2 for request channels [f0, f1) into ``out`` (complex [2][2][f1-f0][ld_out])."""
        spec, mode = plan["spec"], plan["mode"]
        nf = f1 - f0
        fptr = None
        if plan["fdev"] is not None:
            fptr = ctypes.c_void_p(plan["fdev"].data_ptr() + 8 * f0)
        saved_map = spec.freq_map
        if mode == _lib.FULL and (f0 != 0 or f1 != plan["nf"]):
            if not spec.freq_map:  # identity map needed once a slab does not start at channel 0
                ident = torch.arange(plan["nf"], dtype=torch.int32, device=self.device)
                plan["keep"].append(ident)
                saved_map = ident.data_ptr()
            spec.freq_map = saved_map + 4 * f0
        events = getattr(self, "_timing_events", None)
        if events is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(torch.cuda.current_stream(self.device))
        for sp in (spec, *plan.get("group_specs", ())):
            head = (ctypes.byref(sp), ctypes.byref(plan["coords"]), _lib.ptr(plan["altaz"]), plan["ravg"],
                    None if plan["idx"] is None else _lib.ptr(plan["idx"]), _lib.ptr(plan.get("cols")), plan["n_above"],
                    _lib.ptr(plan["count_dev"]), fptr, nf)
            if plan.get("beam") is not None:
                self._call("rsk_apparent_coherency", *head, ctypes.byref(plan["beam"]), _lib.ptr(out), ld_out, col0)
            else:
                self._call("rsk_coherency_packed" if packed else "rsk_coherency", *head, _lib.ptr(out), ld_out, col0)
        if events is not None:
            ev1.record(torch.cuda.current_stream(self.device))
            events.append((ev0, ev1))
        spec.freq_map = saved_map

    def coherency_calc_to_host(self, freqs, host_out=None, slab=64, freq_interp_kind="cubic", packed=True,
                               nthreads=None, nan_handling="clip"):
        """``coherency_calc(freqs=...)`` delivered into HOST memory (extension).

        The frequency axis is processed in slabs: while the kernel fills one device slab, the previous
        one streams to the host on a second CUDA stream, so HBM never holds more than two slabs and the
        PCIe copy overlaps the compute.

        ``packed=True`` (default): a 2x2 coherency is Hermitian with a real diagonal, so the kernel writes
        4 real numbers per evaluation (``rsk_coherency_packed``), half the bytes cross PCIe into small
        pinned staging buffers, and ``rsk_unpack_coherency_host`` expands each slab with ``nthreads`` host
        threads (pure data movement) into the reference-layout array while the next slab is in flight.
        ``host_out`` may then be any C-contiguous complex128 numpy array / CPU tensor with room for
        4 * Nfreqs * N_above values (allocated if None); the result is complete when the call returns.

        ``packed=False``: the full complex array is copied; ``host_out`` must be a flat *pinned* complex128
        tensor, and the call returns after the last copy has been queued (synchronise before reading).

        ``packed=x`` with 0 < x < 1 (``host_out`` pinned): that share of the slabs goes packed, the rest plain,
        evenly interleaved (for tuning: the best share balances host memory traffic against PCIe time).

        ``packed="auto"`` (``host_out`` pinned): every slab takes whichever of the two routes finishes earlier
        under a two-resource model -- the PCIe link (32 or 64 bytes per evaluation) and the host threads that
        expand packed slabs -- so a slow or busy host is not the bottleneck and a fast one still halves the
        wire bytes.  The unpack cost per evaluation is measured on every call and carried to the next one.

        Returns the (2, 2, Nfreqs, N_above) view of ``host_out``.
        """
        plan = self._plan_coherency(freqs, freq_interp_kind, nan_handling=nan_handling)
        nf, n_above = plan["nf"], plan["n_above"]
        need = 4 * nf * n_above
        if host_out is None:
            host_out = torch.empty(max(need, 1), dtype=torch.complex128, pin_memory=not packed)
        if isinstance(host_out, np.ndarray):
            host_out = torch.from_numpy(host_out.reshape(-1))
        if host_out.dtype != torch.complex128 or host_out.numel() < need or host_out.is_cuda:
            raise ValueError(f"host_out must be a host complex128 buffer with at least {need} elements")
        if packed is False and not host_out.is_pinned():
            raise ValueError("host_out must be pinned when packed=False")
        view = host_out.reshape(-1)[:need].view(2, 2, nf, n_above)
        if need == 0:
            return view
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        copy = self._copy_stream
        slab = max(1, min(slab, nf))
        if packed is False:
            bufs = [torch.empty((2, 2, slab, n_above), dtype=torch.complex128, device=self.device) for _ in range(2)]
            freed = [None, None]
            for i, f0 in enumerate(range(0, nf, slab)):
                f1 = min(nf, f0 + slab)
                buf = bufs[i % 2]
                if freed[i % 2] is not None:
                    main.wait_event(freed[i % 2])
                dst = buf if f1 - f0 == slab else buf.reshape(-1)[: 4 * (f1 - f0) * n_above].view(2, 2, f1 - f0, n_above)
                self._launch_coherency(plan, f0, f1, dst, n_above, 0)
                ready = torch.cuda.Event()
                ready.record(main)
                copy.wait_event(ready)
                with torch.cuda.stream(copy):
                    for a in range(2):
                        for b in range(2):
                            view[a, b, f0:f1].copy_(dst[a, b], non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(copy)
                freed[i % 2] = done
            main.wait_stream(copy)
            for b in bufs:
                b.record_stream(copy)
            return view
        # packed / mixed route: device slab (4, slab, ld) real -> pinned staging -> threaded unpack into host_out;
        # in "auto" mode some slabs instead leave as full complex rows straight into (pinned) host_out
        import os
        import time as _time

        ratio = float(packed) if isinstance(packed, float) and 0.0 < packed < 1.0 else None
        auto = packed == "auto" or ratio is not None
        if auto and not host_out.is_pinned():
            auto, ratio = False, None  # plain slabs need a DMA-able destination
        if nthreads is None:
            nthreads = max(1, min(32, (3 * (os.cpu_count() or 1)) // 4))  # leave cores for the DMA / main thread
        ld = -(-n_above // 32) * 32  # 256-byte aligned rows of doubles
        nbuf = 3  # ring of staging pairs: two slabs stay queued on the GPU while one is unpacked
        key = (slab, ld, auto)
        st = self.__dict__.get("_staging")
        if st is None or st["key"] != key:
            st = dict(key=key,
                      dev=[torch.empty((4, slab, ld), dtype=torch.float64, device=self.device) for _ in range(nbuf)],
                      host=[torch.empty((4, slab, ld), dtype=torch.float64, pin_memory=True) for _ in range(nbuf)],
                      cdev=[torch.empty((2, 2, slab, n_above), dtype=torch.complex128, device=self.device)
                            for _ in range(2)] if auto else None)
            self._staging = st
        lib = _lib.load()
        out_ptr = view.data_ptr()
        pending = []      # (copy-done event, host staging view, f0, nfs), oldest first
        copied = [None] * nbuf
        cfreed = [None, None]
        # two-resource model for "auto": PCIe busy-until (tp) and host-unpack busy-until (th), in seconds from
        # now, with running estimates of the link rate and of the unpack cost per evaluation
        rate = _E2E_STATE["link_bytes_per_s"]
        h_eval = _E2E_STATE["unpack_s_per_eval"]
        tp = th = 0.0
        n_packed = n_plain = 0
        unpack_s = unpack_evals = 0.0
        errors = []

        def drain(item):
            nonlocal unpack_s, unpack_evals
            ev, hbuf, f0, nfs, fin = item
            try:
                ev.synchronize()
                t0 = _time.perf_counter()
                _lib.check("rsk_unpack_coherency_host",
                           lib.rsk_unpack_coherency_host(hbuf.data_ptr(), nfs, n_above, ld, out_ptr, nf, f0, n_above,
                                                         nthreads))
                unpack_s += _time.perf_counter() - t0
                unpack_evals += nfs * n_above
            except Exception as exc:  # surfaced on the calling thread below
                errors.append(exc)
            finally:
                fin.set()

        # the unpack runs on a helper thread (ctypes releases the GIL) so this thread keeps the GPU queue fed
        work = _queue.Queue()

        def worker():
            while True:
                item = work.get()
                if item is None:
                    return
                drain(item)

        helper = _threading.Thread(target=worker, daemon=True)
        helper.start()

        kp = kc = 0
        for f0 in range(0, nf, slab):
            f1 = min(nf, f0 + slab)
            nfs = f1 - f0
            ev_n = nfs * n_above
            use_plain = False
            if auto:
                # Packed halves the wire bytes but costs host time.  Keep sending packed slabs while the host
                # keeps up (at most about one slab of unpacking queued by the time this slab would arrive);
                # when it falls behind, ship this slab in full instead: the link has the spare capacity and
                # the host catches up meanwhile.  In steady state this balances the two resources.
                arrive = tp + 32.0 * ev_n / rate
                backlog = th - arrive
                # ... nor when its unpack would end after everything left could have been shipped in full
                late = max(th, arrive) + ev_n * h_eval > tp + 64.0 * (nf - f0) * n_above / rate
                use_plain = backlog >= ev_n * h_eval or late
                if ratio is not None:  # fixed share of packed slabs, evenly interleaved
                    use_plain = (n_packed + n_plain + 1) * (1.0 - ratio) - n_plain >= 1.0 - 1e-9
            if use_plain:
                k = kc % 2
                kc += 1
                cbuf = st["cdev"][k]
                dst = cbuf if nfs == slab else cbuf.reshape(-1)[: 4 * nfs * n_above].view(2, 2, nfs, n_above)
                if cfreed[k] is not None:
                    main.wait_event(cfreed[k])
                self._launch_coherency(plan, f0, f1, dst, n_above, 0)
                ready = torch.cuda.Event()
                ready.record(main)
                copy.wait_event(ready)
                with torch.cuda.stream(copy):
                    for a in range(2):
                        for b in range(2):
                            view[a, b, f0:f1].copy_(dst[a, b], non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(copy)
                cfreed[k] = done
                tp += 64.0 * ev_n / rate
                n_plain += 1
                continue
            k = kp % nbuf
            kp += 1
            if len(pending) == nbuf:  # host staging buffer k is about to be overwritten: its unpack must be done
                pending.pop(0)[4].wait()
            dbuf, hbuf = st["dev"][k], st["host"][k]
            dview = dbuf if nfs == slab else dbuf.reshape(-1)[: 4 * nfs * ld].view(4, nfs, ld)
            hview = hbuf if nfs == slab else hbuf.reshape(-1)[: 4 * nfs * ld].view(4, nfs, ld)
            if copied[k] is not None:
                main.wait_event(copied[k])  # device slab k must have left for the host before it is refilled
            self._launch_coherency(plan, f0, f1, dview, ld, 0, packed=True)
            ready = torch.cuda.Event()
            ready.record(main)
            copy.wait_event(ready)
            with torch.cuda.stream(copy):
                hview.copy_(dview, non_blocking=True)
                done = torch.cuda.Event()
                done.record(copy)
            copied[k] = done
            item = (done, hview, f0, nfs, _threading.Event())
            pending.append(item)
            work.put(item)
            tp += 32.0 * ev_n / rate
            th = max(th, tp) + ev_n * h_eval
            n_packed += 1
        work.put(None)
        helper.join()
        if errors:
            raise errors[0]
        if auto:
            copy.synchronize()  # plain slabs: complete when the call returns, like the packed ones
            if unpack_evals > 0:  # refine the cost estimate for the next call (the host's speed is not known a priori)
                _E2E_STATE["unpack_s_per_eval"] = 0.5 * h_eval + 0.5 * unpack_s / unpack_evals
            self._e2e_route_counts = (n_packed, n_plain)
        return view

    # ------------------------------------------------------------------------------------------
    # spectra
    # ------------------------------------------------------------------------------------------
    def _check_request(self, fhost, freq_interp_kind):
        """Argument checks of at_frequencies that raise before any GPU work (skymodel.py:2156-2192)."""
        if self.spectral_type == "spectral_index":
            if self._has_nan("spectral_index"):
                raise ValueError("Some spectral index values are NaNs.")
        elif self.spectral_type == "subband":
            if np.max(fhost) > np.max(self.freq_array):
                raise ValueError("A requested frequency is larger than the highest subband frequency.")
            if np.min(fhost) < np.min(self.freq_array):
                raise ValueError("A requested frequency is smaller than the lowest subband frequency.")
            spectral.interp_degree(freq_interp_kind)
        return fhost

    def _nan_groups(self, fhost, freq_interp_kind, nan_handling):
        """Per-NaN-pattern groups of a subband model (skymodel.py:2196-2360), or None without NaNs.

        Components are grouped by the set of stored channels at which all four Stokes parameters are
        finite (``rsk_valid_channels``); every group gets one dense operator built on its valid channels
        (``spectral.build_nan_group_operator``).  Emits the reference's warnings.  Cached until the Stokes
        tensor changes."""
        if self.spectral_type != "subband" or not self._has_nan("stokes"):
            return None
        if nan_handling not in ("propagate", "interp", "clip"):
            raise ValueError("nan_handling must be one of ['propagate', 'interp', 'clip']")
        message = "Some Stokes values are NaNs."
        if nan_handling == "propagate":
            message += " All output Stokes values for sources with any NaN values will be NaN."
        else:
            message += " Interpolating using the non-NaN values only."
        message += " You can change the way NaNs are handled using the `nan_handling` keyword."
        warnings.warn(message)
        key = (self.stokes.data_ptr(), self.stokes._version, self.__dict__.get("_stokes_gen", 0),
               tuple(self.stokes.shape), str(freq_interp_kind), nan_handling, self.freq_array.tobytes())
        cache = self.__dict__.get("_nan_group_cache")
        if cache is None or cache["key"] != key:
            nwords = -(-self.Nfreqs // 64)
            mask = torch.empty((nwords, self.Ncomponents), dtype=torch.int64, device=self.device)
            self._call("rsk_valid_channels", _lib.ptr(self.stokes), self.Nfreqs, self.Ncomponents, _lib.ptr(mask))
            mask_h = mask.cpu().numpy().view(np.uint64)  # (nwords, N): once per catalogue, cached below
            full = np.array([(1 << min(64, self.Nfreqs - 64 * w)) - 1 for w in range(nwords)], dtype=np.uint64)
            bad = np.nonzero(np.any(mask_h != full[:, None], axis=0))[0]
            patterns, inverse = np.unique(mask_h[:, bad].T, axis=0, return_inverse=True)
            inverse = np.asarray(inverse).reshape(-1)
            group_of = np.full(self.Ncomponents, -1, dtype=np.int32)  # component -> its NaN-pattern group
            group_of[bad] = inverse.astype(np.int32)
            group_of_dev = torch.as_tensor(group_of).to(self.device)
            groups = []
            for gi, pat in enumerate(patterns):
                comps = bad[inverse == gi]
                use = np.array([j for j in range(self.Nfreqs) if (int(pat[j // 64]) >> (j % 64)) & 1], dtype=np.int64)
                degree, breaks, ppoly, chan, fallback = spectral.build_nan_group_operator(
                    self.freq_array, use, freq_interp_kind, nan_handling)
                groups.append(dict(
                    index=gi, group_of=group_of_dev,
                    comps_host=comps, use=use, fallback=fallback and use.size > 0, degree=degree,
                    comps=torch.as_tensor(comps.astype(np.int64)).to(self.device),
                    breaks=torch.as_tensor(breaks).to(self.device),
                    ppoly=torch.as_tensor(np.ascontiguousarray(ppoly)).to(self.device),
                    chan=torch.as_tensor(np.ascontiguousarray(chan, dtype=np.int32)).to(self.device)))
            cache = dict(key=key, groups=groups)
            self._nan_group_cache = cache
        groups = cache["groups"]
        if nan_handling != "propagate":  # the per-category warnings of skymodel.py:2315-2359
            n_all = sum(g["comps_host"].size for g in groups if g["use"].size == 0)
            n_high = sum(g["comps_host"].size for g in groups
                         if g["use"].size and np.max(fhost) > np.max(self.freq_array[g["use"]]))
            n_low = sum(g["comps_host"].size for g in groups
                        if g["use"].size and np.min(fhost) < np.min(self.freq_array[g["use"]]))
            # counted only when something is left to interpolate (at_freq_inds_use.size > 0, skymodel.py:2295-2303)
            n_many = sum(g["comps_host"].size for g in groups if g["fallback"] and np.any(
                (fhost >= np.min(self.freq_array[g["use"]])) & (fhost <= np.max(self.freq_array[g["use"]]))))
            if n_all:
                warnings.warn(f"{n_all} components had all NaN Stokes values. "
                              "Output Stokes for these components will all be NaN.")
            tail = ("The Stokes for these components at these frequencies will be NaN."
                    if nan_handling == "interp" else None)
            if n_high:
                warnings.warn(f"{n_high} components had all NaN Stokes values above one or more of the requested "
                              "frequencies. " + (tail or "Using the Stokes value at the highest frequency without a "
                                                 "NaN for these components at these frequencies."))
            if n_low:
                warnings.warn(f"{n_low} components had all NaN Stokes values below one or more of the requested "
                              "frequencies. " + (tail or "Using the Stokes value at the lowest frequency without a "
                                                 "NaN for these components at these frequencies."))
            if n_many:
                warnings.warn(f"{n_many} components had too few non-NaN Stokes values for chosen interpolation. "
                              "Using linear interpolation for these components instead.")
        return groups

    def _group_spectrum(self, group, polarized):
        """RskSpectrum of one NaN-pattern group: dense operator on the group's valid channels."""
        s = _lib.RskSpectrum()
        s.mode = _lib.SUBBAND
        s.stokes = self.stokes.data_ptr()
        s.nf_in = self.Nfreqs
        s.n = self.Ncomponents
        s.degree = group["degree"]
        s.breaks = group["breaks"].data_ptr()
        s.ppoly = group["ppoly"].data_ptr()
        s.nb = group["breaks"].numel() - 1
        s.chan_map = group["chan"].data_ptr()
        s.nf_use = group["chan"].numel()
        if polarized:
            # coherency launches run over the above-horizon columns: the kernel keeps this group's components
            # (at_frequencies instead hands the kernel the group's component list)
            s.group_of = group["group_of"].data_ptr()
            s.group = group["index"]
            if self._n_polarized > 0:
                s.polarized = self._polarized_mask.data_ptr()
        return s

    def _has_nan(self, attr):
        """isnan(...).any() of a device array, cached until the tensor is modified in place or replaced
        (the reference re-scans on every call, skymodel.py:2157, :2196; here that would cost a
        reduction kernel and a host sync per time step)."""
        t = getattr(self, attr)
        key = (t.data_ptr(), t._version, self.__dict__.get("_stokes_gen", 0), tuple(t.shape))
        cache = self.__dict__.setdefault("_nan_cache", {})
        hit = cache.get(attr)
        if hit is None or hit[0] != key:
            hit = (key, self._count(t.contiguous(), _lib.COUNT_NAN_F64) > 0)
            cache[attr] = hit
        return hit[1]

    def _full_freq_map(self, fhost, atol=None):
        """skymodel.py:2165-2181: stored channels matching the request within ``atol`` (1 Hz)."""
        tol = self.freq_tol if atol is None else float(atol.to_value("Hz") if hasattr(atol, "to_value") else atol)
        matches = np.fromiter(
            (np.isclose(freq, fhost, atol=tol).any() for freq in self.freq_array), dtype=bool
        )
        if np.sum(matches) != fhost.size:
            raise ValueError("Some requested frequencies are not present in the current SkyModel.")
        return np.nonzero(matches)[0]

    def _evaluate_stokes(self, fhost, freq_interp_kind="cubic", nan_handling="clip", atol=None):
        """The (4, Nfreqs, Ncomponents) Stokes cube of ``at_frequencies`` (K3 ``rsk_at_frequencies``), NaN-pattern
        groups included.  Returns ``(stokes, freq_map)``; ``freq_map`` is the channel subset of a 'full' model."""
        self._check_request(fhost, freq_interp_kind)
        groups = self._nan_groups(fhost, freq_interp_kind, nan_handling)
        mode = _lib.SPECTRAL_MODES[self.spectral_type]
        fmap = self._full_freq_map(fhost, atol) if mode == _lib.FULL else None
        keep = []
        spec = self._spectrum_struct(mode, keep, freq_interp_kind, fmap, polarized=False)
        nf = fhost.size if fmap is None else fmap.size
        new_stokes = torch.empty((4, nf, self.Ncomponents), dtype=torch.float64, device=self.device)
        fdev = self._freqs_dev(fhost)
        self._call("rsk_at_frequencies", ctypes.byref(spec), _lib.ptr(fdev), nf, None, 0, _lib.ptr(new_stokes))
        for g in groups or ():  # components with NaNs: re-evaluated per pattern on their valid channels
            gs = self._group_spectrum(g, polarized=False)
            self._call("rsk_at_frequencies", ctypes.byref(gs), _lib.ptr(fdev), nf, _lib.ptr(g["comps"]),
                       g["comps"].numel(), _lib.ptr(new_stokes))
        return new_stokes, fmap

    def at_frequencies(self, freqs, inplace=True, freq_interp_kind="cubic", nan_handling="clip",
                       run_check=True, check_extra=True, run_check_acceptability=True, atol=None):
        """Evaluate the stokes array at the specified frequencies (skymodel.py:2084-2389).

        full: subset of existing channels; subband: spline interpolation; spectral_index: power law;
        flat: copy.  The object becomes ``spectral_type='full'``.  Returns the new object iff
        ``inplace`` is False.
        """
        sky = self if inplace else self.copy()
        fhost = _to_hz_host(freqs)
        new_stokes, fmap = self._evaluate_stokes(fhost, freq_interp_kind, nan_handling, atol)
        nf = int(new_stokes.shape[1])
        was_subband = self.spectral_type == "subband"
        sky.stokes = new_stokes
        sky._stokes_changed()
        sky.reference_frequency = None
        sky.Nfreqs = nf
        sky.freq_array = fhost.copy()  # the request itself, as upstream (:2375)
        if fmap is not None and sky.freq_edge_array is not None:
            sky.freq_edge_array = sky.freq_edge_array[:, fmap]  # :2179-2180
        if was_subband and sky.freq_edge_array is not None:
            sky.freq_edge_array = None  # :2376-2377
        sky.spectral_type = "full"
        if sky.frame_coherency is not None:
            sky.calc_frame_coherency()
        if run_check:  # :2382-2385
            sky.check(check_extra=check_extra, run_check_acceptability=run_check_acceptability)
        if not inplace:
            return sky

    # ------------------------------------------------------------------------------------------
    # units
    # ------------------------------------------------------------------------------------------
    def _convert(self, to_jansky):
        if self.spectral_type == "spectral_index" or (
            self.spectral_type == "flat" and self.reference_frequency is not None
        ):
            fchan, fcomp = None, self.reference_frequency
        elif self.freq_array is not None:
            fchan, fcomp = self._freqs_dev(self.freq_array), None
        else:
            raise ValueError(
                "Either reference_frequency or freq_array must be set to convert to "
                + ("Jy." if to_jansky else "K.")
            )
        self._call("rsk_convert_units", _lib.ptr(self.stokes), self.Nfreqs, self.Ncomponents, _lib.ptr(fchan),
                   _lib.ptr(fcomp), 1 if to_jansky else 0)
        self._stokes_changed()
        if self.stokes_error is not None:  # skymodel.py:1583-1584, :1630-1631
            self._call("rsk_convert_units", _lib.ptr(self.stokes_error), self.Nfreqs, self.Ncomponents, _lib.ptr(fchan),
                       _lib.ptr(fcomp), 1 if to_jansky else 0)
        if self.frame_coherency is not None:
            self.calc_frame_coherency()

    def jansky_to_kelvin(self):
        """Jy -> K sr (point) or Jy/sr -> K (healpix) in place (skymodel.py:1593-1635)."""
        if self.stokes_unit == _K_UNITS[self.component_type]:
            return
        self._convert(to_jansky=False)
        self.stokes_unit = _K_UNITS[self.component_type]

    def kelvin_to_jansky(self):
        """K sr -> Jy (point) or K -> Jy/sr (healpix) in place (skymodel.py:1541-1591)."""
        if self.stokes_unit == _JY_UNITS[self.component_type]:
            return
        self._convert(to_jansky=True)
        self.stokes_unit = _JY_UNITS[self.component_type]


class StepGraph:
    """``update_positions`` + horizon compaction + fused ``coherency_calc`` of one :class:`SkyModel` as a CUDA graph.

    Everything that changes from one time step to the next is the 44-double record of ``rsk_pack_step`` (packed
    astrometry context + R_avg); the graph copies it from a pinned host buffer to the device and the position kernel
    reads it there (``rsk_positions_step``), so no launch argument ever changes and the captured graph is replayed
    as is.  The outputs live in fixed buffers: ``sky.alt_az``, ``sky.pos_lmn``, ``sky.above_horizon`` and
    ``self.out`` (2, 2, Nfreqs, ld) with the number of valid columns in ``self.count`` (CUDA int64), exactly what
    ``coherency_calc(out=..., sync=False)`` returns.
    """

    def __init__(self, sky, freqs=None, freq_interp_kind="cubic"):
        self.sky = sky
        self.freqs = None if freqs is None else _to_hz_host(freqs)
        self.kind = freq_interp_kind
        dev, n = sky.device, sky.Ncomponents
        nf = sky.Nfreqs if self.freqs is None else self.freqs.size
        sky.alt_az = torch.empty((2, n), dtype=torch.float64, device=dev)
        sky.pos_lmn = torch.empty((3, n), dtype=torch.float64, device=dev)
        sky.above_horizon = torch.zeros(n, dtype=torch.bool, device=dev)
        if sky._n_polarized > 0:
            sky._rot = torch.empty((3, n), dtype=torch.float64, device=dev)
            sky._rot_key = (sky.alt_az.data_ptr(), sky.alt_az._version)
        self.ld = max(16, -(-n // 16) * 16)
        self.out = torch.empty((2, 2, nf, self.ld), dtype=torch.complex128, device=dev)
        self.step_host = torch.zeros(_lib.STEP_LEN, dtype=torch.float64).pin_memory()
        self.step_dev = torch.zeros(_lib.STEP_LEN, dtype=torch.float64, device=dev)
        self._step_np = self.step_host.numpy()
        self.count = None
        self._graph = None

    def _pack(self, time, telescope_location, astrom):
        sky = self.sky
        if astrom is None:
            if not astrometry.is_time(time):
                raise ValueError(f"time must be an astropy Time object. value was: {time}")
            sky._check_tel_location(telescope_location)
            astrom = astrometry.astrom_context(time, telescope_location)
        astrom = np.ascontiguousarray(astrom, dtype=np.float64).ravel()
        if astrom.size != _lib.ASTROM_LEN:
            raise ValueError(f"astrom must hold {_lib.ASTROM_LEN} doubles (eraASTROM)")
        _lib.check("rsk_pack_step", _lib.load().rsk_pack_step(
            _lib.host_array(astrom), _lib.host_array(sky._frame_to_icrs.ravel()), 1 if sky._n_polarized > 0 else 0,
            self._step_np.ctypes.data_as(_lib.c_double_p)))
        sky.time, sky.telescope_location, sky._astrom = time, telescope_location, astrom

    def _body(self):
        sky = self.sky
        self.step_dev.copy_(self.step_host, non_blocking=True)
        sky._positions_from_step(self.step_dev)
        _, self.count = sky.coherency_calc(freqs=self.freqs, freq_interp_kind=self.kind, out=self.out, ld_out=self.ld,
                                           sync=False)

    def run(self, time, telescope_location, astrom=None):
        """One time step.  Returns ``(out, count)``: columns ``[0, count)`` of ``out`` are valid."""
        self._pack(time, telescope_location, astrom)
        if self._graph is None:
            # eager warm-up on a side stream (fills the caches that would otherwise synchronise during capture)
            side = torch.cuda.Stream(self.sky.device)
            side.wait_stream(torch.cuda.current_stream(self.sky.device))
            with torch.cuda.stream(side):
                self._body()
            torch.cuda.current_stream(self.sky.device).wait_stream(side)
            torch.cuda.synchronize(self.sky.device)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._body()
        self._graph.replay()
        return self.out, self.count
