"""SkyH5 files <-> device-resident :class:`SkyModel`.

Mirrors ``SkyModel.read_skyh5`` (src/pyradiosky/skymodel.py:3491-3766), ``from_skyh5`` (:3768-3774) and
``write_skyh5`` (:4871-5029): same groups and dataset names, same unit / object_type attributes, same
``ValueError`` texts and warnings.  h5py is replaced by librsk's own HDF5 reader/writer
(include/rsk_io.h): header columns are small and come back as numpy arrays, the ``(4, Nfreqs,
Ncomponents)`` Stokes cube streams from the file straight into HBM (``rsk_h5_read_device``: chunks are
gathered/inflated into pinned slabs by host threads while the previous slab is in flight).
"""

import ctypes
import os
import warnings

import numpy as np
import torch

from . import _lib

_VERSION_STR = "  Read/written with pyradiosky_b200 version: 0.1.0."

_LONLAT_NAMES = {"icrs": ("ra", "dec"), "fk5": ("ra", "dec"), "fk4": ("ra", "dec"), "galactic": ("l", "b")}
_ANGLE_TO_RAD = {"rad": 1.0, "deg": np.pi / 180.0, "hourangle": np.pi / 12.0, "arcsec": np.pi / 648000.0,
                 "arcmin": np.pi / 10800.0}
_FREQ_TO_HZ = {"Hz": 1.0, "kHz": 1e3, "MHz": 1e6, "GHz": 1e9}


def _ascii_bytes_to_str(raw):
    """Fixed-length byte strings -> numpy str array.  ASCII (every SkyH5 name in practice) is widened bytewise,
    ~20x faster than numpy's own S -> U cast; anything else is decoded as UTF-8 like the reference
    (skymodel.py:297)."""
    k = raw.dtype.itemsize
    codes = raw.view(np.uint8).reshape(-1, k)
    if codes.size == 0 or codes.max() < 128:
        return np.ascontiguousarray(codes.astype(np.uint32)).view(f"U{k}").reshape(raw.shape)
    return np.char.decode(raw, "utf8")


def _str_to_bytes(arr):
    """numpy str array -> fixed-length UTF-8 byte strings (bytewise narrowing for ASCII)."""
    arr = np.ascontiguousarray(arr)
    k = arr.dtype.itemsize // 4
    if k == 0:
        return arr.astype("S1")
    codes = arr.view(np.uint32).reshape(-1, k)
    if codes.size == 0 or codes.max() < 128:
        return np.ascontiguousarray(codes.astype(np.uint8)).view(f"S{k}").reshape(arr.shape)
    return np.char.encode(arr, "utf8")


class H5Error(OSError):
    """librsk's HDF5 layer returned an error."""


def _h5check(lib, handle, func, code):
    if code < 0 or (code > 0 and func in ("rsk_h5_read_device",)):
        detail = lib.rsk_h5_last_error(handle).decode() if handle else ""
        raise H5Error(f"{func}: {lib.rsk_h5_strerror(int(code)).decode()}" + (f" ({detail})" if detail else ""))
    return code


class H5File:
    """Read-only view of an HDF5 file through ``rsk_h5_*`` (the subset listed in include/rsk_io.h)."""

    def __init__(self, filename):
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        rc = self._lib.rsk_h5_open(os.fspath(filename).encode(), ctypes.byref(self._h))
        if rc != 0:
            self._h = None
            raise H5Error(f"cannot read {filename}: {self._lib.rsk_h5_strerror(rc).decode()}")

    def close(self):
        if self._h:
            self._lib.rsk_h5_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        self.close()

    def kind(self, path):
        return _h5check(self._lib, self._h, "rsk_h5_kind", self._lib.rsk_h5_kind(self._h, path.encode()))

    def __contains__(self, path):
        return self.kind(path) != _lib.H5_ABSENT

    def is_dataset(self, path):
        return self.kind(path) == _lib.H5_DATASET

    def keys(self, group):
        need = _h5check(self._lib, self._h, "rsk_h5_list", self._lib.rsk_h5_list(self._h, group.encode(), None, 0))
        buf = ctypes.create_string_buffer(int(need))
        self._lib.rsk_h5_list(self._h, group.encode(), buf, need)
        text = buf.value.decode("utf8")
        return text.split("\n") if text else []

    def info(self, path):
        info = _lib.RskH5Info()
        _h5check(self._lib, self._h, "rsk_h5_info", self._lib.rsk_h5_info(self._h, path.encode(), ctypes.byref(info)))
        return info

    def shape(self, path):
        info = self.info(path)
        return tuple(info.dims[: info.rank])

    def attr(self, path, name):
        """String attribute, or None if it does not exist."""
        need = self._lib.rsk_h5_attr_string(self._h, path.encode(), name.encode(), None, 0)
        if need == _lib.H5_ERR_NOTFOUND:
            return None
        _h5check(self._lib, self._h, "rsk_h5_attr_string", need)
        buf = ctypes.create_string_buffer(int(need))
        self._lib.rsk_h5_attr_string(self._h, path.encode(), name.encode(), buf, need)
        return buf.value.decode("utf8")

    def read_strings(self, path):
        info = self.info(path)
        if info.type_class == _lib.H5_STRING and info.rank > 0 and info.count > 0:
            # fixed-length records: one raw read, then a bytewise widening to numpy str
            raw = np.empty(int(info.count), dtype=f"S{info.elem_size}")
            _h5check(self._lib, self._h, "rsk_h5_read",
                     self._lib.rsk_h5_read(self._h, path.encode(), raw.ctypes.data, raw.nbytes, 0))
            return _ascii_bytes_to_str(raw).reshape(tuple(info.dims[: info.rank]))
        need = _h5check(self._lib, self._h, "rsk_h5_read_strings",
                        self._lib.rsk_h5_read_strings(self._h, path.encode(), None, 0))
        buf = ctypes.create_string_buffer(max(int(need), 1))
        self._lib.rsk_h5_read_strings(self._h, path.encode(), buf, need)
        vals = [v.decode("utf8") for v in buf.raw[: int(need)].split(b"\0")[: int(info.count)]]
        return vals[0] if info.rank == 0 else np.array(vals).reshape(tuple(info.dims[: info.rank]))

    def read(self, path, nthreads=0):
        """Numeric dataset as a host numpy array (float64 for floats, int64 for integers), strings as str."""
        info = self.info(path)
        if info.kind != _lib.H5_DATASET:
            raise H5Error(f"{path} is not a dataset")
        if info.type_class in (_lib.H5_STRING, _lib.H5_VLEN_STRING):
            return self.read_strings(path)
        is_float = info.type_class == _lib.H5_FLOAT
        out = np.empty(int(info.count), dtype=np.float64 if is_float else np.int64)
        _h5check(self._lib, self._h, "rsk_h5_read_as",
                 self._lib.rsk_h5_read_as(self._h, path.encode(), _lib.H5_AS_F64 if is_float else _lib.H5_AS_I64,
                                          out.ctypes.data, int(info.count), nthreads))
        out = out.reshape(tuple(info.dims[: info.rank]))
        return out[()] if info.rank == 0 else out

    def read_device(self, path, device, dtype=torch.float64, nthreads=0, slab_bytes=0):
        """Numeric dataset as a CUDA tensor, streamed file -> pinned slabs -> HBM."""
        info = self.info(path)
        if dtype not in (torch.float64, torch.int64):
            raise ValueError("dtype must be torch.float64 or torch.int64")
        out = torch.empty(tuple(info.dims[: info.rank]), dtype=dtype, device=device)
        with torch.cuda.device(out.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)
            rc = self._lib.rsk_h5_read_device(self._h, path.encode(),
                                              _lib.H5_AS_F64 if dtype == torch.float64 else _lib.H5_AS_I64,
                                              ctypes.c_void_p(out.data_ptr()), int(info.count), nthreads, slab_bytes,
                                              stream)
        if rc != 0:
            _h5check(self._lib, self._h, "rsk_h5_read_device", rc)
        return out


def _get_value(f, group, name, str_type=None):
    """``_get_value_hdf5_group`` (skymodel.py:270-302) without astropy: returns (value, unit, object_type)."""
    path = f"{group}/{name}"
    value = f.read(path)
    unit = f.attr(path, "unit")
    object_type = f.attr(path, "object_type") or f.attr(path, "angtype")
    return value, unit, object_type


def _angle_rad(value, unit, what):
    if unit is None:
        unit = "rad"
    if unit not in _ANGLE_TO_RAD:
        raise ValueError(f"unsupported angle unit {unit!r} for {what}")
    return np.asarray(value, dtype=np.float64) * _ANGLE_TO_RAD[unit]


def _freq_hz(value, unit):
    if unit is None:
        unit = "Hz"
    if unit not in _FREQ_TO_HZ:
        raise ValueError(f"unsupported frequency unit {unit!r}")
    return np.asarray(value, dtype=np.float64) * _FREQ_TO_HZ[unit]


def _parse_equinox(text):
    """'J2000.000' / 'B1950.000' / '2000.0' -> Julian epoch (the FK5 equinox argument of SkyModel)."""
    t = str(text).strip()
    if t[:1] in ("J", "j"):
        return float(t[1:])
    if t[:1] in ("B", "b"):
        # Besselian -> Julian epoch (Lieske 1979)
        jd = 2415020.31352 + (float(t[1:]) - 1900.0) * 365.242198781
        return 2000.0 + (jd - 2451545.0) / 365.25
    return float(t)


def _frame_from_group(f, group):
    """Frame name and equinox from a nested ``skycoord`` / ``hpx_frame`` group (skymodel.py:3569-3623)."""
    frame = f.read_strings(f"{group}/frame")
    equinox = 2000.0
    if f"{group}/equinox" in f:
        equinox = _parse_equinox(f.read_strings(f"{group}/equinox"))
    return str(frame).lower(), equinox


def get_freq_edges_from_centers(freq_array, atol=1.0):
    """``_get_freq_edges_from_centers`` (skymodel.py:305-329); frequencies in Hz, tolerance 1 Hz (:491)."""
    freq_array = np.asarray(freq_array, dtype=np.float64)
    if freq_array.size == 1:
        raise ValueError(
            "Cannot calculate frequency edges from frequency center array because "
            "there is only one frequency center."
        )
    diff = np.diff(freq_array)
    if not np.isclose(diff.min(), diff.max(), rtol=0, atol=atol):
        raise ValueError(
            "Cannot calculate frequency edges from frequency center array because "
            "frequency center spacing is not constant."
        )
    delta = np.mean(diff)
    return np.stack([freq_array - delta / 2.0, freq_array + delta / 2.0])


def read_skyh5_params(filename, device, skip_params=False, nthreads=0):
    """Everything ``SkyModel.__init__`` needs, read from a SkyH5 file (skymodel.py:3525-3759)."""
    with H5File(filename) as f:
        if f.kind("/Header") != _lib.H5_GROUP:
            raise ValueError("This is not a proper skyh5 file.")
        header = set(f.keys("/Header"))
        init = {"filename": [os.path.basename(os.fspath(filename))]}

        header_params = ["Ncomponents", "Nfreqs", "component_type", "spectral_type", "history", "name", "nside",
                         "hpx_order", "hpx_inds", "freq_array", "freq_edge_array", "reference_frequency",
                         "spectral_index", "extended_model_group"]
        optional = ["hpx_order", "freq_array", "freq_edge_array", "reference_frequency", "spectral_index",
                    "extended_model_group"]
        skippable = ["extended_model_group"]
        string_params = {"component_type", "spectral_type", "history", "name", "hpx_order", "extended_model_group",
                         "frame"}

        component_type = f.read_strings("/Header/component_type")
        equinox = 2000.0
        if component_type != "healpix":
            optional += ["nside", "hpx_inds"]
            skippable += ["nside", "hpx_inds", "hpx_order"]
            if "skycoord" in header:
                frame, equinox = _frame_from_group(f, "/Header/skycoord")
                keys = set(f.keys("/Header/skycoord"))
                lon_name, lat_name = _LONLAT_NAMES.get(frame, ("lon", "lat"))
                if lon_name not in keys or lat_name not in keys:
                    raise ValueError(f"skycoord group lacks {lon_name}/{lat_name} for frame {frame!r}")
                if "representation_type" in keys:
                    rep = f.read_strings("/Header/skycoord/representation_type")
                    if rep not in ("spherical", "unitspherical"):
                        raise ValueError(f"unsupported skycoord representation_type {rep!r}")
                lon, lon_unit, _ = _get_value(f, "/Header/skycoord", lon_name)
                lat, lat_unit, _ = _get_value(f, "/Header/skycoord", lat_name)
                init["lon"] = _angle_rad(lon, lon_unit, lon_name)
                init["lat"] = _angle_rad(lat, lat_unit, lat_name)
                init["frame"] = frame
            else:
                if "lat" in header and "lon" in header and "frame" in header:
                    names = ("lon", "lat")
                elif "ra" in header and "dec" in header:
                    names = ("ra", "dec")
                else:
                    raise ValueError("No component location information found in file.")
                warnings.warn(
                    "Parameter skycoord not found in skyh5 file. "
                    "This skyh5 file was written by an older version of "
                    "pyradiosky. Consider re-writing this file to ensure "
                    "future compatibility"
                )
                lon, lon_unit, _ = _get_value(f, "/Header", names[0])
                lat, lat_unit, _ = _get_value(f, "/Header", names[1])
                init["lon"] = _angle_rad(lon, lon_unit, names[0])
                init["lat"] = _angle_rad(lat, lat_unit, names[1])
        else:
            optional.append("name")
            skippable.append("name")
            if "hpx_frame" in header:
                if f.is_dataset("/Header/hpx_frame"):
                    init["hpx_frame"] = str(f.read_strings("/Header/hpx_frame")).lower()
                else:
                    init["hpx_frame"], equinox = _frame_from_group(f, "/Header/hpx_frame")
            elif "frame" in header:
                init["hpx_frame"] = str(f.read_strings("/Header/frame")).lower()

        if isinstance(skip_params, bool):
            skip_params = skippable if skip_params else []
        if isinstance(skip_params, str):
            skip_params = [skip_params]

        units_of = {}
        for par in header_params:
            if par in optional and (par in skip_params or par not in header):
                continue
            if par not in header:
                raise ValueError(f"Expected parameter {par} is missing in file.")
            path = f"/Header/{par}"
            if par in string_params:
                value = f.read_strings(path)
            elif par == "hpx_inds":
                value = f.read_device(path, device, torch.int64, nthreads)
            elif par in ("reference_frequency", "spectral_index"):
                value = f.read_device(path, device, torch.float64, nthreads)
                units_of[par] = f.attr(path, "unit")
            else:
                value = f.read(path)
                units_of[par] = f.attr(path, "unit")
            if par == "nside":
                value = int(value)
            init[par] = value

        # checks on the parameters that are not passed to __init__ (skymodel.py:3673-3685)
        ncomp, nfreqs = int(init["Ncomponents"]), int(init["Nfreqs"])
        if component_type == "healpix":
            if ncomp != init["hpx_inds"].numel():
                raise ValueError("Ncomponents is not equal to the size of 'hpx_inds'.")
        else:
            if ncomp != np.size(init["name"]):
                raise ValueError("Ncomponents is not equal to the size of 'name'.")

        if "freq_array" in init:
            init["freq_array"] = _freq_hz(init["freq_array"], units_of.get("freq_array"))
            if nfreqs != init["freq_array"].size:
                raise ValueError("Nfreqs is not equal to the size of 'freq_array'.")
            if "freq_edge_array" in init:
                init["freq_edge_array"] = _freq_hz(init["freq_edge_array"], units_of.get("freq_edge_array"))
            elif init["spectral_type"] == "subband":
                try:
                    init["freq_edge_array"] = get_freq_edges_from_centers(init["freq_array"])
                except ValueError:
                    warnings.warn(
                        "No freq_edge_array in this file and frequencies are "
                        "not evenly spaced, so spectral_type will be set to "
                        "'full' rather than 'subband'."
                    )
                    init["spectral_type"] = "full"
        if "reference_frequency" in init:
            scale = _FREQ_TO_HZ.get(units_of.get("reference_frequency") or "Hz")
            if scale is None:
                raise ValueError("unsupported reference_frequency unit")
            if scale != 1.0:
                init["reference_frequency"] = init["reference_frequency"] * scale
        init.pop("Ncomponents")
        init.pop("Nfreqs")

        if "extra_columns" in header:
            init["extra_columns"] = {k: f.read(f"/Header/extra_columns/{k}") for k in f.keys("/Header/extra_columns")}

        # the Stokes cube goes from the file straight to the device
        if "/Data/stokes" not in f:
            raise ValueError("Expected parameter stokes is missing in file.")
        init["stokes"] = f.read_device("/Data/stokes", device, torch.float64, nthreads)
        init["stokes_unit"] = f.attr("/Data/stokes", "unit")
        for extra in ("stokes_error", "beam_amp"):
            for grp in ("/Data", "/Header"):  # "/Header" is the old location (skymodel.py:3724-3736)
                if f"{grp}/{extra}" in f:
                    init[extra] = f.read_device(f"{grp}/{extra}", device, torch.float64, nthreads)
                    break

        if component_type != "healpix" and "frame" not in init:
            if "frame" in header:
                init["frame"] = str(f.read_strings("/Header/frame")).lower()
            else:
                warnings.warn(
                    "No frame available in this file, assuming 'icrs'. "
                    "Consider re-writing this file to ensure future compatibility."
                )
                init["frame"] = "icrs"

    if component_type == "healpix":
        if "hpx_frame" not in init:
            warnings.warn(
                "No frame available in this file, assuming 'icrs'. "
                "Consider re-writing this file to ensure future compatibility."
            )
            init["hpx_frame"] = "icrs"
    init["equinox"] = equinox
    return init


# ---------------------------------------------------------------------------------------------------
# writer
# ---------------------------------------------------------------------------------------------------
class H5Writer:
    """``rsk_h5w_*`` wrapper: groups, numpy datasets, string attributes."""

    def __init__(self, filename):
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        rc = self._lib.rsk_h5w_create(os.fspath(filename).encode(), ctypes.byref(self._h))
        if rc != 0:
            self._h = None
            raise H5Error(f"cannot create {filename}: {self._lib.rsk_h5_strerror(rc).decode()}")

    def _check(self, func, rc):
        if rc != 0:
            raise H5Error(f"{func}: {self._lib.rsk_h5_strerror(rc).decode()}")

    def group(self, path):
        self._check("rsk_h5w_group", self._lib.rsk_h5w_group(self._h, path.encode()))

    def dataset(self, path, value, chunks=None, gzip=0, unit=None, object_type=None):
        if isinstance(value, str):
            value = np.bytes_(value.encode("utf8"))
        arr = np.asarray(value)
        if arr.dtype.kind == "U":
            arr = _str_to_bytes(arr) if arr.ndim > 0 else np.char.encode(arr, "utf8")
        if arr.dtype.kind == "b":
            arr = arr.astype(np.int8)
        if arr.ndim > 0:
            arr = np.ascontiguousarray(arr)
        kind = {"i": _lib.H5_INT, "u": _lib.H5_UINT, "f": _lib.H5_FLOAT, "S": _lib.H5_STRING}.get(arr.dtype.kind)
        if kind is None:
            raise TypeError(f"cannot store dtype {arr.dtype} in a SkyH5 file")
        if kind == _lib.H5_STRING and arr.dtype.itemsize == 0:
            arr = arr.astype("S1")
        dims = (ctypes.c_int64 * max(arr.ndim, 1))(*arr.shape)
        ch = None
        if chunks is not None and arr.ndim > 0 and arr.size > 0:
            ch = (ctypes.c_int64 * arr.ndim)(*[int(c) for c in chunks])
        self._check("rsk_h5w_dataset",
                    self._lib.rsk_h5w_dataset(self._h, path.encode(), kind, arr.dtype.itemsize, arr.ndim, dims,
                                              ctypes.c_void_p(arr.ctypes.data), ch, int(gzip)))
        if unit is not None:
            self.attr(path, "unit", unit)
        if object_type is not None:
            self.attr(path, "object_type", object_type)

    def dataset_device(self, path, tensor, chunks=None, gzip=0, unit=None):
        """float64 / int64 CUDA tensor (up to 3-D) -> dataset, streamed HBM -> pinned slabs -> file."""
        t = tensor.contiguous()
        if t.dtype not in (torch.float64, torch.int64) or not t.is_cuda or t.dim() > 3:
            raise TypeError("dataset_device takes float64 or int64 CUDA tensors of up to three dimensions")
        kind = _lib.H5_FLOAT if t.dtype == torch.float64 else _lib.H5_INT
        dims = (ctypes.c_int64 * max(t.dim(), 1))(*t.shape)
        ch = None
        if chunks is not None and t.dim() > 0 and t.numel() > 0:
            ch = (ctypes.c_int64 * t.dim())(*[int(c) for c in chunks])
        with torch.cuda.device(t.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
            self._check("rsk_h5w_dataset_device",
                        self._lib.rsk_h5w_dataset_device(self._h, path.encode(), kind, 8, t.dim(), dims,
                                                         ctypes.c_void_p(t.data_ptr()), ch, int(gzip), stream))
        if unit is not None:
            self.attr(path, "unit", unit)

    def attr(self, path, name, value):
        self._check("rsk_h5w_attr_string",
                    self._lib.rsk_h5w_attr_string(self._h, path.encode(), name.encode(), str(value).encode("utf8")))

    def close(self):
        if self._h:
            h, self._h = self._h, None
            self._check("rsk_h5w_close", self._lib.rsk_h5w_close(h))

    def __enter__(self):
        return self

    def __exit__(self, exc_type, *exc):
        if exc_type is None:
            self.close()
        elif self._h:
            h, self._h = self._h, None
            self._lib.rsk_h5w_close(h)


def guess_chunks(shape, itemsize=8, target_bytes=1 << 20):
    """Chunk shape for ``chunks=True`` (skymodel.py:5004): whole runs of the fastest axis, about 1 MiB."""
    shape = [max(int(s), 1) for s in shape]
    chunk = [1] * len(shape)
    budget = max(target_bytes // itemsize, 1)
    for ax in range(len(shape) - 1, -1, -1):
        chunk[ax] = int(min(shape[ax], max(budget, 1)))
        budget //= chunk[ax]
        if budget < 1:
            break
    return tuple(chunk)


def _host(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def write_skyh5(sky, filename, clobber=False, data_compression=None):
    """``SkyModel.write_skyh5`` (skymodel.py:4871-5029) for the device-resident model."""
    if data_compression not in (None, "gzip"):
        raise NotImplementedError("data_compression must be None or 'gzip'")
    history = getattr(sky, "history", None)
    if history is None:
        history = _VERSION_STR
    elif _VERSION_STR.strip() not in history:
        history = history + _VERSION_STR
    sky.history = history
    if os.path.exists(filename):
        if not clobber:
            raise OSError("File exists; If overwriting is desired set the clobber keyword to True.")
        print("File exists; clobbering.")
    gz = 4 if data_compression == "gzip" else 0
    with H5Writer(filename) as w:
        w.group("/Header")
        w.dataset("/Header/Ncomponents", np.int64(sky.Ncomponents))
        w.dataset("/Header/Nfreqs", np.int64(sky.Nfreqs))
        w.dataset("/Header/component_type", sky.component_type)
        w.dataset("/Header/spectral_type", sky.spectral_type)
        w.dataset("/Header/history", history)
        if sky.name is not None:
            w.dataset("/Header/name", np.asarray(sky.name).astype(str))
        # every header parameter that is set is written (skymodel.py:4925-4956): a point object made by
        # healpix_to_point keeps nside / hpx_order / hpx_inds
        if sky.nside is not None:
            w.dataset("/Header/nside", np.int64(sky.nside))
        if sky.hpx_order is not None:
            w.dataset("/Header/hpx_order", sky.hpx_order)
        inds = sky.hpx_inds
        if inds is None and sky.component_type == "healpix":  # implicit pixel range
            inds = np.arange(sky.hpx_first, sky.hpx_first + sky.Ncomponents, dtype=np.int64)
        if inds is not None:
            w.dataset("/Header/hpx_inds", _host(inds).astype(np.int64))
        if sky.freq_array is not None:
            w.dataset("/Header/freq_array", np.asarray(sky.freq_array, dtype=np.float64), unit="Hz")
        if getattr(sky, "freq_edge_array", None) is not None:
            w.dataset("/Header/freq_edge_array", np.asarray(sky.freq_edge_array, dtype=np.float64), unit="Hz")
        if sky.reference_frequency is not None:
            w.dataset("/Header/reference_frequency", _host(sky.reference_frequency), unit="Hz")
        if sky.spectral_index is not None:
            w.dataset("/Header/spectral_index", _host(sky.spectral_index))
        if getattr(sky, "extended_model_group", None) is not None:
            w.dataset("/Header/extended_model_group", np.asarray(sky.extended_model_group).astype(str))
        # nested frame description, as skycoord.info._represent_as_dict() lays it out (skymodel.py:4958-4987)
        if sky.component_type == "point":
            lon_name, lat_name = _LONLAT_NAMES.get(sky.frame, ("lon", "lat"))
            w.group("/Header/skycoord")
            w.dataset(f"/Header/skycoord/{lon_name}", _host(sky.lon), unit="rad", object_type="longitude")
            w.dataset(f"/Header/skycoord/{lat_name}", _host(sky.lat), unit="rad", object_type="latitude")
            w.dataset("/Header/skycoord/representation_type", "spherical")
            w.dataset("/Header/skycoord/frame", sky.frame)
            if sky.frame == "fk5":
                w.dataset("/Header/skycoord/equinox", f"J{sky.equinox:.3f}", object_type="time")
        else:
            w.group("/Header/hpx_frame")
            w.dataset("/Header/hpx_frame/frame", sky.frame)
            if sky.frame == "fk5":
                w.dataset("/Header/hpx_frame/equinox", f"J{sky.equinox:.3f}", object_type="time")
        if getattr(sky, "extra_columns", None):
            w.group("/Header/extra_columns")
            for key, value in sky.extra_columns.items():
                w.dataset(f"/Header/extra_columns/{key}", value)
        w.group("/Data")
        # the cubes go HBM -> pinned slabs -> file, without a host copy of the whole array
        w.dataset_device("/Data/stokes", sky.stokes, chunks=guess_chunks(sky.stokes.shape), gzip=gz, unit=sky.stokes_unit)
        for extra in ("stokes_error", "beam_amp"):
            val = getattr(sky, extra, None)
            if val is not None:
                w.dataset_device(f"/Data/{extra}", val, chunks=guess_chunks(val.shape), gzip=gz,
                                 unit=sky.stokes_unit if extra == "stokes_error" else None)
