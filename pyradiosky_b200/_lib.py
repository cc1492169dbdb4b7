"""ctypes binding of librsk.so -- the thin layer between host Python and the sm_100a kernels.

This file is exactly the stub a pyradiosky maintainer would add to call the C ABI declared in
``include/rsk.h``.  There is no fallback: if the shared library has not been
built (``python -c "import __graft_entry__ as g; g.build()"``) importing it raises.
"""

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "librsk.so")

ASTROM_LEN = 31
CTX_LEN = 34
STEP_LEN = 44

FLAT, SPECTRAL_INDEX, SUBBAND, FULL = 0, 1, 2, 3
SPECTRAL_MODES = {"flat": FLAT, "spectral_index": SPECTRAL_INDEX, "subband": SUBBAND, "full": FULL}
FRAME_ICRS, FRAME_FK5, FRAME_GALACTIC = 0, 1, 2
OP_COMPACT = 1

c_double_p = ctypes.POINTER(ctypes.c_double)


class RskCoords(ctypes.Structure):
    _fields_ = [
        ("lon", ctypes.c_void_p),
        ("lat", ctypes.c_void_p),
        ("nside", ctypes.c_int64),
        ("nested", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("hpx_inds", ctypes.c_void_p),
        ("hpx_first", ctypes.c_int64),
        ("rot", ctypes.c_void_p),
    ]


class RskSpectrum(ctypes.Structure):
    _fields_ = [
        ("mode", ctypes.c_int32),
        ("degree", ctypes.c_int32),
        ("stokes", ctypes.c_void_p),
        ("nf_in", ctypes.c_int64),
        ("n", ctypes.c_int64),
        ("spectral_index", ctypes.c_void_p),
        ("ref_freq", ctypes.c_void_p),
        ("breaks", ctypes.c_void_p),
        ("ppoly", ctypes.c_void_p),
        ("nb", ctypes.c_int64),
        ("freq_map", ctypes.c_void_p),
        ("polarized", ctypes.c_void_p),
        ("bcoef", ctypes.c_void_p),
        ("nc", ctypes.c_int64),
        ("first", ctypes.c_void_p),
        ("pband", ctypes.c_void_p),
        ("chan_map", ctypes.c_void_p),
        ("nf_use", ctypes.c_int64),
        ("pivot_hz", ctypes.c_double),
        ("group_of", ctypes.c_void_p),
        ("group", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("stokes_cm", ctypes.c_void_p),
    ]


class RskBeam(ctypes.Structure):
    _fields_ = [
        ("type", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("jones1", ctypes.c_void_p),
        ("jones2", ctypes.c_void_p),
        ("ld_jones", ctypes.c_int64),
    ]


BEAM_JONES, BEAM_SHORT_DIPOLE, BEAM_UNIFORM = 0, 1, 2
COUNT_NAN_F64, COUNT_NONZERO_U8, COUNT_NEGATIVE_F64 = 0, 1, 2


class RskSelect(ctypes.Structure):
    _fields_ = [
        ("use_lat", ctypes.c_int32),
        ("use_lon", ctypes.c_int32),
        ("use_min", ctypes.c_int32),
        ("use_max", ctypes.c_int32),
        ("non_nan", ctypes.c_int32),
        ("non_negative", ctypes.c_int32),
        ("cut_nonrising", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("lat_lo", ctypes.c_double),
        ("lat_hi", ctypes.c_double),
        ("lon_lo", ctypes.c_double),
        ("lon_hi", ctypes.c_double),
        ("min_brightness", ctypes.c_double),
        ("max_brightness", ctypes.c_double),
        ("tan_tel_lat", ctypes.c_double),
        ("freq_use", ctypes.c_void_p),
    ]


_SIGNATURES = {
    "rsk_version": (ctypes.c_int, []),
    "rsk_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "rsk_astrom_context": (ctypes.c_int, [ctypes.c_double] * 8 + [c_double_p]),
    "rsk_astrom_contexts": (ctypes.c_int, [c_double_p, c_double_p, ctypes.c_int64] + [ctypes.c_double] * 6 + [c_double_p]),
    "rsk_geodetic_to_geocentric": (ctypes.c_int, [ctypes.c_double] * 3 + [c_double_p]),
    "rsk_frame_to_icrs": (ctypes.c_int, [ctypes.c_int, ctypes.c_double, c_double_p]),
    "rsk_ravg": (ctypes.c_int, [c_double_p, c_double_p]),
    "rsk_pack_context": (ctypes.c_int, [c_double_p, c_double_p, c_double_p]),
    "rsk_pack_contexts": (ctypes.c_int, [c_double_p, ctypes.c_int64, c_double_p, c_double_p]),
    "rsk_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int64, ctypes.c_int64]),
    "rsk_positions": (
        ctypes.c_int,
        [ctypes.POINTER(RskCoords), ctypes.c_int64, c_double_p, c_double_p, c_double_p, ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "rsk_positions_step": (
        ctypes.c_int,
        [ctypes.POINTER(RskCoords), ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_void_p],
    ),
    "rsk_pack_step": (ctypes.c_int, [c_double_p, c_double_p, ctypes.c_int, c_double_p]),
    "rsk_positions_sweep": (
        ctypes.c_int,
        [ctypes.POINTER(RskCoords), ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "rsk_healpix_lonlat": (
        ctypes.c_int,
        [ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "rsk_lonlat_to_healpix": (
        ctypes.c_int,
        [ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
         ctypes.c_void_p],
    ),
    "rsk_scale": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_int, ctypes.c_void_p]),
    "rsk_segment_sum": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
         ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "rsk_select_mask": (
        ctypes.c_int,
        [ctypes.POINTER(RskCoords), ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
         ctypes.POINTER(RskSelect), ctypes.c_void_p, ctypes.c_void_p],
    ),
    "rsk_gather_rows": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
         ctypes.c_void_p, ctypes.c_void_p],
    ),
    "rsk_concat_rows": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
         ctypes.c_void_p, ctypes.c_void_p],
    ),
    "rsk_rotate_lonlat": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, c_double_p, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_void_p],
    ),
    "rsk_coherency_to_stokes": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "rsk_healpix_interp": (
        ctypes.c_int,
        [ctypes.c_int64, ctypes.c_int, c_double_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
         ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "rsk_rise_set_lsts": (
        ctypes.c_int,
        [ctypes.POINTER(RskCoords), ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_void_p],
    ),
    "rsk_compact_index": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_size_t, ctypes.c_void_p],
    ),
    "rsk_count": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "rsk_polarized_mask": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "rsk_coherency": (
        ctypes.c_int,
        [ctypes.POINTER(RskSpectrum), ctypes.POINTER(RskCoords), ctypes.c_void_p, c_double_p,
         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
         ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p],
    ),
    "rsk_coherency_packed": (
        ctypes.c_int,
        [ctypes.POINTER(RskSpectrum), ctypes.POINTER(RskCoords), ctypes.c_void_p, c_double_p,
         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
         ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p],
    ),
    "rsk_apparent_coherency": (
        ctypes.c_int,
        [ctypes.POINTER(RskSpectrum), ctypes.POINTER(RskCoords), ctypes.c_void_p, c_double_p,
         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
         ctypes.POINTER(RskBeam), ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p],
    ),
    "rsk_visibility": (
        ctypes.c_int,
        [ctypes.POINTER(RskSpectrum), ctypes.POINTER(RskCoords), ctypes.c_void_p, c_double_p, ctypes.c_void_p,
         ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(RskBeam), ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p],
    ),
    "rsk_visibility_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int64, ctypes.c_int64, ctypes.c_int64]),
    "rsk_unpack_coherency_host": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
         ctypes.c_int64, ctypes.c_int64, ctypes.c_int],
    ),
    "rsk_spline_coefficients": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
         ctypes.c_void_p, ctypes.c_void_p],
    ),
    "rsk_at_frequencies": (
        ctypes.c_int,
        [ctypes.POINTER(RskSpectrum), ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
         ctypes.c_void_p, ctypes.c_void_p],
    ),
    "rsk_valid_channels": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "rsk_copy_2d": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t,
         ctypes.c_void_p],
    ),
    "rsk_stokes_component_major": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "rsk_stokes_to_coherency": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "rsk_convert_units": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_int, ctypes.c_void_p],
    ),
}


class RskH5Info(ctypes.Structure):
    """``RskH5Info`` of include/rsk_io.h."""

    _fields_ = [
        ("kind", ctypes.c_int32),
        ("type_class", ctypes.c_int32),
        ("elem_size", ctypes.c_int32),
        ("rank", ctypes.c_int32),
        ("dims", ctypes.c_int64 * 8),
        ("count", ctypes.c_int64),
        ("layout", ctypes.c_int32),
        ("nfilters", ctypes.c_int32),
        ("chunk", ctypes.c_int64 * 8),
        ("nchunks", ctypes.c_int64),
    ]


H5_ABSENT, H5_GROUP, H5_DATASET = 0, 1, 2
H5_INT, H5_UINT, H5_FLOAT, H5_STRING, H5_VLEN_STRING = 0, 1, 2, 3, 4
H5_AS_F64, H5_AS_I64 = 0, 1
H5_ERR_NOTFOUND = -105

_c_int64_p = ctypes.POINTER(ctypes.c_int64)

# include/rsk_io.h: the SkyH5 (HDF5 subset) reader / writer and the file -> HBM streaming path
_IO_SIGNATURES = {
    "rsk_h5_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    "rsk_h5_close": (None, [ctypes.c_void_p]),
    "rsk_h5_kind": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p]),
    "rsk_h5_list": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int64]),
    "rsk_h5_info": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(RskH5Info)]),
    "rsk_h5_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]),
    "rsk_h5_read_as": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int],
    ),
    "rsk_h5_read_strings": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int64]),
    "rsk_h5_attr_string": (
        ctypes.c_int64,
        [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int64],
    ),
    "rsk_h5_attr_f64": (
        ctypes.c_int64,
        [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, c_double_p, ctypes.c_int64],
    ),
    "rsk_h5_read_device": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
         ctypes.c_int64, ctypes.c_void_p],
    ),
    "rsk_h5_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "rsk_h5_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "rsk_h5w_create": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    "rsk_h5w_group": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p]),
    "rsk_h5w_dataset": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_int64_p, ctypes.c_void_p,
         _c_int64_p, ctypes.c_int],
    ),
    "rsk_h5w_dataset_device": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_int64_p, ctypes.c_void_p,
         _c_int64_p, ctypes.c_int, ctypes.c_void_p],
    ),
    "rsk_h5w_attr_string": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]),
    "rsk_h5w_close": (ctypes.c_int, [ctypes.c_void_p]),
}
_SIGNATURES.update(_IO_SIGNATURES)

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

# Kernel launches per entry point (for callers that report how many of the library's kernels they queued, e.g.
# bench.py's ``gpu_launches``): every call that is not listed launches exactly one kernel, host helpers none.
LAUNCHES_PER_CALL = {"rsk_compact_index": 3, "rsk_visibility": 2, "rsk_count": 1}
LAUNCHES = [0]  # process-wide running total, advanced by SkyModel._call


def count_launches(fname):
    LAUNCHES[0] += LAUNCHES_PER_CALL.get(fname, 1)

_lib = None


class RskError(RuntimeError):
    """A librsk call returned a non-zero status."""

    def __init__(self, func, code, message):
        super().__init__(f"{func} failed with status {code}: {message}")
        self.func = func
        self.code = code


def load():
    """Load librsk.so (once) and declare every signature.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` from the repository root. "
            "pyradiosky_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(func, code):
    """Raise RskError for a non-zero status."""
    if code != 0:
        msg = load().rsk_strerror(code).decode()
        raise RskError(func, code, msg)


def host_array(values, n=None):
    """Pointer to a host double array for a ctypes call (or zeros of length n).

    numpy input is passed without a Python-level copy loop: the returned pointer keeps the (contiguous float64)
    array alive."""
    if values is None:
        return (ctypes.c_double * n)()
    import numpy as np

    arr = np.ascontiguousarray(values, dtype=np.float64)
    return arr.ctypes.data_as(c_double_p)


def ptr(tensor):
    """Device (or host) address of a torch tensor, or None."""
    if tensor is None:
        return None
    return ctypes.c_void_p(tensor.data_ptr())
