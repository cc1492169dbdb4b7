"""Module-level helpers of the hot path, mirroring ``pyradiosky.utils``.

* :func:`stokes_to_coherency` -- src/pyradiosky/utils.py:17-61
* :func:`jy_to_ksr`           -- src/pyradiosky/utils.py:105-126
"""

import numpy as np
import torch

from . import _lib

C_MPS = 299792458.0
K_B = 1.380649e-23


def _unit_of(arr, default=None):
    unit = getattr(arr, "unit", None)
    return default if unit is None else str(unit)


def _as_device_f64(arr, device=None):
    """torch float64 CUDA tensor from numpy / torch / astropy-Quantity-like input."""
    if hasattr(arr, "unit") and hasattr(arr, "value"):
        arr = arr.value
    if isinstance(arr, torch.Tensor):
        t = arr.to(dtype=torch.float64)
        if not t.is_cuda:
            t = t.to(device if device is not None else "cuda")
        return t.contiguous()
    return torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float64)).to(
        device if device is not None else "cuda"
    )


def stokes_to_coherency(stokes_arr):
    """Convert a Stokes array to coherency matrices on the GPU.

    Parameters
    ----------
    stokes_arr : array_like, shape (4,) or (4, Nfreqs, Ncomponents)
        [I, Q, U, V].  numpy arrays, torch tensors and Quantity-like objects are accepted (the
        reference insists on an astropy Quantity, utils.py:32-33; astropy is optional here).

    Returns
    -------
    complex128 array of shape (2, 2) or (2, 2, Nfreqs, Ncomponents); a torch CUDA tensor if the input
    was a torch tensor, otherwise a numpy array.
    """
    shape = tuple(stokes_arr.shape)
    if shape[0] != 4:
        raise ValueError("First dimension of stokes_vector must be length 4.")
    single = len(shape) == 1
    want_torch = isinstance(stokes_arr, torch.Tensor)
    st = _as_device_f64(stokes_arr)
    if single:
        st = st.reshape(4, 1, 1)
    if st.dim() != 3:
        raise ValueError("stokes_arr must have shape (4,) or (4, Nfreqs, Ncomponents).")
    nf, n = int(st.shape[1]), int(st.shape[2])
    out = torch.empty((2, 2, nf, n), dtype=torch.complex128, device=st.device)
    with torch.cuda.device(st.device):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(
            "rsk_stokes_to_coherency",
            _lib.load().rsk_stokes_to_coherency(_lib.ptr(st), nf, n, _lib.ptr(out), stream),
        )
    if single:
        out = out.reshape(2, 2)
    return out if want_torch else out.cpu().numpy()


def coherency_to_stokes(coherency_matrix):
    """Convert coherency matrix to vector of 4 Stokes parameter in order [I, Q, U, V] (utils.py:64-103).

    ``coherency_matrix``: complex, shape (2, 2) or (2, 2, ...) (numpy, torch or Quantity-like).  Returns float64 of shape
    (4,) or (4, ...): a torch CUDA tensor if the input was a torch tensor, otherwise a numpy array."""
    shape = tuple(coherency_matrix.shape)
    if len(shape) < 2 or shape[0] != 2 or shape[1] != 2:
        raise ValueError("First two dimensions of coherency_matrix must be length 2.")
    want_torch = isinstance(coherency_matrix, torch.Tensor)
    arr = coherency_matrix
    if hasattr(arr, "unit") and hasattr(arr, "value"):
        arr = arr.value
    if want_torch:
        coh = arr.to(dtype=torch.complex128)
        coh = coh if coh.is_cuda else coh.to("cuda")
    else:
        coh = torch.as_tensor(np.ascontiguousarray(arr, dtype=np.complex128)).to("cuda")
    coh = coh.contiguous()
    count = coh.numel() // 4
    out = torch.empty((4,) + shape[2:], dtype=torch.float64, device=coh.device)
    with torch.cuda.device(coh.device):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check("rsk_coherency_to_stokes",
                   _lib.load().rsk_coherency_to_stokes(_lib.ptr(coh), count, _lib.ptr(out), stream))
    return out if want_torch else out.cpu().numpy()


def jy_to_ksr(freqs):
    """Multiplicative factors taking [Jy] to [K sr] at ``freqs`` (Hz): 1e-26 c^2 / (2 k_B nu^2).

    This is astropy's ``brightness_temperature(freqs, beam_area=1 sr)`` equivalency with CODATA-2018
    constants, as used by utils.py:123-124.  Same type out as in (numpy or torch).
    """
    if hasattr(freqs, "to_value"):
        freqs = freqs.to_value("Hz")
    if isinstance(freqs, torch.Tensor):
        f = freqs.to(torch.float64)
        return 1e-26 * C_MPS**2 / (2.0 * K_B * f * f)
    f = np.atleast_1d(np.asarray(freqs, dtype=np.float64))
    return 1e-26 * C_MPS**2 / (2.0 * K_B * f**2)
