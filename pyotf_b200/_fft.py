"""Device N-d FFT with folded shifts (host wrapper of pyotf_b200_fftn / pyotf_b200_abs2_sum0)."""
import ctypes as C

import numpy as np

from . import _engine, _lib


def _norm_axes(ndim, axes):
    if axes is None:
        return list(range(ndim))
    if isinstance(axes, int):
        axes = (axes,)
    return [a % ndim for a in axes]


def fftn_dev(x, axes=None, inverse=False, shift_in=False, shift_out=False):
    """FFT of a CUDA tensor (real or complex, fp32/fp64) -> complex CUDA tensor of the same shape."""
    torch = _engine._torch()
    x = x.contiguous()
    if x.dtype in (torch.float32, torch.complex64):
        code, ct = 0, torch.complex64
    elif x.dtype in (torch.float64, torch.complex128):
        code, ct = 1, torch.complex128
    else:
        raise TypeError(f"unsupported dtype {x.dtype}")
    real_in = not x.is_complex()
    axes = _norm_axes(x.dim(), axes)
    out = torch.empty(x.shape, dtype=ct, device=x.device)
    shape = (C.c_longlong * x.dim())(*x.shape)
    ax = (C.c_int * len(axes))(*axes)
    _lib.check(_lib.load().pyotf_b200_fftn(_engine.ptr(x), _engine.ptr(out), x.dim(), shape, ax, len(axes),
                                           int(inverse), int(real_in), int(shift_in), int(shift_out), code,
                                           _engine.stream_ptr()))
    return out


def shifted_fft_dev(x, axes=None, inverse=False):
    """easy_fft / easy_ifft on a CUDA tensor (pyotf/utils.py:15-22)."""
    return fftn_dev(x, axes, inverse, True, True)


def shifted_fft(data, axes=None, inverse=False):
    """easy_fft / easy_ifft on a host array; returns a numpy array (complex64 for 32-bit input)."""
    torch = _engine._torch()
    a = np.asarray(data)
    if a.dtype in (np.float32, np.complex64):
        pass
    elif np.iscomplexobj(a):
        a = a.astype(np.complex128, copy=False)
    else:
        a = a.astype(np.float64, copy=False)
    x = torch.as_tensor(np.ascontiguousarray(a), device="cuda")
    return shifted_fft_dev(x, axes, inverse).cpu().numpy()


def abs2_sum0(a):
    """sum over axis 0 of |a|^2 for a complex CUDA tensor (nvec, ...) -> real tensor (...)."""
    torch = _engine._torch()
    a = a.contiguous()
    code = 0 if a.dtype == torch.complex64 else 1
    rt = torch.float32 if code == 0 else torch.float64
    out = torch.empty(a.shape[1:], dtype=rt, device=a.device)
    n = out.numel()
    _lib.check(_lib.load().pyotf_b200_abs2_sum0(_engine.ptr(a), int(a.shape[0]), n, _engine.ptr(out), code,
                                                _engine.stream_ptr()))
    return out
