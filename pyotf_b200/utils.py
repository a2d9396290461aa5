"""Utility layer of the drop-in API (mirrors pyotf/utils.py).

``easy_fft`` / ``easy_ifft`` run the hand-written device FFT (csrc/fftnd.cuh) when the input is
large enough to matter and is a shape the device path supports; ``psqrt`` / ``cart2pol`` /
``NumericProperty`` are host-side parameter helpers with the reference's exact semantics.
The data-prep helpers (``center_data``, ``remove_bg``, ``prep_data_for_PR``) are one-off host
preparation steps that sit *before* the hot path.
"""
import numpy as np


def easy_fft(data, axes=None):
    """Centered forward FFT, ``fftshift(fftn(ifftshift(x)))`` (pyotf/utils.py:15-17)."""
    from . import _fft

    return _fft.shifted_fft(data, axes=axes, inverse=False)


def easy_ifft(data, axes=None):
    """Centered inverse FFT, ``ifftshift(ifftn(fftshift(x)))`` (pyotf/utils.py:20-22)."""
    from . import _fft

    return _fft.shifted_fft(data, axes=axes, inverse=True)


def cart2pol(y, x):
    """Return ``(rho, theta)`` for cartesian ``(y, x)`` (pyotf/utils.py:25-29)."""
    return np.hypot(y, x), np.arctan2(y, x)


class NumericProperty(property):
    """A validated numeric attribute (pyotf/utils.py:32-73).

    Wrong type -> TypeError, negative -> ValueError; a *changed* value invalidates the owner's
    cached results through ``owner._attribute_changed()``.
    """

    def __init__(self, fget=None, fset=None, fdel=None, doc=None, attr=None, vartype=None):
        if attr is not None and vartype is not None:
            self.attr, self.vartype = attr, vartype

            def fget(owner, _name=attr):
                return getattr(owner, _name)

            def fset(owner, value, _name=attr, _types=vartype):
                if not isinstance(value, _types):
                    raise TypeError(f"{_name} must be an {_types}, var = {value}")
                if value < 0:
                    raise ValueError(f"{_name} must be larger than 0")
                if getattr(owner, _name, None) != value:
                    setattr(owner, _name, value)
                    owner._attribute_changed()

        super().__init__(fget, fset, fdel, doc)


def center_data(data):
    """Roll ``data`` so that its maximum sits at the array centre (pyotf/utils.py:76-124).

    >>> center_data(np.array(((1, 0, 0), (0, 0, 0), (0, 0, 0))))
    array([[0, 0, 0],
           [0, 1, 0],
           [0, 0, 0]])
    """
    peak = np.unravel_index(np.argmax(data), data.shape)
    shifts = tuple(n // 2 - p for p, n in zip(peak, data.shape))
    return np.roll(data, shifts, axis=tuple(range(data.ndim)))


def remove_bg(data, multiplier):
    """Subtract ``multiplier`` x the mode of integer ``data`` (pyotf/utils.py:127-147).

    >>> remove_bg(np.array([1, 1, 1, 0, 2]), 1)
    array([ 0.,  0.,  0., -1.,  1.])
    """
    data = np.asarray(data)
    if not np.issubdtype(data.dtype, np.integer):
        raise ValueError("Floating point numbers unsupported")
    if (data < 0).any():
        raise ValueError("Negative values unsupported")
    mode = int(np.bincount(data.ravel()).argmax())
    return data - multiplier * float(mode)


def psqrt(data):
    """Square root with negatives clamped to zero (pyotf/utils.py:150-158).

    >>> psqrt((-4, 4))
    array([0., 2.])
    """
    return np.sqrt(np.fmax(0, data))


def slice_maker(centers, widths):
    """Slices of ``widths`` centred on ``centers`` (clean-room stand-in for dphtools.slice_maker).

    ``widths`` may be a scalar (applied to every axis).  A window is ``[c - w // 2, c - w // 2 + w)``; only its START is
    clipped at 0, so a window that hangs over the lower edge comes back shorter (the published dphtools rule as far as
    it is known here), and one that ends at or before 0 is empty.  UNPINNED: dphtools is not vendored by the reference,
    so edge-clipped windows could not be compared with the original; windows that fit are unambiguous and are what the
    reference's own shape / centre tests exercise (tests/test_api_gpu.py).  In ``prep_data_for_PR`` the centre is
    ``(n + 1) // 2``: a clipped start there implies a stop beyond the array, so either clipping rule crops the same
    elements (tests/test_utils_cpu.py).
    """
    centers = np.atleast_1d(centers)
    widths = np.broadcast_to(np.atleast_1d(widths), centers.shape)
    out = []
    for c, w in zip(centers, widths):
        w = int(round(float(w)))
        if w < 0:
            raise ValueError("width must be non-negative")
        lo = int(round(float(c))) - w // 2
        hi = lo + w
        if hi <= 0:
            lo, hi = 0, 0
        out.append(slice(max(lo, 0), hi))
    return tuple(out)


def fft_pad(data, newshape, mode="constant"):
    """Pad ``data`` symmetrically (FFT-centre preserving) up to ``newshape``.

    Clean-room stand-in for dphtools.utils.fft_pad: the element at index ``n // 2`` of every
    axis stays at index ``new_n // 2``.  UNPINNED for the same reason as ``slice_maker``; shrinking raises instead of
    cropping (``prep_data_for_PR`` crops through ``slice_maker`` before it pads).
    """
    data = np.asarray(data)
    pads = []
    for n, m in zip(data.shape, newshape):
        if m < n:
            raise ValueError("fft_pad cannot crop")
        before = m // 2 - n // 2
        pads.append((before, m - n - before))
    return np.pad(data, pads, mode=mode)


def prep_data_for_PR_dev(data, xysize=None, multiplier=1.05, precision=None):
    """``prep_data_for_PR`` on the GPU for unsigned 16-bit stacks; returns a CUDA tensor (nz, xysize, xysize).

    ``data`` is a uint16 numpy array or CUDA tensor (nz, ny, nx).  Histogram mode, arg-max, roll / pad / crop and the
    clip run on the device (``pyotf_b200_prep_data_u16``); values are bit-identical to the host function in fp64.
    The result can be handed to ``retrieve_phase`` directly, so a measured stack never visits the host as floats.
    """
    import ctypes as C

    from . import _engine, _lib

    torch = _engine._torch()
    prec = precision or _engine.get_precision()
    if isinstance(data, np.ndarray):
        if data.dtype != np.uint16:
            raise ValueError("prep_data_for_PR_dev handles uint16 stacks; use prep_data_for_PR for other dtypes")
        dev = torch.as_tensor(np.ascontiguousarray(data).view(np.int16), device="cuda")
    else:
        dev = data.contiguous()
        if dev.dtype not in (torch.uint16, torch.int16):
            raise ValueError("prep_data_for_PR_dev handles uint16 stacks")
    nz, ny, nx = (int(v) for v in dev.shape)
    size = max(ny, nx) if xysize is None else int(xysize)
    out = torch.empty((nz, size, size), dtype=torch.float32 if prec == "fp32" else torch.float64, device="cuda")
    mode = C.c_int(0)
    _lib.check(_lib.load().pyotf_b200_prep_data_u16(_engine.ptr(dev), nz, ny, nx, size, float(multiplier),
                                                    _engine.dtype_code(prec), _engine.ptr(out), C.byref(mode),
                                                    _engine.stream_ptr()))
    return out


def prep_data_for_PR(data, xysize=None, multiplier=1.05):
    """Background-subtract, pad/crop to ``xysize`` and clip at zero (pyotf/utils.py:161-201)."""
    nz, ny, nx = data.shape
    cleaned = remove_bg(data, multiplier)
    if xysize is None:
        xysize = max(ny, nx)
    if xysize == ny == nx:
        out = cleaned
    elif xysize >= max(ny, nx):
        out = fft_pad(cleaned, (nz, xysize, xysize), mode="constant")
    else:
        window = slice_maker(((ny + 1) // 2, (nx + 1) // 2), xysize)
        out = center_data(cleaned)[(Ellipsis,) + window]
    return np.fmax(0, out)
