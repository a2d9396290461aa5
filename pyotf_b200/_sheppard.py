"""Host side of SheppardPSF (pyotf/otf.py:497-592): parameter marshalling for the K3 generator and
the shifted 3D inverse FFT that turns the spherical-cap OTF into the amplitude PSF."""
import ctypes as C

import numpy as np

from . import _engine, _fft, _lib


def _params(model):
    return _lib.SheppardParams(float(model.wl), float(model.na), float(model.ni), float(model.res), float(model.zres),
                               int(model.size), int(model.zsize), _engine._VEC[model.vec_corr],
                               _engine._COND[model.condition], int(bool(model.dual)),
                               _engine.dtype_code(model.precision))


def _nvec(model):
    return 1 if model.vec_corr == "none" else (6 if model.vec_corr == "total" else 2)


def otfa(model):
    """SheppardPSF.OTFa as a CUDA tensor (nvec, nz, N, N), fftshift-ed (otf.py:535-587)."""
    torch = _engine._torch()
    _, ct = _engine.torch_dtypes(model.precision)
    out = torch.empty((_nvec(model), model.zsize, model.size, model.size), dtype=ct, device="cuda")
    p = _params(model)
    _lib.check(_lib.load().pyotf_b200_sheppard_otfa(C.byref(p), _engine.ptr(out), C.c_void_p(0), _engine.stream_ptr()))
    return out


def _fused(model, want_psfa, want_psfi):
    """One pyotf_b200_sheppard_psf call without OTFa: the shell is generated inside the first FFT pass."""
    torch = _engine._torch()
    rt, ct = _engine.torch_dtypes(model.precision)
    shape = (_nvec(model), model.zsize, model.size, model.size)
    psfa_t = torch.empty(shape, dtype=ct, device="cuda") if want_psfa else None
    psfi_t = torch.empty(shape[1:], dtype=rt, device="cuda") if want_psfi else None
    p = _params(model)
    _lib.check(_lib.load().pyotf_b200_sheppard_psf(C.byref(p), C.c_void_p(0), _engine.ptr(psfa_t), _engine.ptr(psfi_t),
                                                   _engine.stream_ptr()))
    return psfa_t, psfi_t


def psfa(model):
    """SheppardPSF.PSFa = easy_ifft(OTFa, axes=(1, 2, 3)) (otf.py:589-592)."""
    if model._has_result("OTFa"):      # an OTFa was assigned / cached: transform exactly that array
        return _fft.shifted_fft_dev(model._device_result("OTFa"), axes=(1, 2, 3), inverse=True)
    return _fused(model, True, False)[0]


def psfi(model):
    if model._has_result("OTFa"):
        return _fft.abs2_sum0(model._device_result("PSFa"))
    return _fused(model, False, True)[1]


def gen_kr(model):
    """Sets kmag, dk, dkz, valid_points, kzz, kyy, kxx like SheppardPSF._gen_kr (otf.py:497-533)."""
    torch = _engine._torch()
    k = np.fft.fftfreq(model.size, model.res)
    kz = np.fft.fftfreq(model.zsize, model.zres)
    model.kmag = model.ni / model.wl
    model.dk, model.dkz = k[1] - k[0], kz[1] - kz[0]
    valid = torch.empty((model.zsize, model.size, model.size), dtype=torch.uint8, device="cuda")
    p = _params(model)
    _lib.check(_lib.load().pyotf_b200_sheppard_otfa(C.byref(p), C.c_void_p(0), _engine.ptr(valid), _engine.stream_ptr()))
    model.valid_points = valid.cpu().numpy().astype(bool)
    iz, iy, ix = np.nonzero(model.valid_points)
    model.kzz, model.kyy, model.kxx = kz[iz], k[iy], k[ix]
