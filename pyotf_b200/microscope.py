"""Microscope models on the device-resident PSFi (first slice of SURVEY.md section 8f, N4).

Mirrors ``pyotf.microscope`` (pyotf/microscope.py:49-227) for the widefield and the confocal microscope: same
constructor keywords, validation and cache behaviour; the arithmetic -- oversample roll + lateral binning +
normalisation (``WidefieldMicroscope.PSF``), the pinhole convolution times the excitation PSF
(``ConfocalMicroscope.model_psf``) and the centred 3D FFT (``OTF``) -- runs in CUDA on the stacks the PSF models
leave in HBM.  ``PSF`` / ``OTF`` / ``model_psf`` return numpy arrays like the reference; ``PSF_dev`` / ``OTF_dev`` /
``model_psf_dev`` give the CUDA tensors.  The SIM family (``BaseSIMMicroscope``, ``SIM2DMicroscope``,
``SIM3DMicroscope``, ``LatticeSIMMicroscope``, pyotf/microscope.py:265-425) is built the same way: the optional Wiener
filter of the OTF and the structured-illumination intensity are CUDA kernels on the resident stacks.  The apotome
model is not built: its radial average comes from the un-vendored ``dphtools.utils.radial_profile``.

dphtools.utils.bin_ndarray is not vendored by the reference; binning is a block sum (the default of that function),
which after the normalisation ``psf / psf.sum()`` is indistinguishable from a block mean.
"""
import copy
import ctypes as C
import logging

import numpy as np

from . import _engine, _fft, _lib
from .otf import HanserPSF, SheppardPSF
from .utils import NumericProperty

logger = logging.getLogger(__name__)

MODELS = {"hanser": HanserPSF, "sheppard": SheppardPSF}


def _choose_model(model):
    """Validate model (pyotf/microscope.py:39-46)."""
    try:
        return MODELS[model.lower()]
    except KeyError:
        raise ValueError(f"Model {model:} doesn't exist please choose one of: " + ", ".join(MODELS.keys()))


def _code(t):
    torch = _engine._torch()
    if t.dtype == torch.float32:
        return 0
    if t.dtype == torch.float64:
        return 1
    raise TypeError(f"expected a float32 / float64 CUDA tensor, not {t.dtype}")


def bin_lateral_dev(psf_dev, factor, shift=0, normalise=True):
    """np.roll(psf, (shift, shift), (1, 2)) summed over factor x factor lateral bins (and divided by the total)."""
    torch = _engine._torch()
    psf_dev = psf_dev.contiguous()
    nz, n, n2 = psf_dev.shape
    if n != n2 or n % factor:
        raise ValueError("the stack must be square with a size divisible by the bin factor")
    out = torch.empty((nz, n // factor, n // factor), dtype=psf_dev.dtype, device=psf_dev.device)
    _lib.check(_lib.load().pyotf_b200_bin_lateral(_engine.ptr(psf_dev), int(nz), int(n), int(factor), int(shift), _engine.ptr(out),
                                                  int(bool(normalise)), _code(psf_dev), _engine.stream_ptr()))
    return out


def disk_conv_mul_dev(em_dev, exc_dev, radius):
    """fftconvolve(em, disk(radius)[None], "same", axes=(1, 2)) * exc on CUDA tensors [nz, N, N]."""
    torch = _engine._torch()
    em_dev, exc_dev = em_dev.contiguous(), exc_dev.contiguous()
    if em_dev.shape != exc_dev.shape or em_dev.dtype != exc_dev.dtype:
        raise ValueError("emission and excitation stacks must have the same shape and dtype")
    nz, n, _ = em_dev.shape
    out = torch.empty_like(em_dev)
    _lib.check(_lib.load().pyotf_b200_disk_conv_mul(_engine.ptr(em_dev), _engine.ptr(exc_dev), int(nz), int(n), float(radius),
                                                    _engine.ptr(out), _code(em_dev), _engine.stream_ptr()))
    return out


class WidefieldMicroscope(object):
    """A base class for microscope models (pyotf/microscope.py:49-129)."""

    oversample_factor = NumericProperty(attr="_oversample_factor", vartype=int,
                                        doc="By how much do you want to oversample the simulation")

    def __init__(self, *, model, na, ni, wl, size, pixel_size, oversample_factor, **kwargs):
        # zsize and zres are set here because the simulation is not oversampled along z
        self.oversample_factor = oversample_factor
        self.psf_params = dict(na=na, ni=ni, wl=wl, zres=pixel_size, zsize=size)
        self.psf_params.update(kwargs)
        assert self.oversample_factor % 2 == 1, "oversample_factor must be odd"
        if oversample_factor == 1:
            self.psf_params["size"] = size
            self.psf_params["res"] = pixel_size
        elif oversample_factor > 1:
            self.psf_params["size"] = size * oversample_factor
            self.psf_params["res"] = pixel_size / oversample_factor
        else:
            raise ValueError("oversample_factor must be positive")
        self.oversample_factor = oversample_factor
        self.pixel_size = pixel_size
        self.model = _choose_model(model)(**self.psf_params)

    def __repr__(self):
        psf_kwargs = self.psf_params.copy()
        psf_kwargs.pop("res", None)
        psf_kwargs_str = ", ".join(f"{k}={v}" for k, v in psf_kwargs.items())
        return (f"{self.__class__.__name__}(model='{self.model.__class__.__name__[:-3].lower()}', "
                + f"{psf_kwargs_str}, pixel_size={self.pixel_size}, oversample_factor={self.oversample_factor})")

    def _attribute_changed(self):
        self.__dict__.pop("_PSF_dev", None)
        self.__dict__.pop("_OTF_dev", None)

    @property
    def model_psf_dev(self):
        """Internal model PSF (full grid resolution) as a CUDA tensor."""
        return self.model.PSFi_dev

    @property
    def model_psf(self):
        """Return internal model PSF (i.e. at full grid resolution)."""
        return _engine.to_host(self.model_psf_dev)

    @property
    def PSF_dev(self):
        if "_PSF_dev" not in self.__dict__:
            psf = self.model_psf_dev
            f = self.oversample_factor
            # an even oversampled grid has its peak in the upper left corner of the super pixel: roll by f // 2
            shift = f // 2 if (f > 1 and self.psf_params["size"] % 2 == 0) else 0
            self.__dict__["_PSF_dev"] = bin_lateral_dev(psf, f, shift, normalise=True)
        return self.__dict__["_PSF_dev"]

    @property
    def PSF(self):
        """Point spread function of the microscope."""
        return _engine.to_host(self.PSF_dev)

    @property
    def OTF_dev(self):
        if "_OTF_dev" not in self.__dict__:
            self.__dict__["_OTF_dev"] = _fft.shifted_fft_dev(self.PSF_dev, axes=None, inverse=False)
        return self.__dict__["_OTF_dev"]

    @property
    def OTF(self):
        """Optical transfer function of the microscope."""
        return _engine.to_host(self.OTF_dev)

    def plot_psf(self, **kwargs):
        raise NotImplementedError("plotting is out of scope for pyotf_b200; use pyotf.display on .PSF")

    def plot_otf(self, **kwargs):
        raise NotImplementedError("plotting is out of scope for pyotf_b200; use pyotf.display on .OTF")


class ConfocalMicroscope(WidefieldMicroscope):
    """A class representing a confocal microscope (pyotf/microscope.py:190-227)."""

    pinhole_size = NumericProperty(attr="_pinhole_size", vartype=(float, int),
                                   doc="Size of the pinhole (in airy units relative to emission wavelength")
    wl_exc = NumericProperty(attr="_wl_exc", vartype=float, doc="The excitation wavelength")

    def __init__(self, *, wl_exc, pinhole_size, **kwargs):
        super().__init__(**kwargs)
        self.pinhole_size = pinhole_size
        # the excitation PSF: same model at the excitation wavelength
        self.model_exc = copy.copy(self.model)
        self.model_exc._attribute_changed()
        self.model_exc.wl = wl_exc

    def __repr__(self):
        return super().__repr__()[:-1] + f", wl_exc={self.model_exc.wl}, pinhole_size={self.pinhole_size})"

    @property
    def model_psf_dev(self):
        """Oversampled confocal PSF: (detection PSF * pinhole) x excitation PSF."""
        airy_unit = 1.22 * self.model.wl / self.model.na / self.model.res      # in pixels
        logger.debug(f"Airy unit = {airy_unit:}")
        pixel_pinhole_radius = self.pinhole_size * airy_unit / 2
        em, exc = self.model.PSFi_dev, self.model_exc.PSFi_dev
        if pixel_pinhole_radius > 1.5:
            return disk_conv_mul_dev(em, exc, pixel_pinhole_radius)
        return em * exc


def wiener_otf_dev(otf_dev, wiener):
    """abs(otf) ** 2 / (abs(otf) ** 2 + max * wiener ** 2), or the support (> 1e-16) for wiener == 0; real CUDA tensor."""
    torch = _engine._torch()
    otf_dev = otf_dev.contiguous()
    rt = torch.float32 if otf_dev.dtype == torch.complex64 else torch.float64
    out = torch.empty(otf_dev.shape, dtype=rt, device=otf_dev.device)
    _lib.check(_lib.load().pyotf_b200_wiener_otf(_engine.ptr(otf_dev), int(otf_dev.numel()), float(wiener), _engine.ptr(out),
                                                 0 if rt == torch.float32 else 1, _engine.stream_ptr()))
    return out


def sim_psf_dev(psf_dev, res, zres, ni, wl_exc, na_exc, orientations, coherent, dc, dc_suppress):
    """psf * structured-illumination intensity (pyotf/microscope.py:357-401) on a CUDA stack [nz, N, N] (real, or complex:
    its real part is used)."""
    torch = _engine._torch()
    psf_dev = psf_dev.contiguous()
    cplx = psf_dev.is_complex()
    rt = torch.float32 if psf_dev.dtype in (torch.float32, torch.complex64) else torch.float64
    nz, n, n2 = psf_dev.shape
    if n != n2:
        raise ValueError("the stack must be square")
    out = torch.empty((nz, n, n), dtype=rt, device=psf_dev.device)
    ori = np.ascontiguousarray(np.asarray(list(orientations), dtype=np.float64))
    _lib.check(_lib.load().pyotf_b200_sim_psf(_engine.ptr(psf_dev), int(cplx), int(nz), int(n), float(res), float(zres), float(ni),
                                              float(wl_exc), float(na_exc), ori.ctypes.data_as(C.POINTER(C.c_double)), int(ori.size),
                                              int(bool(coherent)), int(bool(dc)), int(bool(dc_suppress)), _engine.ptr(out),
                                              0 if rt == torch.float32 else 1, _engine.stream_ptr()))
    return out


class BaseSIMMicroscope(WidefieldMicroscope):
    """A base class for SIM and SIM like microscopes (pyotf/microscope.py:265-401)."""

    na_exc = NumericProperty(attr="_na_exc", vartype=float, doc="The excitation NA")
    wl_exc = NumericProperty(attr="_wl_exc", vartype=float, doc="The excitation wavelength")
    coherent = NumericProperty(attr="_coherent", vartype=bool, doc="Treat the orientations coherently?")
    dc = NumericProperty(attr="_dc", vartype=bool, doc="Include the DC component")
    dc_suppress = NumericProperty(attr="_dc_suppress", vartype=bool, doc="Suppress the DC component")

    def __init__(self, *, na_exc, wl_exc, wiener, coherent, dc, dc_suppress, orientations, **kwargs):
        super().__init__(**kwargs)
        if na_exc is None:
            na_exc = self.psf_params["na"]
        self.na_exc = na_exc
        if wl_exc is None:
            wl_exc = self.psf_params["wl"]
        self.wl_exc = wl_exc
        self.coherent = coherent
        self.dc = dc
        self.dc_suppress = dc_suppress
        self.orientations = orientations
        self.wiener = wiener
        if self.wiener < 0:
            raise ValueError(f"self.wiener is {self.wiener:} which should be greater than zero")

    def __repr__(self):
        return (super().__repr__()[:-1]
                + f", na_exc={self.na_exc}, wl_exc={self.wl_exc}, wiener={self.wiener}, "
                + f"coherent={self.coherent}, dc={self.dc}, dc_suppress={self.dc_suppress}, "
                + f"orientations={self.orientations!r})")

    @property
    def model_psf_dev(self):
        """Oversampled SIM PSF (pyotf/microscope.py:316-401)."""
        if self.wiener is None:
            psf = self.model.PSFi_dev
        elif self.wiener >= 0:
            # https://en.wikipedia.org/wiki/Wiener_deconvolution ; the PSF is real, the imaginary part is discarded
            psf = _fft.shifted_fft_dev(wiener_otf_dev(self.model.OTFi_dev, self.wiener), axes=None, inverse=True)
        else:
            raise ValueError(f"self.wiener is {self.wiener:} which should be greater than zero")
        return sim_psf_dev(psf, self.psf_params["res"], self.psf_params["zres"], self.psf_params["ni"], self.wl_exc, self.na_exc,
                           self.orientations, self.coherent, self.dc, self.dc_suppress)


class SIM2DMicroscope(BaseSIMMicroscope):
    """A class for 2D-SIM, including optical sectioning SIM (pyotf/microscope.py:404-408)."""

    def __init__(self, **kwargs):
        super().__init__(coherent=False, dc=False, **kwargs)


class SIM3DMicroscope(BaseSIMMicroscope):
    """A class for 3D-SIM using multiple pattern orientations (pyotf/microscope.py:411-415)."""

    def __init__(self, **kwargs):
        super().__init__(coherent=False, dc=True, **kwargs)


class LatticeSIMMicroscope(BaseSIMMicroscope):
    """A class for Zeiss Lattice SIM (pyotf/microscope.py:418-425)."""

    def __init__(self, **kwargs):
        super().__init__(coherent=True, dc=True, orientations=(0, np.pi / 2), **kwargs)
