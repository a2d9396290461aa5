"""PSF / OTF models with pyotf's object API, computed by sm_100a CUDA kernels.

Drop-in for ``pyotf.otf`` (reference: pyotf/otf.py): same class names, constructor
signatures, validating setters, exception types, cache invalidation and array conventions
(``PSFa`` is ``(nvec, nz, N, N)`` complex, spatially fftshift-ed; ``PSFi`` is ``(nz, N, N)``
real; ``OTFa`` / ``OTFi`` are centred 3D spectra).  Arrays are returned as numpy arrays;
``<name>_dev`` (e.g. ``model.PSFi_dev``) gives the device-resident ``torch`` tensor without a
device-to-host copy.  ``precision="fp32"|"fp64"`` selects the kernel arithmetic (default: the
package-wide ``pyotf_b200.get_precision()``).
"""
import copy
import logging

import numpy as np

from . import _engine
from .utils import NumericProperty, easy_fft, easy_ifft, psqrt, slice_maker  # noqa: F401
from .zernike import name2noll, noll2degrees

logger = logging.getLogger(__name__)

_RESULTS = ("PSFa", "OTFa", "PSFi", "OTFi")


def _result_property(name, doc):
    """A cached array attribute: computed on first read, assignable, deletable.

    Two per-object caches back it: ``_device`` (CUDA tensors, what the kernels produce) and
    ``_results`` (numpy copies, made on first host access).
    """

    def getter(self):
        host = self.__dict__.setdefault("_results", {})
        if name not in host:
            host[name] = _engine.to_host(self._device_result(name))
        return host[name]

    def setter(self, value):
        self.__dict__.setdefault("_results", {}).pop(name, None)
        self.__dict__.setdefault("_device", {}).pop(name, None)
        if isinstance(value, np.ndarray):
            self.__dict__["_results"][name] = value
        else:
            self.__dict__["_device"][name] = value

    def deleter(self):
        a = self.__dict__.setdefault("_results", {}).pop(name, None)
        b = self.__dict__.setdefault("_device", {}).pop(name, None)
        if a is None and b is None:
            raise AttributeError(name)

    return property(getter, setter, deleter, doc)


class BasePSF(object):
    """Parameter holder shared by the PSF models (pyotf/otf.py:30-262).

    wl, na, ni : wavelength, numerical aperture, refractive index
    res, size  : lateral pixel size (same unit as wl) and lateral size in pixels
    zres, zsize: axial step and number of planes (default to res / size)
    vec_corr   : "none" | "x" | "y" | "z" | "total"
    condition  : "none" | "sine" | "herschel"
    """

    wl = NumericProperty(attr="_wl", vartype=(float, int), doc="Wavelength of emission, in nm")
    na = NumericProperty(attr="_na", vartype=(float, int), doc="Numerical Aperature")
    ni = NumericProperty(attr="_ni", vartype=(float, int), doc="Refractive index")
    size = NumericProperty(attr="_size", vartype=int, doc="x/y size")
    zsize = NumericProperty(attr="_zsize", vartype=int, doc="z size")

    def __init__(self, wl, na, ni, res, size, zres=None, zsize=None, vec_corr="none", condition="sine",
                 *, precision=None):
        self.precision = precision
        self.wl = wl
        self.na = na
        self.ni = ni
        self.res = res
        self.size = size
        self.zres = res if zres is None else zres
        self.zsize = size if zsize is None else zsize
        self.vec_corr = vec_corr
        self.condition = condition

    def __repr__(self):
        return (
            f"{self.__class__.__name__}(wl={self.wl}, na={self.na}, ni={self.ni},"
            f" res={self.res}, size={self.size}, zres={self.zres}, zsize={self.zsize},"
            f" vec_corr='{self.vec_corr}', condition='{self.condition}')"
        )

    def __copy__(self):
        new = self.__class__.__new__(self.__class__)
        new.__dict__.update(self.__dict__)
        # cached results are per-object (the arrays themselves are shared, like the reference's
        # shallow copy of cached_property values)
        new.__dict__["_results"] = dict(self.__dict__.get("_results", {}))
        new.__dict__["_device"] = dict(self.__dict__.get("_device", {}))
        return new

    # -- precision ------------------------------------------------------------------------
    @property
    def precision(self):
        """Kernel arithmetic: "fp32" or "fp64"."""
        return self._precision or _engine.get_precision()

    @precision.setter
    def precision(self, value):
        if value not in (None, "fp32", "fp64"):
            raise ValueError("precision must be None, 'fp32' or 'fp64'")
        self._precision = value
        self._attribute_changed()

    # -- cache ----------------------------------------------------------------------------
    def _attribute_changed(self):
        """Drop every cached result (pyotf/otf.py:114-124)."""
        self.__dict__["_results"] = {}
        self.__dict__["_device"] = {}
        self.__dict__["_applied_pupil"] = None

    def _has_result(self, name):
        return name in self.__dict__.get("_results", {}) or name in self.__dict__.get("_device", {})

    def _device_result(self, name):
        """CUDA tensor of a result, computing (or uploading an assigned numpy array) if needed."""
        dev = self.__dict__.setdefault("_device", {})
        if name not in dev:
            host = self.__dict__.setdefault("_results", {})
            if name in host:
                import torch

                dev[name] = torch.as_tensor(host[name], device="cuda")
            else:
                dev[name] = getattr(self, "_compute_" + name)()
        return dev[name]

    PSFa_dev = property(lambda self: self._device_result("PSFa"), doc="PSFa as a CUDA tensor")
    PSFi_dev = property(lambda self: self._device_result("PSFi"), doc="PSFi as a CUDA tensor")
    OTFa_dev = property(lambda self: self._device_result("OTFa"), doc="OTFa as a CUDA tensor")
    OTFi_dev = property(lambda self: self._device_result("OTFi"), doc="OTFi as a CUDA tensor")

    # -- validated attributes ---------------------------------------------------------------
    @property
    def zres(self):
        """Z resolution."""
        return self._zres

    @zres.setter
    def zres(self, value):
        if not value > 0:
            raise ValueError("zres must be positive")
        self._zres = value
        self._attribute_changed()

    @property
    def res(self):
        """X/Y resolution; must be finer than the Abbe limit wl / (2 na) (pyotf/otf.py:139-163)."""
        return self._res

    @res.setter
    def res(self, value):
        abbe_limit = 1 / (2 * self.na / self.wl)
        if value >= abbe_limit:
            raise ValueError(f"{value} is larger than the Abbe Limit, try a number smaller than {abbe_limit}")
        if value >= abbe_limit / 2:
            logger.info(f"res has been set to {value} which is greater than the Nyquist limit of {abbe_limit / 2}")
        self._res = value
        self._attribute_changed()

    @property
    def vec_corr(self):
        """Vectorial correction: "none", "x", "y", "z" or "total"."""
        return self._vec_corr

    @vec_corr.setter
    def vec_corr(self, value):
        valid = {"none", "x", "y", "z", "total"}
        if value not in valid:
            raise ValueError("Vector correction must be one of {}".format(", ".join(valid)))
        self._vec_corr = value
        self._attribute_changed()

    @property
    def condition(self):
        """Imaging condition: "none", "sine" or "herschel"."""
        return self._condition

    @condition.setter
    def condition(self, value):
        valid = {"none", "sine", "herschel"}
        if value not in valid:
            raise ValueError("Condition must be one of {}".format(", ".join(valid)))
        self._condition = value
        self._attribute_changed()

    # -- results ----------------------------------------------------------------------------
    OTFa = _result_property("OTFa", "Amplitude OTF (coherent transfer function), complex array.")
    PSFa = _result_property("PSFa", "Amplitude PSF, complex array.")
    PSFi = _result_property("PSFi", "Intensity PSF, real array.")
    OTFi = _result_property("OTFi", "Intensity OTF, complex array.")

    def _compute_OTFa(self):
        raise NotImplementedError

    def _compute_PSFa(self):
        raise NotImplementedError

    def _compute_PSFi(self):
        # sum over polarisation components of |PSFa|^2 (pyotf/otf.py:204-207)
        from . import _fft

        return _fft.abs2_sum0(self._device_result("PSFa"))

    def _compute_OTFi(self):
        # centred 3D FFT of PSFi (pyotf/otf.py:209-212)
        from . import _fft

        return _fft.shifted_fft_dev(self._device_result("PSFi"), axes=None, inverse=False)

    def _validate_zrange(self):
        try:
            steps = np.diff(self.zrange)
            if len(steps) < 2 or not np.allclose(steps, steps.mean()):
                raise RuntimeError(f"{self} doesn't have uniform z-steps ---> {steps}")
        except AttributeError:
            pass

    def plot_psf(self, **kwargs):
        """Plotting is outside the hot path."""
        raise NotImplementedError("plotting is out of scope for pyotf_b200; use pyotf.display on .PSFi")

    def plot_otf(self, **kwargs):
        """Plotting is outside the hot path."""
        raise NotImplementedError("plotting is out of scope for pyotf_b200; use pyotf.display on .OTFi")


class HanserPSF(BasePSF):
    """Scalar / vectorial pupil-function PSF model (Hanser et al. 2004; pyotf/otf.py:265-443).

    zrange : array-like, optional keyword -- explicit plane positions (same unit as wl)
             instead of the regular ``zsize`` x ``zres`` grid.
    """

    def __init__(self, *args, zrange=None, **kwargs):
        super().__init__(*args, **kwargs)
        if zrange is None:
            self._gen_zrange()
        else:
            self.zrange = zrange

    def __repr__(self):
        return super().__repr__()[:-1] + f", zrange={self.zrange!r})"

    def _gen_zrange(self):
        """Planes at (arange(zsize) - (zsize + 1) // 2) * zres (pyotf/otf.py:296-298)."""
        self.zrange = (np.arange(self.zsize) - (self.zsize + 1) // 2) * self.zres

    @BasePSF.zsize.setter
    def zsize(self, value):
        BasePSF.zsize.fset(self, value)
        if hasattr(self, "_zres"):
            self._gen_zrange()

    @BasePSF.zres.setter
    def zres(self, value):
        BasePSF.zres.fset(self, value)
        if hasattr(self, "_zsize"):
            self._gen_zrange()

    @property
    def zrange(self):
        """Plane positions over which the PSF is calculated."""
        return self._zrange

    @zrange.setter
    def zrange(self, value):
        z = np.asarray(value)
        if not z.shape:
            z = z.reshape(1)
        self._zrange = z
        self._attribute_changed()

    # -- hooks used by phase retrieval / aberration code (pyotf/otf.py:335-360) ------------
    def _gen_kr(self):
        """Pupil-plane coordinates (unshifted): ``_k, _kr, _phi, _kmag, _kz``."""
        key = (self.wl, self.ni, self.res, self.size)
        if self.__dict__.get("_kr_key") == key:         # same geometry as the last call: the arrays are still right
            return
        self._k = np.fft.fftfreq(self.size, self.res)
        kr, phi, kz = _engine.hanser_kspace(self.wl, self.ni, self.res, self.size)
        self._kr, self._phi, self._kz = kr.cpu().numpy(), phi.cpu().numpy(), kz.cpu().numpy()
        self._kmag = self.ni / self.wl
        self.__dict__["_kr_key"] = key

    def _gen_pupil(self):
        """Ideal pupil: complex ones where kr <= na / wl (pyotf/otf.py:346-355)."""
        return (self._kr <= self._na / self._wl).astype(complex)

    def _calc_defocus(self):
        """exp(2 pi i kz z) for every plane, (nz, N, N) complex128 (pyotf/otf.py:357-360)."""
        kz = _engine.to_device_f64(self._kz)
        return _engine.hanser_defocus(kz, _engine.to_device_f64(self.zrange)).cpu().numpy()

    # -- K1 ---------------------------------------------------------------------------------
    def _check_size(self):
        n = self.size
        if n < 2 or n > 16384:
            raise NotImplementedError(f"size={n}: the CUDA path covers sizes 2..16384 (there is no CPU fallback)")

    def _handle_and_pupil(self, pupil_base):
        """Resolve (HanserHandle, compact pupil tensor or None) for a pupil_base argument."""
        self._check_size()
        args = (self.wl, self.na, self.ni, self.res, self.size, self.vec_corr, self.condition, self.precision)
        if pupil_base is None:
            return _engine.hanser_handle(*args, -1), None
        if isinstance(pupil_base, _CompactPupil):
            return pupil_base.handle(self), pupil_base.tensor
        if isinstance(pupil_base, np.ndarray) or not hasattr(pupil_base, "device"):
            pupil_base = np.asarray(pupil_base)
            assert pupil_base.ndim == 2, f"`pupil_base` is wrong shape: {pupil_base.shape}"
            pupil_dev = _engine.to_device_c128(pupil_base)
        else:
            assert pupil_base.dim() == 2, f"`pupil_base` is wrong shape: {tuple(pupil_base.shape)}"
            import torch

            pupil_dev = pupil_base.to(device="cuda", dtype=torch.complex128).contiguous()
        if tuple(pupil_dev.shape) != (self.size, self.size):
            raise ValueError(f"pupil shape {tuple(pupil_dev.shape)} does not match size {self.size}")
        h = _engine.hanser_handle(*args, _engine.pupil_support(pupil_dev))
        return h, h.pack_pupil(pupil_dev[None])

    def _run(self, pupil_base, want_psfa, want_psfi):
        h, pupilc = self._handle_and_pupil(pupil_base)
        # zrange stays on the host, as in the reference: pyotf_b200_hanser_psf_zhost
        psfa, psfi = h.psf(np.asarray(self.zrange, dtype=np.float64), pupilc, want_psfa=want_psfa, want_psfi=want_psfi)
        return (psfa[0] if psfa is not None else None), (psfi[0] if psfi is not None else None)

    def _gen_psf(self, pupil_base=None):
        """Amplitude PSF for ``pupil_base`` (unshifted N x N; ideal pupil if None).

        Replaces pyotf/otf.py:362-428 with the fused kernel K1.  Like the reference it clears
        the cached results first.  Returns a numpy array (nvec, nz, N, N).
        """
        self._attribute_changed()
        psfa, _ = self._run(pupil_base, True, False)
        return psfa.cpu().numpy()

    def apply_pupil(self, pupil):
        """Use ``pupil`` (unshifted N x N complex) instead of the ideal one (pyotf/otf.py:430-433)."""
        if not isinstance(pupil, _CompactPupil):
            ndim = pupil.ndim if hasattr(pupil, "ndim") else np.asarray(pupil).ndim
            assert ndim == 2, f"`pupil_base` is wrong shape: {getattr(pupil, 'shape', None)}"
        self._attribute_changed()
        self.__dict__["_applied_pupil"] = pupil

    def _compute_PSFa(self):
        return self._run(self.__dict__.get("_applied_pupil"), True, False)[0]

    def _compute_PSFi(self):
        if self._has_result("PSFa"):
            return super()._compute_PSFi()  # a PSFa was assigned / already computed: reduce it
        return self._run(self.__dict__.get("_applied_pupil"), False, True)[1]

    def _compute_OTFa(self):
        # centred FFT over (z, y, x) of every component (pyotf/otf.py:435-438)
        from . import _fft

        return _fft.shifted_fft_dev(self._device_result("PSFa"), axes=(1, 2, 3), inverse=False)


class _CompactPupil:
    """A device-resident pupil already in a handle's compact K x K layout (internal)."""

    def __init__(self, tensor, handle_args):
        self.tensor, self.handle_args = tensor, handle_args
        self.ndim = 2

    def handle(self, model):
        args = (model.wl, model.na, model.ni, model.res, model.size, model.vec_corr, model.condition,
                model.precision)
        if args[:5] + (args[7],) != self.handle_args:
            raise ValueError("model optics changed since the pupil was built")
        return _engine.hanser_handle(*args, -2)


class SheppardPSF(BasePSF):
    """3D spherical-shell PSF model (Arnison & Sheppard 2002; pyotf/otf.py:446-592).

    dual : bool -- simulate opposing objectives (mirrored cap).
    """

    dual = NumericProperty(attr="_dual", vartype=bool, doc="Simulate dual objectives")

    def __init__(self, *args, dual=False, **kwargs):
        super().__init__(*args, **kwargs)
        self.dual = dual

    def __repr__(self):
        return super().__repr__()[:-1] + f", dual={self.dual})"

    @property
    def dual(self):
        """Simulate opposing objectives."""
        return self._dual

    @dual.setter
    def dual(self, value):
        if not isinstance(value, bool):
            raise TypeError("`dual` must be a boolean")
        self._dual = value
        self._attribute_changed()

    @BasePSF.zres.setter
    def zres(self, value):
        # axial Nyquist limit of the amplitude shell (pyotf/otf.py:484-495)
        max_val = self.wl / 2 / self.ni
        if value >= max_val:
            raise ValueError(f"{value} is too large try a number smaller than {max_val}")
        BasePSF.zres.fset(self, value)

    def _gen_kr(self):
        from . import _sheppard

        _sheppard.gen_kr(self)

    def _gen_otf(self):
        from . import _sheppard

        self._attribute_changed()
        return _sheppard.otfa(self)

    def _compute_OTFa(self):
        from . import _sheppard

        return _sheppard.otfa(self)

    def _compute_PSFa(self):
        from . import _sheppard

        return _sheppard.psfa(self)

    def _compute_PSFi(self):
        from . import _sheppard

        if self._has_result("PSFa"):
            return super()._compute_PSFi()
        return _sheppard.psfi(self)


def apply_aberration(model, mcoefs, pcoefs):
    """Return a copy of ``model`` aberrated by Noll-ordered Zernike coefficients (pyotf/otf.py:595-645).

    mcoefs / pcoefs : (n,) magnitude / phase coefficients; either may be None.
    The pupil ``(sum m_j Z_j + |ideal|) * exp(i sum p_j Z_j)`` is built on the GPU directly
    in K1's compact layout; nothing O(N^2) touches the host.
    """
    assert isinstance(model, HanserPSF), "Model must be a HanserPSF"
    model = copy.copy(model)
    if mcoefs is None and pcoefs is None:
        logger.warning("No abberation applied")
        return model
    if mcoefs is None:
        mcoefs = np.zeros_like(pcoefs)
    if pcoefs is None:
        pcoefs = np.zeros_like(mcoefs)
    assert len(mcoefs) == len(pcoefs), "Coefficient lengths don't match"
    model.apply_pupil(_aberrated_pupils(model, np.asarray(mcoefs, float)[None], np.asarray(pcoefs, float)[None]))
    return model


def _aberrated_pupils(model, mcoefs, pcoefs):
    model._check_size()
    nmodes = (mcoefs if mcoefs is not None else pcoefs).shape[-1]
    n, m = noll2degrees(np.arange(nmodes) + 1)
    args = (model.wl, model.na, model.ni, model.res, model.size, model.vec_corr, model.condition, model.precision)
    h = _engine.hanser_handle(*args, -2)
    tensor = _engine.aberration_pupil(h, n, m, mcoefs, pcoefs)
    return _CompactPupil(tensor, args[:5] + (args[7],))


def psf_library(model, pcoefs, mcoefs=None, out=None):
    """Batched Zernike-aberrated PSF library (BASELINE config 5): one PSFi stack per coefficient row.

    pcoefs / mcoefs : (B, n) Noll-ordered phase / magnitude coefficients (either may be None).
    Returns a CUDA tensor (B, nz, N, N); equivalent to
    ``stack([apply_aberration(model, m, p).PSFi for m, p in zip(mcoefs, pcoefs)])``.
    """
    assert isinstance(model, HanserPSF), "Model must be a HanserPSF"
    pc = None if pcoefs is None else np.atleast_2d(np.asarray(pcoefs, float))
    mc = None if mcoefs is None else np.atleast_2d(np.asarray(mcoefs, float))
    if pc is None and mc is None:
        raise ValueError("need at least one of pcoefs / mcoefs")
    pupils = _aberrated_pupils(model, mc, pc)
    h = pupils.handle(model)
    z = _engine.to_device_f64(model.zrange)
    _, psfi = h.psf(z, pupils.tensor, want_psfa=False, want_psfi=True, psfi_out=out)
    return psfi


def apply_named_aberration(model, aberration, magnitude):
    """Apply one named phase aberration (pyotf/otf.py:648-651)."""
    return apply_aberration(model, None, named_aberration_to_pcoefs(aberration, magnitude))


def named_aberration_to_pcoefs(aberration, magnitude):
    """Noll-ordered phase-coefficient vector with one named entry set (pyotf/otf.py:654-679)."""
    try:
        noll = name2noll[aberration]
    except KeyError:
        raise KeyError(
            f"Aberration '{aberration}' unknown, choose from: '" + "', '".join(name2noll.keys()) + "'"
        ) from None
    pcoefs = np.zeros(max(name2noll.values()))
    pcoefs[noll - 1] = magnitude
    return pcoefs


def apply_named_aberrations(model, aberrations):
    """Apply a dict of name -> magnitude phase aberrations (pyotf/otf.py:682-702)."""
    pcoefs = np.zeros(max(name2noll.values()))
    for aberration, magnitude in aberrations.items():
        pcoefs = pcoefs + named_aberration_to_pcoefs(aberration, magnitude)
    return apply_aberration(model, None, pcoefs)


def otfi_zsharded(model, nz_total, group=None, mode="auto"):
    """OTFi of a stack whose planes are sharded over the ranks of a torch.distributed group (one process per GPU):
    `model` holds this rank's contiguous block of planes (its ``zrange`` slice -- the reference's own shard API,
    pyotf/otf.py:321-333) and the result is this rank's z block of ``fftshift(fftn(ifftshift(PSFi)))`` over the
    WHOLE stack (pyotf/otf.py:209-212, pyotf/utils.py:15-17) as a CUDA tensor.  ``mode``: "gather" (all-gather the
    real PSFi, transform locally), "alltoall" (2D transforms locally, all-to-all transpose, z transform, transpose
    back) or "auto"; see pyotf_b200/_dist.py."""
    from . import _dist

    return _dist.otfi_zsharded(model.PSFi_dev, int(nz_total), group=group, mode=mode)
