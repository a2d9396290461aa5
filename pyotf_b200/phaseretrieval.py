"""Phase retrieval with pyotf's API, iterated by the fused sm_100a kernel K2.

Drop-in for ``pyotf.phaseretrieval`` (reference: pyotf/phaseretrieval.py): ``retrieve_phase``
returns a ``PhaseRetrievalResult`` with the same fields (``mag``, ``phase``, ``mse``,
``pupil_diff``, ``mse_diff``, ``model``) and helpers (``fit_to_zernikes``, ``generate_psf``,
``complex_pupil``, ...).  The whole loop of the reference (pyotf/phaseretrieval.py:95-135) --
model PSF, mean squared error, magnitude replacement, back-transform, z-mean, convergence test
-- runs on the GPU behind ``pyotf_b200_pr_run`` (include/pyotf_b200.h); the host only post-
processes the final N x N pupil (phase unwrapping, Zernike least squares), as the reference does.
"""
import copy
import logging

import numpy as np
from numpy.fft import fftshift, ifftshift
from numpy.linalg import lstsq

from . import _engine
from .otf import HanserPSF
from .utils import fft_pad, psqrt  # noqa: F401
from .zernike import noll2degrees, noll2name, zernike  # noqa: F401

logger = logging.getLogger(__name__)


def unwrap_phase(image):
    """2D phase unwrapping by reliability-sorted path following (Herraez et al. 2002).

    Stand-in for ``skimage.restoration.unwrap_phase`` (pyotf/phaseretrieval.py:22,146), which is
    not vendored by the reference; runs in the library's host code (the algorithm is a serial
    union-find over sorted edges and sits after the hot path).
    """
    from . import _lib
    import ctypes as C

    a = np.ascontiguousarray(image, dtype=np.float64)
    if a.ndim != 2:
        raise ValueError("unwrap_phase: image must be 2D")
    out = np.empty_like(a)
    _lib.check(_lib.load().pyotf_b200_unwrap_phase_2d(a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                                      a.shape[0], a.shape[1]))
    return out


def retrieve_phase(data, params, max_iters=200, pupil_tol=1e-8, mse_tol=1e-8, phase_only=False, *,
                   precision=None, comm=None, nz_total=None, rows_per_cta=0):
    """Retrieve the pupil function from a measured PSF stack (Hanser et al. 2003).

    Same signature and semantics as pyotf/phaseretrieval.py:32-149.

    data : (nz, N, N) real array (numpy or CUDA tensor), background-subtracted, square in x/y
    params : dict for ``HanserPSF`` (``size``/``zsize``/``vec_corr``/``condition`` are overwritten,
        the dict is updated in place like the reference does)
    max_iters, pupil_tol, mse_tol, phase_only : as in the reference
    precision : "fp32" | "fp64" kernel arithmetic (keyword-only extension)
    rows_per_cta : tuning knob of the fused kernel (0 = auto; 1 forces the four-pass path K2b)
    comm, nz_total : z-sharded multi-GPU retrieval -- ``data`` holds this rank's planes,
        ``params["zrange"]`` their positions, ``nz_total`` the plane count over all ranks and
        ``comm`` a ``pyotf_b200._engine.Comm`` (one all-reduce per iteration)
    """
    assert max_iters > 0, "Must have at least one iteration"
    pr = PhaseRetrieval(data, params, precision=precision, comm=comm, nz_total=nz_total, rows_per_cta=rows_per_cta)
    return pr.run(max_iters, pupil_tol, mse_tol, phase_only)


def _stopped(mse, mse_diff, pupil_diff, pupil_tol, mse_tol):
    with np.errstate(invalid="ignore"):
        return bool(pupil_diff[-1] < pupil_tol or mse_diff[-1] < mse_tol or mse[-1] < mse_tol)


class PhaseRetrieval:
    """Object form of ``retrieve_phase`` that keeps the device state (measured stack, defocus tables, pupil
    estimate) between calls.  Not part of the reference, which only has the function.

    ``PhaseRetrieval(data, params).run(100, 1e-4, 1e-4)`` is ``retrieve_phase(data, params, 100, 1e-4, 1e-4)``.
    ``set_pupil(pupil)`` re-seeds the estimate (``None``: the ideal pupil the reference starts from,
    pyotf/phaseretrieval.py:89); ``step(n)`` runs n more iterations of pyotf/phaseretrieval.py:95-130 from the
    current estimate and returns their ``(mse, mse_diff, pupil_diff)``; ``pupil`` is the current estimate.  ``run``
    always starts from the ideal pupil, like the function.
    """

    def __init__(self, data, params, *, precision=None, comm=None, nz_total=None, rows_per_cta=0):
        assert data.shape[1] == data.shape[2], "Data is not square in x/y"
        assert data.ndim == 3, "Data doesn't have enough dims"
        params.update(dict(vec_corr="none", condition="none", zsize=data.shape[0], size=data.shape[-1]))
        self.params, self.comm = params, comm
        self.model = model = HanserPSF(**params, precision=precision)
        model._check_size()   # 32..256 powers of two: fused kernel K2; every other size: the four-pass path K2b
        torch = _engine._torch()
        want = torch.float32 if model.precision == "fp32" else torch.float64
        if not isinstance(data, np.ndarray) and hasattr(data, "is_cuda"):
            data_dev = data.to(device="cuda", dtype=want)
        else:
            a = np.asarray(data)
            if a.dtype not in (np.float32, np.float64):
                a = a.astype(np.float64)
            data_dev = torch.as_tensor(np.ascontiguousarray(a), device="cuda")
        self.state = _engine.PRState(model.wl, model.na, model.ni, model.res, model.size, model.zrange, data_dev,
                                     model.precision, nz_total=nz_total, rows_per_cta=rows_per_cta)

    def set_pupil(self, pupil=None):
        """Re-seed the estimate from an unshifted N x N complex pupil (numpy or CUDA tensor); None = ideal pupil."""
        if pupil is not None and not hasattr(pupil, "is_cuda"):
            pupil = _engine.to_device_c128(pupil)
        self.state.set_pupil(pupil)

    @property
    def pupil(self):
        """Current pupil estimate, unshifted N x N complex128 (host array)."""
        return _engine.to_host(self.state.get_pupil(applied=False))

    def step(self, n=1, phase_only=False):
        """n more iterations from the current estimate, no convergence test; returns (mse, mse_diff, pupil_diff)."""
        return self.state.run(int(n), 0.0, 0.0, phase_only, self.comm)

    def run(self, max_iters=200, pupil_tol=1e-8, mse_tol=1e-8, phase_only=False):
        assert max_iters > 0, "Must have at least one iteration"
        state, model = self.state, self.model
        state.set_pupil(None)
        mse, mse_diff, pupil_diff = state.run(max_iters, pupil_tol, mse_tol, phase_only, self.comm)
        if len(mse) == max_iters and not _stopped(mse, mse_diff, pupil_diff, pupil_tol, mse_tol):
            logger.info("Reach max iterations without convergence")
        logger.info(
            f"Finished with {len(mse) - 1} iterations and rmse={np.sqrt(mse[-1]):.2g}, "
            + f"mse_diff={mse_diff[-1]:.2g}, pupil_diff={pupil_diff[-1]:.2g}"
        )
        new_pupil = _engine.to_host(state.get_pupil(applied=False))
        # the model keeps the pupil it last applied (phaseretrieval.py:97), device resident
        model.apply_pupil(state.get_pupil(applied=True))
        model._gen_kr()
        mask = fftshift((model._kr <= model.na / model.wl).astype(float))
        phase = unwrap_phase(fftshift(np.angle(new_pupil))) * mask
        magnitude = fftshift(abs(new_pupil)) * mask
        result = PhaseRetrievalResult(magnitude, phase, mse, pupil_diff, mse_diff, model)
        result.pupil = new_pupil  # unshifted complex pupil (extension: no unwrap round trip)
        return result


class PhaseRetrievalResult(object):
    """Holds a retrieved pupil: ``mag``/``phase`` are N x N, fftshift-ed (pyotf/phaseretrieval.py:152-273)."""

    def __init__(self, mag, phase, mse, pupil_diff, mse_diff, model):
        self.mag = mag
        self.phase = phase
        self.mse = mse
        self.pupil_diff = pupil_diff
        self.mse_diff = mse_diff
        self.model = model
        model._gen_kr()
        r, theta = model._kr, model._phi
        self.r, self.theta = fftshift(r), fftshift(theta)
        self.na, self.wl = model.na, model.wl

    @property
    def rms_error(self):
        """RMS wavefront error."""
        return np.sqrt((self.phase[self.r <= self.na / self.wl] ** 2).mean())

    @property
    def pv_error(self):
        """PV wavefront error."""
        wavefront_flat = self.phase[self.r <= self.na / self.wl]
        return wavefront_flat.max() - wavefront_flat.min()

    def fit_to_zernikes(self, num_zerns, *, mapping=noll2degrees):
        """Least-squares fit of mag and phase to the first ``num_zerns`` modes (pyotf/phaseretrieval.py:201-211).

        The modes are evaluated on the in-pupil pixels and the normal equations are accumulated on the GPU
        (``pyotf_b200_zernike`` + ``pyotf_b200_zernike_gram``, fp64); only the num_zerns x num_zerns solve runs on
        the host.  ``zd_result.zerns`` (the full-grid modes the reference stores eagerly) is evaluated on first use.
        """
        r, theta = self.r, self.theta
        r = r / (self.na / self.wl)
        n, m = mapping(np.arange(1, num_zerns + 1))
        valid = r <= 1
        coefs = _engine.zernike_lstsq(r[valid], theta[valid], n, m, np.stack([self.mag[valid], self.phase[valid]]))
        self.zd_result = ZernikeDecomposition(coefs[0], coefs[1], _LazyZerns(r, theta, n, m))
        return self.zd_result

    def _generate_psf(self, complex_pupil, size=None, zsize=None, zrange=None):
        """PSFi of ``complex_pupil`` (shifted) on a copy of the model (pyotf/phaseretrieval.py:213-239)."""
        model = copy.copy(self.model)
        if zsize is not None:
            model.zsize = zsize
        if zrange is not None:
            model.zrange = zrange
        model.apply_pupil(ifftshift(complex_pupil))
        psf = model.PSFi
        nz, ny, nx = psf.shape
        assert ny == nx, "Something is very wrong"
        if size is not None:
            if nx < size:
                psf = fft_pad(psf, (nz, size, size), mode="constant")
            elif nx > size:
                lb = size // 2
                hb = size - lb
                myslice = slice(nx // 2 - lb, nx // 2 + hb)
                psf = psf[:, myslice, myslice]
        return psf

    def generate_zd_psf(self, sphase=slice(4, None, None), size=None, zsize=None, zrange=None):
        """PSF from the Zernike decomposition (if available)."""
        return self._generate_psf(self.zd_result.complex_pupil(sphase=sphase), size, zsize, zrange)

    def generate_psf(self, size=None, zsize=None, zrange=None):
        """PSF from the retrieved pupil."""
        return self._generate_psf(self.complex_pupil, size, zsize, zrange)

    def plot(self, axs=None):
        raise NotImplementedError("plotting is out of scope for pyotf_b200 (SURVEY.md section 2, row 6)")

    def plot_convergence(self):
        raise NotImplementedError("plotting is out of scope for pyotf_b200 (SURVEY.md section 2, row 6)")

    @property
    def complex_pupil(self):
        """Complex pupil function, shifted: ``mag * exp(i phase)``."""
        return self.mag * np.exp(1j * self.phase)


class ZernikeDecomposition(object):
    """Zernike decomposition of a pupil's magnitude and phase (pyotf/phaseretrieval.py:276-370)."""

    def __init__(self, mag_coefs, phase_coefs, zerns):
        assert mag_coefs.size == phase_coefs.size == zerns.shape[0]
        self.mcoefs = mag_coefs
        self.pcoefs = phase_coefs
        self._zerns = zerns

    @property
    def zerns(self):
        """(m, N, N) Zernike modes used in the decomposition (evaluated on first access)."""
        if isinstance(self._zerns, _LazyZerns):
            self._zerns = self._zerns.evaluate()
        return self._zerns

    @zerns.setter
    def zerns(self, value):
        self._zerns = value

    @property
    def rms_error(self):
        """RMS wavefront error."""
        return np.sqrt((self.pcoefs**2).sum())

    def _recon(self, coefs, s):
        return _recon_from_zerns(coefs[s], self.zerns[s])

    def phase(self, s):
        """Phase reconstructed from the modes in slice ``s``."""
        return self._recon(self.pcoefs, s)

    def mag(self, s):
        """Magnitude reconstructed from the modes in slice ``s``."""
        return self._recon(self.mcoefs, s)

    def complex_pupil(self, smag=slice(None), sphase=slice(None)):
        """Complex pupil reconstructed from the chosen mode slices."""
        return self.mag(smag) * np.exp(1j * self.phase(sphase))

    def plot_named_coefs(self):
        raise NotImplementedError("plotting is out of scope for pyotf_b200")

    def plot_coefs(self):
        raise NotImplementedError("plotting is out of scope for pyotf_b200")

    def plot(self, smag=slice(None), sphase=slice(None), axs=None):
        raise NotImplementedError("plotting is out of scope for pyotf_b200")


class _LazyZerns:
    """Full-grid Zernike modes, evaluated (on the GPU) only when somebody asks for them."""

    def __init__(self, r, theta, n, m):
        self.r, self.theta, self.n, self.m = r, theta, n, m
        self.shape = (len(n),) + r.shape

    def evaluate(self):
        return zernike(self.r, self.theta, self.n, self.m)


def _calc_mse(data1, data2):
    """Mean squared difference (pyotf/phaseretrieval.py:401-403)."""
    return ((data1 - data2) ** 2).mean()


def _fit_to_zerns(data, zerns, r, **kwargs):
    """Least squares over the in-pupil pixels r <= 1 (pyotf/phaseretrieval.py:406-432)."""
    valid_points = r <= 1
    data2fit = data[valid_points]
    zerns2fit = zerns[:, valid_points].T
    coefs, _, _, _ = lstsq(zerns2fit, data2fit, rcond=None, **kwargs)
    return coefs


def _recon_from_zerns(coefs, zerns):
    """sum_j coefs[j] * zerns[j] (pyotf/phaseretrieval.py:435-437)."""
    return (coefs[:, np.newaxis, np.newaxis] * zerns).sum(0)
