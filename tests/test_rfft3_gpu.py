"""Real 3D shifted FFT (OTFi = easy_fft(PSFi), pyotf/otf.py:209-212, pyotf/utils.py:15-17) through the Hermitian
half-spectrum passes of csrc/fftnd_impl.cuh (rfft3_shifted_run) against numpy's full transform."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(64, 64, 64), (128, 256, 256), (256, 128, 64), (64, 512, 128), (128, 64, 1024)])
@pytest.mark.parametrize("dtype,tol", [(np.float32, 2e-6), (np.float64, 1e-12)])
def test_real_volume_vs_numpy(shape, dtype, tol):
    from numpy.fft import fftn, fftshift, ifftshift
    from pyotf_b200.utils import easy_fft

    rng = np.random.default_rng(sum(shape))
    a = rng.standard_normal(shape).astype(dtype)
    a[shape[0] // 2, shape[1] // 2, shape[2] // 2] += 50.0        # a PSF-like peak at the centre
    ref = fftshift(fftn(ifftshift(a.astype(np.float64))))
    out = easy_fft(a)
    assert out.shape == ref.shape
    assert rel_l2(out, ref) <= tol
    # every plane on its own (the mirrored half is written by a different code path than the computed half)
    for zs in (0, 1, shape[0] // 2 - 1, shape[0] // 2, shape[0] // 2 + 1, shape[0] - 1):
        assert rel_l2(out[zs], ref[zs]) <= 4 * tol, zs
    # Hermitian symmetry of the result about the centre, exact by construction in the mirrored half
    c = out[1:, 1:, 1:]
    np.testing.assert_array_equal(c[: shape[0] // 2 - 1], np.conj(c[::-1, ::-1, ::-1][: shape[0] // 2 - 1]))


def test_matches_generic_passes_on_cfg1():
    """The cfg1 OTFi through the half-spectrum passes equals the oracle (and so the generic three-pass path)."""
    from oracle import pyotf_oracle as orc
    from pyotf_b200.otf import HanserPSF

    kw = dict(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=256, zsize=128)
    for prec, tol in (("fp32", 1e-5), ("fp64", 1e-11)):
        m = HanserPSF(precision=prec, **kw)
        psfi = orc.psfi(orc.hanser_psfa(kw["wl"], kw["na"], kw["ni"], kw["res"], 256, m.zrange, "none", "sine"))
        assert rel_l2(m.OTFi, orc.otfi(psfi)) <= tol


def test_fft3d_shifted_entry_point():
    """pyotf_b200_fft3d_shifted (the 3D form SURVEY.md section 8b lists): easy_fft / easy_ifft over all three axes."""
    import torch

    from pyotf_b200 import _engine, _lib

    rng = np.random.default_rng(5)
    a = rng.standard_normal((16, 32, 64)) + 1j * rng.standard_normal((16, 32, 64))
    x = torch.as_tensor(a, device="cuda")
    out = torch.empty_like(x)
    lib = _lib.load()
    for direction, ref in ((-1, np.fft.fftshift(np.fft.fftn(np.fft.ifftshift(a)))), (1, np.fft.ifftshift(np.fft.ifftn(np.fft.fftshift(a))))):
        _lib.check(lib.pyotf_b200_fft3d_shifted(_engine.ptr(x), _engine.ptr(out), 16, 32, 64, direction, 0, 1, _engine.stream_ptr()))
        assert rel_l2(out.cpu().numpy(), ref) <= 1e-12
    assert lib.pyotf_b200_version() == lib.pyotf_b200_abi_version()
