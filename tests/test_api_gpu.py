"""Drop-in API parity (GPU): the reference's own tests restated against pyotf_b200, plus
value parity against the oracle / golden vectors.  Mirrors tests/otf_test.py,
tests/zernike_test.py and tests/utils_test.py of the reference."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu


class _BasePSFCase:
    def test_dtypes(self):
        m = self.model
        assert np.issubdtype(m.PSFi.dtype, np.floating)
        assert np.issubdtype(m.PSFa.dtype, np.complexfloating)
        assert np.issubdtype(m.OTFi.dtype, np.complexfloating)
        assert np.issubdtype(m.OTFa.dtype, np.complexfloating)

    def test_PSFi_positive(self):
        assert (self.model.PSFi >= 0).all()

    def test_diffraction_limit(self):
        with pytest.raises(ValueError):
            self.model.res = self.model.wl / self.model.na


class TestHanserPSF(_BasePSFCase):
    @pytest.fixture(autouse=True)
    def _model(self):
        from pyotf_b200.otf import HanserPSF

        self.model = HanserPSF(525, 0.85, 1.0, 140, 64)

    def test_size(self):
        m = self.model
        m.size = 128
        m.zrange = [-1000, 0, 1000]
        assert m.PSFi.shape == (3, 128, 128)
        m.size = 256
        m.zrange = 0
        assert m.PSFi.shape == (1, 256, 256)


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_hanser_results_vs_oracle(precision):
    from oracle import pyotf_oracle as orc
    from pyotf_b200.otf import HanserPSF

    tol = {"fp32": 1e-5, "fp64": 1e-11}[precision]
    kw = dict(wl=525, na=1.27, ni=1.33, res=100, size=64, zres=250, zsize=5)
    for vc, cond in (("none", "sine"), ("total", "herschel"), ("y", "none")):
        m = HanserPSF(vec_corr=vc, condition=cond, precision=precision, **kw)
        a = orc.hanser_psfa(kw["wl"], kw["na"], kw["ni"], kw["res"], kw["size"], m.zrange, vc, cond)
        assert rel_l2(m.PSFi, orc.psfi(a)) <= tol
        assert rel_l2(m.PSFa, a) <= tol
        assert rel_l2(m.OTFi, orc.otfi(orc.psfi(a))) <= tol
        assert rel_l2(m.OTFa, orc.hanser_otfa(a)) <= tol
        # PSFi computed from an existing PSFa (reduction path) agrees with the fused path
        m2 = HanserPSF(vec_corr=vc, condition=cond, precision=precision, **kw)
        _ = m2.PSFa
        assert rel_l2(m2.PSFi, m.PSFi) <= tol


def test_hanser_golden_otf():
    from pyotf_b200.otf import HanserPSF

    g = np.load(os.path.join(GOLDEN, "hanser_small.npz"))
    wl, na, ni, res, size, zres, zsize = g["params"]
    m = HanserPSF(wl, na, ni, res, int(size), zres, int(zsize), precision="fp64")
    assert rel_l2(m.OTFi, g["otfi_none_sine"]) <= 1e-11
    assert rel_l2(m.OTFa, g["otfa_none_sine"]) <= 1e-11
    assert rel_l2(m.PSFa, g["psfa_none_sine"]) <= 1e-11


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_cfg1_otfi_golden(precision):
    """BASELINE config 1 -> OTFi (3D shifted FFT of the 128 x 256 x 256 PSFi)."""
    from pyotf_b200.otf import HanserPSF

    tol = {"fp32": 1e-5, "fp64": 1e-11}[precision]
    g = np.load(os.path.join(GOLDEN, "cfg1_hanser256.npz"))
    m = HanserPSF(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=256, zsize=128, precision=precision)
    o = m.OTFi
    assert o.shape == (128, 256, 256)
    assert rel_l2(o[g["otfi_plane_idx"]], g["otfi_planes"]) <= tol
    assert rel_l2(o[:, 128, :], g["otfi_zx"]) <= tol
    assert abs(np.linalg.norm(o.ravel()) / g["otfi_l2"] - 1) <= tol
    assert abs(o[64, 128, 128] / 23.75643331623897 - 1) <= tol
    assert abs(o[64, 128, 140].real / 0.9735855131724713 - 1) <= 10 * tol


def test_hooks_and_cache_semantics():
    from oracle import pyotf_oracle as orc
    from pyotf_b200.otf import HanserPSF

    m = HanserPSF(520, 0.85, 1.0, 130, 64, 300, 3, precision="fp64")
    m._gen_kr()
    k, kr, phi, kmag, kz = orc.hanser_kspace(520, 1.0, 130, 64)
    np.testing.assert_array_equal(m._k, k)
    np.testing.assert_array_equal(m._kr, kr)          # bit-equal to numpy.hypot
    np.testing.assert_array_equal(m._kz, kz)
    np.testing.assert_allclose(m._phi, phi, rtol=1e-15, atol=1e-16)
    assert m._kmag == kmag
    np.testing.assert_array_equal(m._gen_pupil(), orc.hanser_pupil(kr, 0.85, 520))
    assert rel_l2(m._calc_defocus(), orc.hanser_defocus(kz, m.zrange)) <= 1e-13
    p0 = m.PSFi
    assert m.PSFi is p0  # cached
    m.zres = 200
    assert m.PSFi is not p0  # invalidated
    np.testing.assert_array_equal(m.zrange, orc.default_zrange(3, 200))
    # apply_pupil then an attribute change forgets the pupil (the reference recomputes the ideal PSF)
    rng = np.random.default_rng(0)
    pupil = m._gen_pupil() * np.exp(1j * rng.standard_normal((64, 64)))
    m.apply_pupil(pupil)
    a = orc.hanser_psfa(520, 0.85, 1.0, 130, 64, m.zrange, pupil_base=pupil)
    assert rel_l2(m.PSFa, a) <= 1e-11
    assert rel_l2(m._gen_psf(pupil), a) <= 1e-11
    m.zrange = [0.0, 100.0]
    assert rel_l2(m.PSFi, orc.psfi(orc.hanser_psfa(520, 0.85, 1.0, 130, 64, [0.0, 100.0]))) <= 1e-11
    with pytest.raises(AssertionError):
        m._gen_psf(np.ones((2, 64, 64), complex))
    for bad in (dict(vec_corr="q"), dict(condition="q")):
        with pytest.raises(ValueError):
            HanserPSF(520, 0.85, 1.0, 130, 64, **bad)
    with pytest.raises(TypeError):
        HanserPSF(520, 0.85, 1.0, 130, 64.0)
    with pytest.raises(ValueError):
        HanserPSF(520, 0.85, 1.0, 130, 64, zres=-1)
    assert "zrange=" in repr(m) and repr(m).startswith("HanserPSF(wl=520")


def test_named_aberration_to_pcoefs():
    from pyotf_b200.otf import named_aberration_to_pcoefs

    expected = np.zeros(56)
    expected[1] = 1
    np.testing.assert_array_equal(named_aberration_to_pcoefs("tip", 1), expected)
    with pytest.raises(KeyError):
        named_aberration_to_pcoefs("nope", 1)


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_apply_aberration_golden(precision):
    from pyotf_b200.otf import HanserPSF, apply_aberration, apply_named_aberration

    tol = {"fp32": 1e-5, "fp64": 1e-11}[precision]
    g = np.load(os.path.join(GOLDEN, "aberration_small.npz"))
    b = HanserPSF(wl=520, na=0.85, ni=1.0, res=130, size=64, zres=300, zsize=4, precision=precision)
    assert rel_l2(apply_aberration(b, g["mcoefs"], g["pcoefs"]).PSFi, g["psfi_both"]) <= tol
    assert rel_l2(apply_aberration(b, None, g["pcoefs"]).PSFi, g["psfi_phase"]) <= tol
    assert rel_l2(apply_aberration(b, g["mcoefs"], None).PSFi, g["psfi_mag"]) <= tol
    assert apply_aberration(b, None, None) is not b
    with pytest.raises(AssertionError):
        apply_aberration(b, np.zeros(3), np.zeros(4))
    assert apply_named_aberration(b, "defocus", 0.5).PSFi.shape == (4, 64, 64)


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_cfg5_library_golden(precision):
    """BASELINE config 5 unit + 8 sampled entries of the 8192-entry library (256^2 x 64)."""
    from pyotf_b200.otf import HanserPSF, apply_aberration, psf_library

    tol = {"fp32": 1e-5, "fp64": 1e-11}[precision]
    g = np.load(os.path.join(GOLDEN, "cfg5_library.npz"))
    base = HanserPSF(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=256, zsize=64, precision=precision)
    one = np.random.default_rng(0).standard_normal(120) * 0.1
    p = apply_aberration(base, None, one).PSFi
    assert rel_l2(p[32], g["unit_psfi_plane32"]) <= tol
    assert rel_l2(p.sum((1, 2)), g["unit_psfi_plane_sums"]) <= tol
    assert abs(p[32, 128, 128] / 0.009643743189627898 - 1) <= 10 * tol
    coefs = np.random.default_rng(0).standard_normal((8192, 120)) * 0.1
    idx = g["sample_idx"][:8]
    lib = psf_library(base, coefs[idx]).cpu().numpy()
    assert lib.shape == (8, 64, 256, 256)
    assert rel_l2(lib.sum((2, 3)), g["entry_plane_sums"]) <= tol
    assert rel_l2(lib[:, :, 120:136, 120:136], g["entry_center_crops"]) <= tol
    assert rel_l2(lib[:2, 32], g["entry_plane32"]) <= tol


def test_zernike_golden_and_reference_pins():
    from pyotf_b200.zernike import (cart2pol, degrees2fringe, degrees2name, degrees2noll, degrees2osa,
                                    fringe2degrees, fringe2name, noll2degrees, noll2name, osa2degrees,
                                    osa2name, zernike)

    g = np.load(os.path.join(GOLDEN, "zernike_noll120.npz"))
    z = zernike(g["r"], g["theta"], g["n"], g["m"])
    np.testing.assert_allclose(z, g["zern"], rtol=0, atol=5e-12)
    np.testing.assert_allclose(zernike(g["r"], g["theta"], g["n"][:36], g["m"][:36], norm=False),
                               g["zern_unnormed"], rtol=0, atol=5e-12)
    # index maps (Wikipedia tables, tests/zernike_test.py:66-113)
    deg = np.array(((0, 0), (1, 1), (1, -1), (2, 0), (2, -2), (2, 2), (3, -1), (3, 1), (3, -3), (3, 3)))
    n, m = deg.T
    assert (degrees2noll(n, m) == np.arange(1, 11)).all()
    nn, mm = noll2degrees(np.arange(1, 11))
    assert (nn == n).all() and (mm == m).all()
    assert (degrees2osa(n, m) == np.array((0, 2, 1, 4, 3, 5, 7, 8, 6, 9))).all()
    assert (degrees2fringe(n, m) == np.array((1, 2, 3, 4, 6, 5, 8, 7, 11, 10))).all()
    j = np.random.default_rng(0).integers(1, 36, 10)
    assert (degrees2noll(*noll2degrees(j)) == j).all()
    assert (degrees2osa(*osa2degrees(j)) == j).all()
    for bad in (0, -1):
        with pytest.raises(ValueError):
            noll2degrees(bad)
    for fn, args in ((noll2degrees, (2.5,)), (noll2degrees, (1.0,)), (osa2degrees, (2.5,)),
                     (degrees2noll, (1.0, 3.0)), (degrees2noll, (1.5, 3.5)), (degrees2osa, (1.0, 3.0)),
                     (degrees2noll, (1, 2)), (degrees2osa, (1, 2))):
        with pytest.raises(ValueError):
            fn(*args)
    with pytest.raises(NotImplementedError):
        fringe2degrees(3)
    for mapping in (noll2name, osa2name, fringe2name):
        assert (np.diff(list(mapping.keys())) > 0).all()
    # input validation (tests/zernike_test.py:60-63, 116-148)
    with pytest.raises(ValueError):
        zernike(0.5, 0.0, 4, 5)
    for args in ((0, 0, np.ones((2, 2, 2)), np.ones((2, 2, 2))), (-1, 0, 0, 1), (np.ones((10, 10, 2)), 0, 0, 1),
                 (np.ones((3, 3, 3)), np.ones((3, 3, 3)), 0, 0)):
        with pytest.raises(ValueError):
            zernike(*args)
    x = np.linspace(-1, 1, 512)
    xx, yy = np.meshgrid(x, x)
    r, theta = cart2pol(yy, xx)
    assert zernike(r, theta, 0, 0).shape == r.shape
    assert zernike(1.0, 0.3, 4, 2) == zernike(1, 0.3, 4, 2)
    assert np.isfinite(zernike(0.5, 0.7, 97, -33)).all()
    rr = np.random.default_rng(1).random(100) * 2
    assert (zernike(rr, rr, 5, 2) == 0).all()
    # unit RMS normalisation (tests/zernike_test.py:201-213), 1024^2 grid
    x = np.linspace(-1, 1, 1024)
    xx, yy = np.meshgrid(x, x)
    r, theta = cart2pol(yy, xx)
    for (n_, m_), name in sorted(degrees2name.items()):
        zn = zernike(r, theta, n_, m_, norm=True)
        tol = max(10.0 ** (n_ - 6), 5e-3)
        np.testing.assert_allclose(1.0, np.sqrt((zn[r <= 1] ** 2).mean()), err_msg=name, atol=tol, rtol=tol)


def test_utils():
    from pyotf_b200.utils import cart2pol, center_data, easy_fft, easy_ifft, prep_data_for_PR, psqrt, remove_bg

    rng = np.random.default_rng(2)
    d = np.array((1, 2, 3, 3, 3, 4, 5), dtype=np.uint16)
    assert np.allclose(remove_bg(d, 1.0), d - 3.0)
    with pytest.raises(ValueError):
        remove_bg(d.astype(float), 1.0)
    data = np.zeros((37, 12))
    data[5, 7] = 1
    assert np.fft.ifftshift(center_data(data))[0, 0] == 1
    x = rng.integers(-1000, 1000, 20)
    assert (psqrt(x)[x < 0] == 0).all() and np.allclose(psqrt(x)[x >= 0], np.sqrt(x[x >= 0]))
    z = rng.standard_normal(10) + 1j * rng.standard_normal(10)
    rho, th = cart2pol(z.imag, z.real)
    assert np.allclose(rho, abs(z)) and np.allclose(th, np.angle(z))
    for shape_in, xysize, shape_out in (((5, 5), None, (5, 5)), ((5, 4), None, (5, 5)), ((5, 4), 10, (10, 10)),
                                        ((5, 4), 3, (3, 3))):
        a = np.ones((10,) + shape_in, dtype=int)
        a[0, 0, 0], a[-1, -1, -1] = 2, 0
        out = prep_data_for_PR(a, xysize=xysize)
        assert out.shape == (10,) + shape_out and (out >= 0).all()
    for size in (3, 4, 5, 6, 7):
        a = np.ones((10, 8, 8), dtype=int)
        a[0, 0, 0], a[-1, -1, -1] = 2, 0
        out = prep_data_for_PR(a, xysize=size)
        assert np.unravel_index(out.argmax(), out.shape) == (5, size // 2, size // 2)
    # shifted FFTs incl. odd / non power-of-two lengths and partial axes
    from numpy.fft import fftn, fftshift, ifftn, ifftshift

    for shape, axes in (((8, 16, 32), None), ((5, 12, 7), None), ((3, 25, 64, 64), (1, 2, 3)), ((6, 10), (1,)),
                        ((2, 512), (1,)), ((1024, 3), (0,))):
        a = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
        assert rel_l2(easy_fft(a, axes), fftshift(fftn(ifftshift(a, axes=axes), axes=axes), axes=axes)) <= 1e-12
        assert rel_l2(easy_ifft(a, axes), ifftshift(ifftn(fftshift(a, axes=axes), axes=axes), axes=axes)) <= 1e-12
        assert rel_l2(easy_fft(a.real, axes), fftshift(fftn(ifftshift(a.real, axes=axes), axes=axes), axes=axes)) <= 1e-12
        a32 = a.astype(np.complex64)
        assert rel_l2(easy_fft(a32, axes), fftshift(fftn(ifftshift(a, axes=axes), axes=axes), axes=axes)) <= 2e-6


def test_library_and_stack_edge_cases():
    """Single-entry and several-hundred-entry libraries, one-plane and 1000-plane stacks, scalar zrange."""
    from oracle import pyotf_oracle as orc
    from pyotf_b200.otf import HanserPSF, psf_library

    kw = dict(wl=520, na=0.85, ni=1.0, res=130)
    base = HanserPSF(size=64, zres=300, zsize=3, precision="fp64", **kw)
    coefs = np.random.default_rng(5).standard_normal((300, 21)) * 0.2
    lib = psf_library(base, coefs).cpu().numpy()
    assert lib.shape == (300, 3, 64, 64)
    for b in (0, 137, 299):
        pupil = orc.aberrated_pupil(kw["wl"], kw["na"], kw["ni"], kw["res"], 64, None, coefs[b])
        ref = orc.psfi(orc.hanser_psfa(kw["wl"], kw["na"], kw["ni"], kw["res"], 64, base.zrange, pupil_base=pupil))
        assert rel_l2(lib[b], ref) <= 1e-11
    one = psf_library(base, coefs[7]).cpu().numpy()
    assert one.shape == (1, 3, 64, 64) and rel_l2(one[0], lib[7]) == 0.0
    with pytest.raises(ValueError):
        psf_library(base, None)
    long_stack = HanserPSF(size=32, zres=50, zsize=1000, precision="fp32", **kw)
    ref = orc.psfi(orc.hanser_psfa(kw["wl"], kw["na"], kw["ni"], kw["res"], 32, long_stack.zrange))
    assert long_stack.PSFi.shape == (1000, 32, 32) and rel_l2(long_stack.PSFi, ref) <= 1e-5
    single = HanserPSF(size=128, zrange=250.0, precision="fp64", **kw)
    assert rel_l2(single.PSFi, orc.psfi(orc.hanser_psfa(kw["wl"], kw["na"], kw["ni"], kw["res"], 128, [250.0]))) <= 1e-11
