"""Pin oracle/pyotf_oracle.py against the LIVE upstream pyotf (only where /root/reference exists).

These run in the build container; on the GPU box they skip and tests/test_oracle_golden.py
(committed vectors generated from the same live import) carries the pin.
"""
import numpy as np
import pytest

from oracle import pyotf_oracle as orc
from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.reference_available(), reason="no live reference")


@pytest.fixture(scope="module")
def ref():
    refshim.load_reference()
    import pyotf.otf as rotf
    import pyotf.phaseretrieval as rpr
    import pyotf.utils as rutils
    import pyotf.zernike as rzern

    class R:
        otf, pr, utils, zern = rotf, rpr, rutils, rzern

    return R


@pytest.mark.parametrize("vec_corr", ["none", "x", "y", "z", "total"])
@pytest.mark.parametrize("condition", ["none", "sine", "herschel"])
def test_hanser_psfa_all_modes(ref, vec_corr, condition):
    kw = dict(wl=525, na=1.27, ni=1.33, res=100, size=64, zres=250, zsize=5)
    m = ref.otf.HanserPSF(vec_corr=vec_corr, condition=condition, **kw)
    a = orc.hanser_psfa(
        kw["wl"], kw["na"], kw["ni"], kw["res"], kw["size"], m.zrange, vec_corr, condition
    )
    assert a.shape == m.PSFa.shape
    np.testing.assert_array_equal(a, m.PSFa)
    np.testing.assert_array_equal(orc.psfi(a), m.PSFi)
    np.testing.assert_array_equal(orc.otfi(orc.psfi(a)), m.OTFi)
    np.testing.assert_array_equal(orc.hanser_otfa(a), m.OTFa)


def test_hanser_zrange_and_pupil_base(ref):
    rng = np.random.default_rng(3)
    m = ref.otf.HanserPSF(520, 0.85, 1.0, 130, 32, zrange=[-500.0, 0.0, 730.0])
    m._gen_kr()
    pb = m._gen_pupil() * np.exp(1j * rng.standard_normal((32, 32)))
    m.apply_pupil(pb)
    a = orc.hanser_psfa(520, 0.85, 1.0, 130, 32, [-500.0, 0.0, 730.0], pupil_base=pb)
    np.testing.assert_array_equal(a, m.PSFa)
    np.testing.assert_array_equal(orc.default_zrange(7, 300), ref.otf.HanserPSF(520, 0.85, 1.0, 130, 32, 300, 7).zrange)


def test_zernike_noll_1_to_120(ref):
    j = np.arange(1, 121)
    n, m = ref.zern.noll2degrees(j)
    n2, m2 = orc.noll_to_degrees(j)
    np.testing.assert_array_equal(n, n2)
    np.testing.assert_array_equal(m, m2)
    x = np.linspace(-1.2, 1.2, 97)
    xx, yy = np.meshgrid(x, x)
    r, th = ref.utils.cart2pol(yy, xx)
    z_ref = ref.zern.zernike(r, th, n, m)
    z = orc.zernike_modes(r, th, n, m)
    # eval_jacobi (cephes hypergeometric/recurrence) vs the plain recurrence: few ulp
    np.testing.assert_allclose(z, z_ref, rtol=0, atol=2e-12)
    z_ref = ref.zern.zernike(r, th, n, m, norm=False)
    np.testing.assert_allclose(orc.zernike_modes(r, th, n, m, norm=False), z_ref, rtol=0, atol=1e-12)


def test_apply_aberration(ref):
    rng = np.random.default_rng(0)
    pcoefs = rng.standard_normal(37) * 0.2
    mcoefs = rng.standard_normal(37) * 0.05
    base = ref.otf.HanserPSF(520, 0.85, 1.0, 130, 64, 300, 4)
    for mc, pc in ((mcoefs, pcoefs), (None, pcoefs), (mcoefs, None)):
        m = ref.otf.apply_aberration(base, mc, pc)
        pupil = orc.aberrated_pupil(520, 0.85, 1.0, 130, 64, mc, pc)
        a = orc.hanser_psfa(520, 0.85, 1.0, 130, 64, base.zrange, pupil_base=pupil)
        np.testing.assert_allclose(orc.psfi(a), m.PSFi, rtol=1e-10, atol=1e-18)


@pytest.mark.parametrize("phase_only", [False, True])
def test_retrieve_phase_loop(ref, phase_only):
    rng = np.random.default_rng(5)
    kw = dict(wl=525, na=1.27, ni=1.33, res=100, size=64, zrange=[-1000, -500, 0, 250, 1000])
    model = ref.otf.HanserPSF(vec_corr="none", condition="none", **kw)
    model._gen_kr()
    pupil = model._gen_pupil() * np.exp(1j * 0.5 * rng.standard_normal((64, 64)))
    model.apply_pupil(pupil)
    data = model.PSFi + 1e-6 * rng.random(model.PSFi.shape)
    res = ref.pr.retrieve_phase(data, dict(kw), 25, 1e-7, 1e-7, phase_only)
    o = orc.retrieve_phase_loop(
        data, 525, 1.27, 1.33, 100, zrange=kw["zrange"], max_iters=25, pupil_tol=1e-7,
        mse_tol=1e-7, phase_only=phase_only,
    )
    assert o["iters"] == len(res.mse)
    np.testing.assert_array_equal(o["mse"], res.mse)
    np.testing.assert_array_equal(o["mse_diff"], res.mse_diff)
    np.testing.assert_array_equal(o["pupil_diff"], res.pupil_diff)
    np.testing.assert_array_equal(o["psfi"], res.model.PSFi)
    np.testing.assert_array_equal(np.fft.fftshift(abs(o["pupil"])) * np.fft.fftshift(o["mask"]), res.mag)


@pytest.mark.parametrize("vec_corr,condition,dual", [
    ("none", "sine", False), ("total", "sine", False), ("x", "herschel", True), ("z", "none", False),
])
def test_sheppard(ref, vec_corr, condition, dual):
    kw = dict(wl=520, na=1.27, ni=1.33, res=100, size=32, zres=190, zsize=24)
    m = ref.otf.SheppardPSF(vec_corr=vec_corr, condition=condition, dual=dual, **kw)
    o = orc.sheppard_otfa(vec_corr=vec_corr, condition=condition, dual=dual, **kw)
    np.testing.assert_array_equal(o, m.OTFa)
    np.testing.assert_array_equal(orc.sheppard_psfa(o), m.PSFa)
    np.testing.assert_array_equal(orc.psfi(orc.sheppard_psfa(o)), m.PSFi)


def test_sheppard_equal_res(ref):
    kw = dict(wl=500, na=0.85, ni=1.0, res=140, size=32, zres=140, zsize=32)
    m = ref.otf.SheppardPSF(**kw)
    np.testing.assert_array_equal(orc.sheppard_otfa(**kw), m.OTFa)


@pytest.mark.parametrize("size,oversample", [(16, 3), (15, 3), (12, 1), (10, 5)])
def test_microscope_widefield_and_confocal(size, oversample):
    """N4: the oracle's restatement of WidefieldMicroscope.PSF / ConfocalMicroscope.model_psf against the live
    pyotf.microscope (bin_ndarray = block sum, the un-vendored dphtools default, oracle/refshim.py)."""
    refshim.load_reference()
    from pyotf import microscope as rm

    kw = dict(model="hanser", na=0.85, ni=1.0, wl=520, size=size, pixel_size=130, oversample_factor=oversample)
    w = rm.WidefieldMicroscope(**kw)
    np.testing.assert_array_equal(orc.widefield_psf(w.model.PSFi, oversample, size * oversample), w.PSF)
    for pinhole in (1.0, 0.05):                       # with and without the pinhole convolution (radius <= 1.5 px)
        c = rm.ConfocalMicroscope(wl_exc=488.0, pinhole_size=pinhole, **kw)
        got = orc.confocal_model_psf(c.model.PSFi, c.model_exc.PSFi, 520, 0.85, c.model.res, pinhole)
        np.testing.assert_array_equal(got, c.model_psf)
        np.testing.assert_array_equal(orc.widefield_psf(got, oversample, size * oversample), c.PSF)
        np.testing.assert_array_equal(orc.otfi(c.PSF), c.OTF)


@pytest.mark.parametrize("kind", ["sim2d", "os-sim", "sim3d", "lattice", "sim3d-no-suppress", "sim2d-wiener0"])
def test_microscope_sim_family(kind):
    """N4: the oracle's restatement of BaseSIMMicroscope.model_psf (pyotf/microscope.py:316-401) against the live
    SIM2DMicroscope / SIM3DMicroscope / LatticeSIMMicroscope, with and without the Wiener filter."""
    refshim.load_reference()
    from pyotf import microscope as rm

    kw = dict(model="hanser", na=0.85, ni=1.0, wl=0.585, size=16, pixel_size=0.13, oversample_factor=1, vec_corr="none",
              condition="none", na_exc=None, wl_exc=0.561, wiener=0.1, dc_suppress=True)
    ori = (0, 2 * np.pi / 3, 4 * np.pi / 3)
    if kind == "sim2d":
        m = rm.SIM2DMicroscope(orientations=ori, **kw)
    elif kind == "os-sim":
        m = rm.SIM2DMicroscope(orientations=ori, **{**kw, "na_exc": 0.425})
    elif kind == "sim3d":
        m = rm.SIM3DMicroscope(orientations=ori, **kw)
    elif kind == "lattice":
        m = rm.LatticeSIMMicroscope(**kw)
    elif kind == "sim3d-no-suppress":
        m = rm.SIM3DMicroscope(orientations=ori[:2], **{**kw, "wiener": 0.5, "dc_suppress": False})
    else:
        m = rm.SIM2DMicroscope(orientations=ori, **{**kw, "wiener": 0})
    got = orc.sim_model_psf(m.model.PSFi, m.model.OTFi, m.psf_params["res"], m.psf_params["zres"], m.psf_params["ni"], m.wl_exc,
                            m.na_exc, m.wiener, m.coherent, m.dc, m.dc_suppress, m.orientations)
    np.testing.assert_array_equal(got, m.model_psf)
    np.testing.assert_array_equal(orc.widefield_psf(got, 1, 16), m.PSF)


def test_data_prep_restatements(ref):
    """N3: center_data / remove_bg / prep_data_for_PR (no-resize branch) of the oracle against the live
    pyotf.utils (pyotf/utils.py:76-147, 161-201) and its doctest values (utils.py:91-114, :134-136)."""
    rng = np.random.default_rng(2)
    for shape in ((3, 3), (3, 4), (5, 6, 7), (4, 9)):
        a = rng.integers(0, 50, shape)
        np.testing.assert_array_equal(orc.center_data(a), ref.utils.center_data(a))
    np.testing.assert_array_equal(orc.center_data(np.array(((2, 1, 0, 0), (1, 0, 0, 0), (0, 0, 0, 1)))),
                                  [[0, 1, 0, 0], [0, 0, 2, 1], [0, 0, 1, 0]])
    np.testing.assert_array_equal(orc.remove_bg(np.array([1, 1, 1, 0, 2]), 1), [0, 0, 0, -1, 1])
    stack = (rng.poisson(200, (6, 16, 16)) + 100).astype(np.uint16)
    np.testing.assert_array_equal(orc.remove_bg(stack, 1.1), ref.utils.remove_bg(stack, 1.1))
    for xysize in (None, 16):
        np.testing.assert_array_equal(orc.prep_data_for_pr(stack, xysize, 1.1), ref.utils.prep_data_for_PR(stack, xysize, 1.1))
    with pytest.raises(ValueError):
        orc.remove_bg(stack.astype(float), 1.0)
    raw = refshim.read_fixture_stack()
    np.testing.assert_array_equal(orc.prep_data_for_pr(raw, 256, 1.1), ref.utils.prep_data_for_PR(raw, 256, 1.1))
