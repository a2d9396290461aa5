"""N4: WidefieldMicroscope / ConfocalMicroscope on the device-resident PSFi against the oracle's
restatement of pyotf/microscope.py:49-227 (itself pinned to the live reference in tests/test_oracle_vs_reference.py).
Tolerances: rel. L2 <= 1e-5 (fp32), <= 1e-11 (fp64)."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-5, "fp64": 1e-11}
OPT = dict(na=0.85, ni=1.0, wl=520)


def _oracle_psfi(wl, size, res, zres, zsize):
    from oracle import pyotf_oracle as orc

    z = orc.default_zrange(zsize, zres)
    return orc.psfi(orc.hanser_psfa(wl, OPT["na"], OPT["ni"], res, size, z))


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("size,oversample", [(32, 3), (21, 3), (24, 1), (20, 5)])
def test_widefield_and_confocal_match_oracle(precision, size, oversample):
    import pyotf_b200
    from oracle import pyotf_oracle as orc
    from pyotf_b200.microscope import ConfocalMicroscope, WidefieldMicroscope

    pyotf_b200.set_precision(precision)
    try:
        kw = dict(model="hanser", size=size, pixel_size=130, oversample_factor=oversample, **OPT)
        n, res = size * oversample, 130 / oversample
        em = _oracle_psfi(520, n, res, 130, size)
        w = WidefieldMicroscope(**kw)
        ref = orc.widefield_psf(em, oversample, n)
        assert w.PSF.shape == ref.shape == (size, size, size)
        assert rel_l2(w.PSF, ref) <= TOL[precision]
        assert abs(w.PSF.sum() - 1) <= 1e-5
        assert rel_l2(w.OTF, orc.otfi(ref)) <= TOL[precision]
        exc = _oracle_psfi(488.0, n, res, 130, size)
        for pinhole in (1.0, 2.5, 0.05):
            c = ConfocalMicroscope(wl_exc=488.0, pinhole_size=pinhole, **kw)
            mref = orc.confocal_model_psf(em, exc, 520, 0.85, res, pinhole)
            assert rel_l2(c.model_psf, mref) <= TOL[precision]
            pref = orc.widefield_psf(mref, oversample, n)
            assert rel_l2(c.PSF, pref) <= TOL[precision]
            assert rel_l2(c.OTF, orc.otfi(pref)) <= TOL[precision]
    finally:
        pyotf_b200.set_precision("fp64")


def test_validation_like_the_reference():
    """oversample_factor must be an odd positive int; unknown models raise ValueError (pyotf/microscope.py:39-76)."""
    from pyotf_b200.microscope import WidefieldMicroscope

    kw = dict(model="hanser", size=16, pixel_size=130, **OPT)
    with pytest.raises(AssertionError):
        WidefieldMicroscope(oversample_factor=2, **kw)
    with pytest.raises(TypeError):
        WidefieldMicroscope(oversample_factor=3.0, **kw)
    with pytest.raises(ValueError):
        WidefieldMicroscope(oversample_factor=3, **dict(kw, model="nope"))
    assert "WidefieldMicroscope(model='hanser'" in repr(WidefieldMicroscope(oversample_factor=3, **kw))
