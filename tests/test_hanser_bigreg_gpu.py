"""K1bs / K1br parity (GPU): the register-resident two-pass HanserPSF paths at 1024^2 (csrc/hanser_bigsym.cuh) --
the default model through the even-symmetric kernels, aberrated / vectorial models through the general ones --
against the oracle on several planes at once (BASELINE config 4 is tests/test_hanser_big_gpu.py::test_cfg4_golden)."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-5, "fp64": 1e-11}
KW = dict(wl=520, na=0.85, ni=1.0, res=130)
Z = np.array([-7300.0, -300.0, 0.0, 1250.5, 40000.0])


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_default_model_symmetric_kernels(precision):
    from oracle import pyotf_oracle as orc
    from pyotf_b200.otf import HanserPSF

    m = HanserPSF(size=1024, zrange=Z, precision=precision, **KW)
    ref = orc.psfi(orc.hanser_psfa(KW["wl"], KW["na"], KW["ni"], KW["res"], 1024, Z, "none", "sine"))
    p = m.PSFi
    assert p.shape == ref.shape
    for i in range(len(Z)):
        assert rel_l2(p[i], ref[i]) <= TOL[precision], i
    # the symmetries the kernels rely on are those of the reference's own result
    np.testing.assert_array_equal(p[:, 1:, 1:], p[:, 1:, 1:][:, ::-1, ::-1])
    # row 0 / column 0 (the Nyquist row the kernels take from the transpose) are held to the oracle separately
    assert rel_l2(p[:, 0, :], ref[:, 0, :]) <= 10 * TOL[precision]
    assert rel_l2(p[:, :, 0], ref[:, :, 0]) <= 10 * TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("vc,cond", [("none", "sine"), ("total", "herschel")])
def test_general_kernels_user_pupil_and_vectorial(precision, vc, cond):
    from oracle import pyotf_oracle as orc
    from pyotf_b200.otf import HanserPSF

    z = Z[1:4]
    rng = np.random.default_rng(7)
    m = HanserPSF(size=1024, zrange=z, vec_corr=vc, condition=cond, precision=precision, **KW)
    if vc == "none":
        # a non-symmetric complex pupil inside the ideal support
        k = np.fft.fftfreq(1024, KW["res"])
        kr = np.hypot(*np.meshgrid(k, k, indexing="ij"))
        mask = kr <= KW["na"] / KW["wl"]
        pupil = mask * np.exp(1j * rng.standard_normal((1024, 1024))) * (1 + 0.3 * rng.standard_normal((1024, 1024)))
        m.apply_pupil(pupil)
        ref = orc.hanser_psfa(KW["wl"], KW["na"], KW["ni"], KW["res"], 1024, z, vc, cond, pupil)
    else:
        ref = orc.hanser_psfa(KW["wl"], KW["na"], KW["ni"], KW["res"], 1024, z, vc, cond)
    assert rel_l2(m.PSFi, orc.psfi(ref)) <= TOL[precision]
