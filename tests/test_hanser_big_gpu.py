"""K1b parity (GPU): the two-pass HanserPSF path for sizes outside the fused kernel's set
(N > 256 such as BASELINE config 4's 1024^2 planes, odd and non power-of-two N)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-5, "fp64": 1e-11}


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("size,vc,cond", [(512, "none", "sine"), (27, "total", "herschel"), (100, "y", "none"),
                                          (48, "none", "sine"), (1024, "none", "sine")])
def test_sizes_vs_oracle(precision, size, vc, cond):
    from oracle import pyotf_oracle as orc
    from pyotf_b200.otf import HanserPSF

    kw = dict(wl=520, na=0.85, ni=1.0, res=130)
    z = np.array([-1500.0, 0.0, 433.3]) if size < 1024 else np.array([900.0])
    m = HanserPSF(size=size, zrange=z, vec_corr=vc, condition=cond, precision=precision, **kw)
    a = orc.hanser_psfa(kw["wl"], kw["na"], kw["ni"], kw["res"], size, z, vc, cond)
    assert rel_l2(m.PSFi, orc.psfi(a)) <= TOL[precision]
    m2 = HanserPSF(size=size, zrange=z, vec_corr=vc, condition=cond, precision=precision, **kw)
    assert rel_l2(m2.PSFa, a) <= TOL[precision]
    if size <= 100:
        rng = np.random.default_rng(size)
        pupil = rng.standard_normal((size, size)) + 1j * rng.standard_normal((size, size))   # dense: support = grid
        m.apply_pupil(pupil)
        ref = orc.hanser_psfa(kw["wl"], kw["na"], kw["ni"], kw["res"], size, z, vc, cond, pupil)
        assert rel_l2(m.PSFa, ref) <= TOL[precision]
        assert rel_l2(m.PSFi, orc.psfi(ref)) <= TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_cfg4_golden(precision):
    """BASELINE config 4: 1024^2 planes of a 1024-plane stack with 15 Zernike phase terms; four planes
    (0, 511, 512, 1023: |z| up to 153.6 um, 1,856 rad of defocus phase) against the live reference."""
    from pyotf_b200.otf import HanserPSF, apply_aberration

    g = np.load(os.path.join(GOLDEN, "cfg4_hanser1024.npz"))
    tol = TOL[precision]
    chunk = HanserPSF(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=1024, zrange=g["zvals"], precision=precision)
    p = apply_aberration(chunk, None, g["pcoefs"]).PSFi
    assert p.shape == (4, 1024, 1024)
    assert rel_l2(p.sum((1, 2)), g["psfi_plane_sums"]) <= tol
    assert rel_l2(np.sqrt((p.astype(np.float64) ** 2).sum((1, 2))), g["psfi_l2"]) <= tol
    assert rel_l2(p[:, 448:576, 448:576], g["psfi_center_crops"]) <= tol
    assert rel_l2(p[:, 512, :], g["psfi_row512"]) <= tol
    big = HanserPSF(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=1024, zsize=1024)
    np.testing.assert_array_equal(big.zrange[g["plane_idx"]], g["zvals"])
