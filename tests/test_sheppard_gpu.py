"""K3 parity (GPU): SheppardPSF (spherical-cap OTF + shifted 3D inverse FFT) against the oracle and the
golden vectors from the live reference (pyotf/otf.py:446-592)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-5, "fp64": 1e-11}
SK = dict(wl=520, na=1.27, ni=1.33, res=100, size=32, zres=190, zsize=24)


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("vc,cond,dual", [("none", "sine", False), ("total", "sine", False), ("x", "herschel", False),
                                          ("z", "none", False), ("none", "none", True)])
def test_small_golden_and_oracle(precision, vc, cond, dual):
    from oracle import pyotf_oracle as orc
    from pyotf_b200.otf import SheppardPSF

    g = np.load(os.path.join(GOLDEN, "sheppard_small.npz"))
    tag = f"{vc}_{cond}_{int(dual)}"
    s = SheppardPSF(vec_corr=vc, condition=cond, dual=dual, precision=precision, **SK)
    tol = TOL[precision]
    assert rel_l2(s.PSFi, g[f"psfi_{tag}"]) <= tol
    o = orc.sheppard_otfa(SK["wl"], SK["na"], SK["ni"], SK["res"], SK["size"], SK["zres"], SK["zsize"], vc, cond, dual)
    assert s.OTFa.shape == o.shape
    assert rel_l2(s.OTFa, o) <= (1e-7 if precision == "fp32" else 1e-15)
    assert np.array_equal(s.OTFa != 0, o != 0)                      # the shell predicate is exact
    assert rel_l2(s.PSFa, orc.sheppard_psfa(o)) <= tol
    assert rel_l2(s.OTFi, orc.otfi(orc.psfi(orc.sheppard_psfa(o)))) <= tol
    # a fresh model asked for PSFa first takes the fused path (shell generated inside the first FFT pass)
    s2 = SheppardPSF(vec_corr=vc, condition=cond, dual=dual, precision=precision, **SK)
    assert rel_l2(s2.PSFa, orc.sheppard_psfa(o)) <= tol
    assert rel_l2(s2.PSFi, s.PSFi) <= tol
    s._gen_kr()
    assert int(s.valid_points.sum()) == int(g[f"nshell_{tag}"])
    valid, kzz, kyy, kxx, dk, dkz = orc.sheppard_shell(SK["wl"], SK["na"], SK["ni"], SK["res"], SK["size"], SK["zres"],
                                                       SK["zsize"], dual)
    assert np.array_equal(s.valid_points, valid)
    assert np.array_equal(s.kzz, kzz) and np.array_equal(s.kyy, kyy) and np.array_equal(s.kxx, kxx)
    assert s.dk == dk and s.dkz == dkz
    if vc == "none":
        assert rel_l2(s.PSFa, g[f"psfa_{tag}"]) <= tol
        assert rel_l2(s.OTFa, g[f"otfa_{tag}"]) <= tol


def test_same_dk_branch_and_odd_sizes():
    """dk == dkz takes the constant-thickness branch (otf.py:523-524); odd sizes exercise the shift quirk
    of easy_ifft on an fftshift-ed array (otf.py:582,592)."""
    from oracle import pyotf_oracle as orc
    from pyotf_b200.otf import SheppardPSF

    for kw in (dict(wl=520, na=1.27, ni=1.33, res=100, size=32, zres=100, zsize=32),
               dict(wl=520, na=1.1, ni=1.33, res=90, size=27, zres=150, zsize=15)):
        s = SheppardPSF(precision="fp64", **kw)
        o = orc.sheppard_otfa(kw["wl"], kw["na"], kw["ni"], kw["res"], kw["size"], kw["zres"], kw["zsize"])
        fused = SheppardPSF(precision="fp64", **kw)
        assert rel_l2(fused.PSFi, orc.psfi(orc.sheppard_psfa(o))) <= 1e-11      # fused path (no OTFa yet)
        assert rel_l2(fused.PSFa, orc.sheppard_psfa(o)) <= 1e-11
        assert np.array_equal(s.OTFa, o)
        assert rel_l2(s.PSFi, orc.psfi(orc.sheppard_psfa(o))) <= 1e-11          # generator + 3-pass path


def test_api_semantics():
    from pyotf_b200.otf import SheppardPSF

    with pytest.raises(ValueError):                                   # BASELINE config 3 as written: zres=200 >= wl/2/ni
        SheppardPSF(wl=520, na=1.27, ni=1.33, res=100, zres=200, size=256, zsize=256)
    s = SheppardPSF(500, 0.85, 1.0, 140, 32, precision="fp64")       # the reference test's model (tests/otf_test.py)
    assert np.issubdtype(s.PSFi.dtype, np.floating) and np.issubdtype(s.PSFa.dtype, np.complexfloating)
    assert np.issubdtype(s.OTFi.dtype, np.complexfloating) and np.issubdtype(s.OTFa.dtype, np.complexfloating)
    assert (s.PSFi >= 0).all()
    with pytest.raises(TypeError):
        s.dual = 1
    with pytest.raises(ValueError):
        s.res = s.wl / s.na
    assert repr(s).endswith("dual=False)")
    p0 = s.PSFi
    s.dual = True
    assert s.PSFi is not p0


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_cfg3_golden(precision):
    """BASELINE config 3 (with zres=190; 200 raises in the reference): 256^3 shell, 56,986 voxels."""
    from pyotf_b200.otf import SheppardPSF

    g = np.load(os.path.join(GOLDEN, "cfg3_sheppard256.npz"))
    s = SheppardPSF(wl=520, na=1.27, ni=1.33, res=100, zres=190, size=256, zsize=256, precision=precision)
    tol = TOL[precision]
    p = s.PSFi
    assert p.shape == (256, 256, 256)
    assert rel_l2(p.sum((1, 2)), g["psfi_plane_sums"]) <= tol
    assert rel_l2(p[g["psfi_plane_idx"]], g["psfi_planes"]) <= tol
    assert rel_l2(p[:, 128, :], g["psfi_zx"]) <= tol
    assert abs(np.linalg.norm(p.ravel().astype(np.float64)) / float(g["psfi_l2"]) - 1) <= tol
    assert int((s.OTFa != 0).sum()) == int(g["nshell"]) == 56986
    if precision == "fp64":
        assert rel_l2(s.OTFi[:, 128, :], g["otfi_zx"]) <= tol
        s._gen_kr()
        assert s.dk == float(g["dk"]) and s.dkz == float(g["dkz"])


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("size,zsize,dual,na", [(32, 32, False, 1.27), (64, 32, True, 1.27), (128, 64, False, 0.85),
                                                (256, 64, False, 1.27), (64, 128, False, 1.1)])
def test_z_first_path_matches_oracle_and_three_pass(precision, size, zsize, dual, na, monkeypatch):
    """Scalar model at power-of-two sizes: the z-first evaluation (direct z transform of the few shell voxels per
    (ky, kx) column into compact per-plane pupils, then the HanserPSF kernel K1 per plane) against the oracle
    (numpy ifftn of the reference's shell, otf.py:497-592, :204-207) and against the three-pass path."""
    from oracle import pyotf_oracle as orc
    from pyotf_b200.otf import SheppardPSF

    kw = dict(wl=520, na=na, ni=1.33, res=100, size=size, zres=190, zsize=zsize)
    tol = TOL[precision]
    o = orc.sheppard_otfa(kw["wl"], kw["na"], kw["ni"], kw["res"], kw["size"], kw["zres"], kw["zsize"], "none", "none", dual)
    assert np.count_nonzero(o) > 10
    want_a = orc.sheppard_psfa(o)
    monkeypatch.delenv("PYOTF_B200_SHEPPARD_3PASS", raising=False)
    s = SheppardPSF(dual=dual, precision=precision, **kw)
    psfi = s.PSFi                                  # fresh model: PSFi straight from the fused path
    assert rel_l2(psfi, orc.psfi(want_a)) <= tol
    s2 = SheppardPSF(dual=dual, precision=precision, **kw)
    assert rel_l2(s2.PSFa, want_a) <= tol
    monkeypatch.setenv("PYOTF_B200_SHEPPARD_3PASS", "1")
    s3 = SheppardPSF(dual=dual, precision=precision, **kw)
    assert rel_l2(s3.PSFi, psfi) <= 2 * tol
