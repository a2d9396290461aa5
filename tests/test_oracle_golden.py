"""Pin oracle/pyotf_oracle.py against the committed golden vectors (tests/golden/*.npz, generated from the
LIVE upstream pyotf by oracle/make_golden.py).  Runs anywhere (no GPU, no /root/reference): this is the pin
that travels to the GPU box, where the CUDA path is then compared with the oracle on fresh inputs."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_l2
from oracle import pyotf_oracle as orc

TIGHT = 1e-13   # same numpy/pocketfft arithmetic as the reference: only summation-order noise is allowed


def _g(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.mark.parametrize("vc", ["none", "x", "y", "z", "total"])
@pytest.mark.parametrize("cond", ["none", "sine", "herschel"])
def test_hanser_small(vc, cond):
    g = _g("hanser_small.npz")
    wl, na, ni, res, size, zres, zsize = g["params"]
    z = orc.default_zrange(int(zsize), zres)
    a = orc.hanser_psfa(wl, na, ni, res, int(size), z, vc, cond)
    assert rel_l2(orc.psfi(a), g[f"psfi_{vc}_{cond}"]) <= TIGHT
    if f"psfa_{vc}_{cond}" in g.files:
        assert rel_l2(a, g[f"psfa_{vc}_{cond}"]) <= TIGHT
    if (vc, cond) == ("none", "sine"):
        assert rel_l2(orc.otfi(orc.psfi(a)), g["otfi_none_sine"]) <= TIGHT
        assert rel_l2(orc.hanser_otfa(a), g["otfa_none_sine"]) <= TIGHT


def test_cfg1_known_answers():
    g = _g("cfg1_hanser256.npz")
    z = orc.default_zrange(128, 300)
    np.testing.assert_array_equal(z, g["zrange"])
    p = orc.psfi(orc.hanser_psfa(520, 0.85, 1.0, 130, 256, z))
    assert rel_l2(p[g["psfi_plane_idx"]], g["psfi_planes"]) <= TIGHT
    assert rel_l2(p.sum((1, 2)), g["psfi_plane_sums"]) <= TIGHT
    assert abs(p.sum() / 23.75643331623896 - 1) <= 1e-13
    assert abs(p[64, 128, 128] / 0.02609296492831893 - 1) <= 1e-13
    assert abs(p[0, 128, 128] / 1.1464154225339048e-05 - 1) <= 1e-11
    o = orc.otfi(p)
    assert rel_l2(o[g["otfi_plane_idx"]], g["otfi_planes"]) <= TIGHT
    assert abs(o[64, 128, 140].real / 0.9735855131724713 - 1) <= 1e-12


def test_zernike_and_aberration():
    g = _g("zernike_noll120.npz")
    n, m = orc.noll_to_degrees(np.arange(1, 121))
    np.testing.assert_array_equal(n, g["n"])
    np.testing.assert_array_equal(m, g["m"])
    np.testing.assert_allclose(orc.zernike_modes(g["r"], g["theta"], n, m), g["zern"], rtol=0, atol=2e-12)
    np.testing.assert_allclose(orc.zernike_modes(g["r"], g["theta"], n[:36], m[:36], norm=False), g["zern_unnormed"],
                               rtol=0, atol=2e-12)
    g = _g("aberration_small.npz")
    z = orc.default_zrange(4, 300)
    for mc, pc, key in ((g["mcoefs"], g["pcoefs"], "psfi_both"), (None, g["pcoefs"], "psfi_phase"),
                        (g["mcoefs"], None, "psfi_mag")):
        pupil = orc.aberrated_pupil(520, 0.85, 1.0, 130, 64, mc, pc)
        assert rel_l2(orc.psfi(orc.hanser_psfa(520, 0.85, 1.0, 130, 64, z, pupil_base=pupil)), g[key]) <= 1e-12


def test_cfg5_unit_entry():
    g = _g("cfg5_library.npz")
    one = np.random.default_rng(0).standard_normal(120) * 0.1
    pupil = orc.aberrated_pupil(520, 0.85, 1.0, 130, 256, None, one)
    p = orc.psfi(orc.hanser_psfa(520, 0.85, 1.0, 130, 256, orc.default_zrange(64, 300), pupil_base=pupil))
    assert rel_l2(p[32], g["unit_psfi_plane32"]) <= 1e-12
    assert abs(p.sum() / float(g["unit_psfi_sum"]) - 1) <= 1e-12
    assert abs(p[32, 128, 128] / 0.009643743189627898 - 1) <= 1e-11


def test_cfg4_planes():
    g = _g("cfg4_hanser1024.npz")
    pupil = orc.aberrated_pupil(520, 0.85, 1.0, 130, 1024, None, g["pcoefs"])
    p = orc.psfi(orc.hanser_psfa(520, 0.85, 1.0, 130, 1024, g["zvals"][:2], pupil_base=pupil))
    assert rel_l2(p[:, 512, :], g["psfi_row512"][:2]) <= 1e-12
    assert rel_l2(p.sum((1, 2)), g["psfi_plane_sums"][:2]) <= 1e-12


def test_phase_retrieval_fixture_first_iterations():
    """cfg2: the first 12 iterations on the reference's measured stack (the full 51 / 100 iteration runs are
    compared on the GPU against the same golden file; the oracle's full runs are pinned in
    tests/test_oracle_vs_reference.py)."""
    from pyotf_b200.utils import prep_data_for_PR

    g = _g("cfg2_phase_retrieval.npz")
    data = prep_data_for_PR(_g("fixture_psf_520.npz")["raw"], 256, 1.1)
    assert data.sum() == float(g["data_sum"]) and data.max() == float(g["data_max"])
    o = orc.retrieve_phase_loop(data, 520, 0.85, 1.0, 130, zres=300, max_iters=12, pupil_tol=0, mse_tol=0,
                                return_history=True)
    assert rel_l2(o["mse"], g["full_mse"][:12]) <= 1e-13
    np.testing.assert_allclose(o["pupil_diff"][1:], g["full_pupil_diff"][1:12], rtol=1e-10)
    np.testing.assert_allclose(o["mse_diff"][1:], g["full_mse_diff"][1:12], rtol=1e-10)
    hist = o["history"] + [o["pupil"]]
    for k in (1, 2, 10):
        assert rel_l2(hist[k], g[f"pupil_after_{k}"]) <= 1e-12
    assert abs(o["mse"][0] / 43196.9997303445 - 1) <= 1e-13 and abs(o["mse"][10] / 40683.74861451405 - 1) <= 1e-12


def test_integration_twin():
    g = _g("integration_twin.npz")
    o = orc.retrieve_phase_loop(g["psfi"], 525, 1.27, 1.33, 100, zrange=[-1000, -500, 0, 250, 1000, 3000],
                                max_iters=60, pupil_tol=0, mse_tol=0)
    assert rel_l2(o["mse"], g["mse60"]) <= 1e-10
    assert rel_l2(np.fft.fftshift(abs(o["pupil"]) * o["mask"]), g["mag60"]) <= 1e-11


@pytest.mark.parametrize("vc,cond,dual", [("none", "sine", False), ("total", "sine", False), ("x", "herschel", False),
                                          ("z", "none", False), ("none", "none", True)])
def test_sheppard_small(vc, cond, dual):
    g = _g("sheppard_small.npz")
    tag = f"{vc}_{cond}_{int(dual)}"
    o = orc.sheppard_otfa(520, 1.27, 1.33, 100, 32, 190, 24, vc, cond, dual)
    valid = orc.sheppard_shell(520, 1.27, 1.33, 100, 32, 190, 24, dual)[0]
    assert int(valid.sum()) == int(g[f"nshell_{tag}"])
    assert rel_l2(orc.psfi(orc.sheppard_psfa(o)), g[f"psfi_{tag}"]) <= TIGHT
    if vc == "none":
        np.testing.assert_array_equal(o, g[f"otfa_{tag}"])
        assert rel_l2(orc.sheppard_psfa(o), g[f"psfa_{tag}"]) <= TIGHT


def test_sheppard_cfg3_shell_count():
    valid, *_rest, dk, dkz = orc.sheppard_shell(520, 1.27, 1.33, 100, 256, 190, 256)
    g = _g("cfg3_sheppard256.npz")
    assert int(valid.sum()) == int(g["nshell"]) == 56986
    assert dk == float(g["dk"]) and dkz == float(g["dkz"])
