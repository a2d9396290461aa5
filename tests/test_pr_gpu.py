"""K2 parity (GPU): the fused phase-retrieval iteration against the CPU oracle
(oracle.retrieve_phase_loop, a restatement of pyotf/phaseretrieval.py:79-135) and against the
golden vectors generated from the live reference (cfg2 = the fixture stack; the
integration-test twin).  Everything goes through retrieve_phase -> C ABI -> CUDA."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu

OPT = dict(wl=520, na=0.85, ni=1.0, res=130, zres=300)


def _synthetic_stack(size, nz, seed=0, **kw):
    """PSFi of a randomly aberrated pupil (like tests/integration_test.py of the reference), via the oracle."""
    from oracle import pyotf_oracle as orc

    rng = np.random.default_rng(seed)
    p = dict(OPT, **kw)
    z = orc.default_zrange(nz, p["zres"])
    pupil = orc.aberrated_pupil(p["wl"], p["na"], p["ni"], p["res"], size, rng.random(15) * 0.1, rng.random(15))
    psfa = orc.hanser_psfa(p["wl"], p["na"], p["ni"], p["res"], size, z, "none", "none", pupil)
    return orc.psfi(psfa) * 1e4, z


@pytest.mark.parametrize("size,nz,rows", [(64, 5, 0), (64, 7, 16), (64, 4, 64), (128, 6, 32), (128, 3, 64), (256, 4, 0),
                                          (32, 3, 16), (32, 9, 32)])
def test_fp64_free_running_vs_oracle(size, nz, rows):
    from oracle import pyotf_oracle as orc
    from pyotf_b200.phaseretrieval import retrieve_phase

    data, z = _synthetic_stack(size, nz, seed=size + nz)
    iters = 12
    o = orc.retrieve_phase_loop(data, OPT["wl"], OPT["na"], OPT["ni"], OPT["res"], zres=OPT["zres"], max_iters=iters,
                                pupil_tol=0, mse_tol=0)
    r = retrieve_phase(data, dict(OPT), iters, 0, 0, precision="fp64", rows_per_cta=rows)
    assert len(r.mse) == iters
    assert rel_l2(r.mse, o["mse"]) <= 1e-11
    np.testing.assert_allclose(r.mse_diff[1:], o["mse_diff"][1:], rtol=1e-8)
    np.testing.assert_allclose(r.pupil_diff[1:], o["pupil_diff"][1:], rtol=1e-8)
    assert np.isnan(r.mse_diff[0]) and np.isnan(r.pupil_diff[0])
    assert rel_l2(r.pupil, o["pupil"]) <= 1e-11
    # result.model carries the last *applied* pupil: its PSFi is the oracle's last model PSFi
    assert rel_l2(r.model.PSFi, o["psfi"]) <= 1e-11
    mask = np.fft.fftshift(o["mask"])
    assert rel_l2(r.mag, np.fft.fftshift(abs(o["pupil"])) * mask) <= 1e-11


def test_fp32_single_step_and_mse():
    """fp32 mode: each iteration re-seeded from the oracle's pupil reproduces the next pupil to 1e-5
    (free-running fp32 pupils drift, SURVEY.md section 7 hard part 3); the free-running MSE holds 1e-5."""
    import torch

    from oracle import pyotf_oracle as orc
    from pyotf_b200 import _engine as eng
    from pyotf_b200.phaseretrieval import retrieve_phase

    size, nz, iters = 128, 8, 10
    data, z = _synthetic_stack(size, nz, seed=3)
    o = orc.retrieve_phase_loop(data, OPT["wl"], OPT["na"], OPT["ni"], OPT["res"], zres=OPT["zres"], max_iters=iters,
                                pupil_tol=0, mse_tol=0, return_history=True)
    r = retrieve_phase(data, dict(OPT), iters, 0, 0, precision="fp32")
    assert rel_l2(r.mse, o["mse"]) <= 1e-5
    st = eng.PRState(OPT["wl"], OPT["na"], OPT["ni"], OPT["res"], size, z,
                     torch.as_tensor(data.astype(np.float32), device="cuda"), "fp32")
    hist = o["history"] + [o["pupil"]]
    for i in range(iters):
        st.set_pupil(eng.to_device_c128(hist[i]))
        mse, _, _ = st.run(1, 0, 0)
        assert abs(mse[0] / o["mse"][i] - 1) <= 1e-5
        assert rel_l2(st.get_pupil().cpu().numpy(), hist[i + 1]) <= 1e-5, f"iteration {i}"
        assert rel_l2(st.get_pupil(applied=True).cpu().numpy(), hist[i]) <= 1e-6


def test_phase_only_and_early_stop():
    from oracle import pyotf_oracle as orc
    from pyotf_b200.phaseretrieval import retrieve_phase

    data, z = _synthetic_stack(64, 6, seed=11)
    o = orc.retrieve_phase_loop(data, OPT["wl"], OPT["na"], OPT["ni"], OPT["res"], zres=OPT["zres"], max_iters=40,
                                pupil_tol=0, mse_tol=0, phase_only=True)
    r = retrieve_phase(data, dict(OPT), 40, 0, 0, phase_only=True, precision="fp64")
    assert rel_l2(r.mse, o["mse"]) <= 1e-10
    assert rel_l2(r.pupil, o["pupil"]) <= 1e-9
    # early stop: same iteration count and the same final pupil as the reference loop
    for ptol, mtol in ((2e-3, 0), (0, 0.09), (1e-3, 0.05)):
        o = orc.retrieve_phase_loop(data, OPT["wl"], OPT["na"], OPT["ni"], OPT["res"], zres=OPT["zres"], max_iters=60,
                                    pupil_tol=ptol, mse_tol=mtol)
        r = retrieve_phase(data, dict(OPT), 60, ptol, mtol, precision="fp64")
        assert len(r.mse) == o["iters"] and o["iters"] < 60
        assert rel_l2(r.mse, o["mse"]) <= 1e-10
        assert rel_l2(r.pupil, o["pupil"]) <= 1e-9
        assert rel_l2(r.model.PSFi, o["psfi"]) <= 1e-9


def test_argument_checks():
    from pyotf_b200.phaseretrieval import retrieve_phase

    d = np.ones((3, 64, 64))
    with pytest.raises(AssertionError):
        retrieve_phase(d, dict(OPT), 0)
    with pytest.raises(AssertionError):
        retrieve_phase(np.ones((3, 64, 32)), dict(OPT), 5)
    with pytest.raises((AssertionError, IndexError)):   # the reference indexes shape[2] before its ndim assert
        retrieve_phase(np.ones((64, 64)), dict(OPT), 5)
    params = dict(OPT, vec_corr="total", condition="sine")
    r = retrieve_phase(d, params, 2, precision="fp64")
    assert params["vec_corr"] == "none" and params["condition"] == "none" and params["size"] == 64 and params["zsize"] == 3
    assert r.mag.shape == (64, 64) and r.phase.shape == (64, 64) and r.model.size == 64


def _fixture_data():
    from pyotf_b200.utils import prep_data_for_PR

    raw = np.load(os.path.join(GOLDEN, "fixture_psf_520.npz"))["raw"]
    return prep_data_for_PR(raw, 256, 1.1)


def test_cfg2_fixture_golden_fp64():
    """BASELINE config 2 on the reference's measured stack: early stop after 51 iterations at tol 1e-4,
    and the fixed 100-iteration run, against the live reference's numbers."""
    from pyotf_b200.phaseretrieval import retrieve_phase

    g = np.load(os.path.join(GOLDEN, "cfg2_phase_retrieval.npz"))
    data = _fixture_data()
    assert abs(data.sum() / float(g["data_sum"]) - 1) < 1e-15 and data.max() == float(g["data_max"])
    r = retrieve_phase(data, dict(OPT), 100, 1e-4, 1e-4, precision="fp64")
    assert len(r.mse) == len(g["es_mse"]) == 51
    assert rel_l2(r.mse, g["es_mse"]) <= 1e-11
    np.testing.assert_allclose(r.mse_diff[1:], g["es_mse_diff"][1:], rtol=1e-7)
    np.testing.assert_allclose(r.pupil_diff[1:], g["es_pupil_diff"][1:], rtol=1e-7)
    assert rel_l2(r.mag, g["es_mag"]) <= 1e-11
    ref_pupil = g["es_mag"] * np.exp(1j * g["es_angle"])
    assert rel_l2(np.fft.fftshift(r.pupil) * (g["es_mag"] > 0), ref_pupil) <= 1e-10
    assert rel_l2(r.model.PSFi[13], g["es_model_psfi_plane13"]) <= 1e-11
    # known-answer values, SURVEY.md section 8c
    assert abs(r.mse[0] / 43196.9997303445 - 1) <= 1e-11 and abs(r.mse[50] / 5120.289982201339 - 1) <= 1e-10
    assert abs(r.pupil_diff[50] / 9.10097282071048e-05 - 1) <= 1e-7
    r = retrieve_phase(data, dict(OPT), 100, 0, 0, precision="fp64")
    assert len(r.mse) == 100
    assert rel_l2(r.mse, g["full_mse"]) <= 1e-11
    assert rel_l2(r.mag, g["full_mag"]) <= 1e-10
    assert abs(r.mag.sum() / 20776077.656680964 - 1) <= 1e-10
    for k in (1, 2, 10):
        rk = retrieve_phase(data, dict(OPT), k, 0, 0, precision="fp64")
        assert rel_l2(rk.pupil, g[f"pupil_after_{k}"]) <= 1e-11


def test_cfg2_fixture_fp32():
    """fp32 mode on the fixture: per-iteration MSE within 1e-5 and the same early-stop iteration."""
    from pyotf_b200.phaseretrieval import retrieve_phase

    g = np.load(os.path.join(GOLDEN, "cfg2_phase_retrieval.npz"))
    data = _fixture_data()
    r = retrieve_phase(data, dict(OPT), 100, 1e-4, 1e-4, precision="fp32")
    assert len(r.mse) == 51
    assert rel_l2(r.mse, g["es_mse"]) <= 1e-5
    assert rel_l2(r.mag, g["es_mag"]) <= 2e-4     # free-running fp32 pupil: drift documented in DESIGN.md
    r = retrieve_phase(data, dict(OPT), 100, 0, 0, precision="fp32")
    assert rel_l2(r.mse, g["full_mse"]) <= 1e-5


def test_integration_twin_golden():
    """tests/integration_test.py of the reference: explicit zrange, 128^2, random Noll 5-15 pupil."""
    from pyotf_b200.phaseretrieval import retrieve_phase

    g = np.load(os.path.join(GOLDEN, "integration_twin.npz"))
    mk = dict(wl=525, na=1.27, ni=1.33, res=100, size=128, zrange=[-1000, -500, 0, 250, 1000, 3000],
              vec_corr="none", condition="none")
    r = retrieve_phase(g["psfi"], dict(mk), max_iters=60, pupil_tol=0, mse_tol=0, precision="fp64")
    assert rel_l2(r.mse, g["mse60"]) <= 1e-9
    np.testing.assert_allclose(r.pupil_diff[1:], g["pupil_diff60"][1:], rtol=1e-6)
    assert rel_l2(r.mag, g["mag60"]) <= 1e-9
    # the reference's own assertions (integration_test.py:65-90): mse decreases, result is self-consistent
    assert (np.diff(r.mse[1:]) < 0).all()
    from oracle import pyotf_oracle as orc

    ref = orc.psfi(orc.hanser_psfa(mk["wl"], mk["na"], mk["ni"], mk["res"], 128, mk["zrange"], "none", "none", r.pupil))
    assert rel_l2(r.generate_psf(), ref) <= 1e-11          # PSF of the final pupil (mag * exp(i phase))
    assert r.generate_psf(size=64).shape == (6, 64, 64) and r.generate_psf(size=200, zsize=3).shape == (3, 200, 200)
    zd = r.fit_to_zernikes(15)
    assert zd.mcoefs.shape == (15,) and zd.pcoefs.shape == (15,)
    assert np.isfinite(r.rms_error) and np.isfinite(r.pv_error)


def test_integration_twin_400_iterations_recovers_the_pupil():
    """The reference's integration test proper (tests/integration_test.py:60-90): 400 iterations recover
    the magnitude, the (unwrapped) phase, the Zernike coefficients and the PSF of the synthetic pupil."""
    from numpy.fft import fftshift

    from pyotf_b200.phaseretrieval import retrieve_phase

    g = np.load(os.path.join(GOLDEN, "integration_twin.npz"))
    mk = dict(wl=525, na=1.27, ni=1.33, res=100, size=128, zrange=[-1000, -500, 0, 250, 1000, 3000],
              vec_corr="none", condition="none")
    r = retrieve_phase(g["psfi"], dict(mk), max_iters=400, pupil_tol=0, mse_tol=0, precision="fp64")
    assert len(r.mse) == 400
    assert rel_l2(r.mse[:100], g["mse400"][:100]) <= 1e-8
    assert rel_l2(r.mag, g["mag400"]) <= 1e-11
    mask = g["mag400"] > 0
    np.testing.assert_allclose(fftshift(g["pupil_mag"]) * mask, r.mag, rtol=1e-7, atol=1e-10)       # test_mag
    d = (fftshift(g["pupil_phase"]) - r.phase)[mask]                                               # test_phase
    # "a constant offset is normal": piston is not observable.  A pixel on the aperture edge may sit on another
    # 2 pi branch (its unwrapping path runs through the zero region outside the pupil).
    resid = d - np.median(d)
    wraps = np.round(resid / (2 * np.pi))
    assert abs(resid - 2 * np.pi * wraps).max() <= 1e-6
    assert (wraps != 0).sum() <= 3
    r.phase = r.phase + 2 * np.pi * np.where(mask, 0, 0)
    r.phase[mask] += 2 * np.pi * wraps                   # put those pixels on the common branch for the fit below
    np.testing.assert_allclose(r.model.PSFi, g["psfi"], rtol=1e-7, atol=1e-12)                      # test_psf_mse
    zd = r.fit_to_zernikes(15)
    np.testing.assert_allclose(zd.pcoefs[4:], g["pcoefs"], rtol=1e-6, atol=1e-8)                    # test_zernike_modes_phase
    np.testing.assert_allclose(zd.mcoefs[4:], g["mcoefs"], rtol=1e-6, atol=1e-8)                    # test_zernike_modes_mag


@pytest.mark.parametrize("size,nz,precision", [(64, 5, "fp64"), (128, 4, "fp64"), (128, 4, "fp32"), (100, 3, "fp64"),
                                               (45, 4, "fp64"), (63, 5, "fp64"), (512, 3, "fp64"), (512, 3, "fp32")])
def test_generic_size_path_vs_oracle(size, nz, precision):
    """K2b, the four-pass path for sizes outside {32, 64, 128, 256} (forced with rows_per_cta=1 at 64 / 128 so
    that it is also checked against the fused kernel).

    The planes are offset from z = 0: with the ideal start pupil the in-focus amplitude is exactly real and,
    for sizes such as 45 or 63, exactly zero on whole columns, where angle(PSFa) is decided by the rounding
    noise of whichever FFT computed it (the reference's own result is arbitrary there)."""
    from oracle import pyotf_oracle as orc
    from pyotf_b200.phaseretrieval import retrieve_phase

    rng = np.random.default_rng(size)
    z = orc.default_zrange(nz, OPT["zres"]) + 41.0
    pupil = orc.aberrated_pupil(OPT["wl"], OPT["na"], OPT["ni"], OPT["res"], size, rng.random(15) * 0.1, rng.random(15))
    data = orc.psfi(orc.hanser_psfa(OPT["wl"], OPT["na"], OPT["ni"], OPT["res"], size, z, "none", "none", pupil)) * 1e4
    params = dict(wl=OPT["wl"], na=OPT["na"], ni=OPT["ni"], res=OPT["res"], zrange=z)
    iters = 8
    o = orc.retrieve_phase_loop(data, OPT["wl"], OPT["na"], OPT["ni"], OPT["res"], zrange=z, max_iters=iters,
                                pupil_tol=0, mse_tol=0)
    r = retrieve_phase(data, dict(params), iters, 0, 0, precision=precision, rows_per_cta=1)
    tol = 1e-11 if precision == "fp64" else 1e-5
    assert len(r.mse) == iters
    assert rel_l2(r.mse, o["mse"]) <= tol
    assert rel_l2(r.pupil, o["pupil"]) <= (tol if precision == "fp64" else 2e-5)
    assert rel_l2(r.model.PSFi, o["psfi"]) <= tol
    if size in (64, 128):
        f = retrieve_phase(data, dict(params), iters, 0, 0, precision=precision)       # fused kernel
        assert rel_l2(r.pupil, f.pupil) <= (1e-12 if precision == "fp64" else 2e-5)
    # early stop decides on the same iteration as the reference loop
    o2 = orc.retrieve_phase_loop(data, OPT["wl"], OPT["na"], OPT["ni"], OPT["res"], zrange=z, max_iters=40,
                                 pupil_tol=3e-3, mse_tol=0)
    r2 = retrieve_phase(data, dict(params), 40, 3e-3, 0, precision=precision, rows_per_cta=1)
    assert len(r2.mse) == o2["iters"]


def test_fit_to_zernikes_device_normal_equations_vs_host_lstsq():
    """fit_to_zernikes: GPU normal equations + host solve against the reference's formulation
    (numpy.linalg.lstsq on the full mode stack, phaseretrieval.py:406-432), 120 Noll modes at 256^2."""
    from pyotf_b200.phaseretrieval import _fit_to_zerns, _recon_from_zerns, retrieve_phase
    from pyotf_b200.zernike import noll2degrees, zernike

    data = _fixture_data()
    r = retrieve_phase(data, dict(OPT), 20, 0, 0, precision="fp64")
    zd = r.fit_to_zernikes(120)
    rr = r.r / (r.na / r.wl)
    zerns = zernike(rr, r.theta, *noll2degrees(np.arange(1, 121)))
    ref_m, ref_p = _fit_to_zerns(r.mag, zerns, rr), _fit_to_zerns(r.phase, zerns, rr)
    assert rel_l2(zd.mcoefs, ref_m) <= 1e-9 and rel_l2(zd.pcoefs, ref_p) <= 1e-9
    assert zd.zerns.shape == (120, 256, 256) and rel_l2(zd.zerns, zerns) == 0.0
    assert rel_l2(zd.phase(slice(4, None)), _recon_from_zerns(ref_p[4:], zerns[4:])) <= 1e-9
    assert zd.complex_pupil().shape == (256, 256) and np.isfinite(zd.rms_error)
    assert r.generate_zd_psf().shape == (25, 256, 256)


@pytest.mark.parametrize("size,nz", [(64, 1), (64, 300), (256, 70), (32, 2)])
def test_stack_length_edge_cases(size, nz):
    """One plane, and stacks long enough for several waves of plane kernels / the narrow-CTA variant."""
    from oracle import pyotf_oracle as orc
    from pyotf_b200.phaseretrieval import retrieve_phase

    rng = np.random.default_rng(nz)
    z = (np.arange(nz) - nz // 2) * 110.0 + 17.0
    pupil = orc.aberrated_pupil(OPT["wl"], OPT["na"], OPT["ni"], OPT["res"], size, rng.random(8) * 0.1, rng.random(8))
    data = orc.psfi(orc.hanser_psfa(OPT["wl"], OPT["na"], OPT["ni"], OPT["res"], size, z, "none", "none", pupil)) * 1e4
    params = dict(wl=OPT["wl"], na=OPT["na"], ni=OPT["ni"], res=OPT["res"], zrange=z)
    o = orc.retrieve_phase_loop(data, OPT["wl"], OPT["na"], OPT["ni"], OPT["res"], zrange=z, max_iters=5, pupil_tol=0,
                                mse_tol=0)
    for precision, tol in (("fp64", 1e-11), ("fp32", 1e-5)):
        r = retrieve_phase(data, dict(params), 5, 0, 0, precision=precision)
        assert rel_l2(r.mse, o["mse"]) <= tol
        assert rel_l2(r.pupil, o["pupil"]) <= 2 * tol


def test_phase_retrieval_object_keeps_state_step_and_set_pupil():
    """PhaseRetrieval (object form): run() == retrieve_phase(); step(n) continues from the current estimate, so
    5 + 7 single-launch steps equal one 12-iteration run; set_pupil() re-seeds the estimate."""
    from pyotf_b200.otf import HanserPSF, apply_aberration
    from pyotf_b200.phaseretrieval import PhaseRetrieval, retrieve_phase

    z = (np.arange(6) - 3) * 250.0 + 40.0
    m = HanserPSF(520, 0.85, 1.0, 130, 64, zrange=z, vec_corr="none", condition="none", precision="fp64")
    data = apply_aberration(m, None, np.array([0, 0, 0, 0, 0.3, 0.2, -0.4, 0.1])).PSFi * 1e4
    params = dict(wl=520, na=0.85, ni=1.0, res=130, zrange=z)
    full = retrieve_phase(data, dict(params), 12, 0, 0, precision="fp64")
    pr = PhaseRetrieval(data, dict(params), precision="fp64")
    r = pr.run(12, 0, 0)
    np.testing.assert_array_equal(r.mse, full.mse)
    np.testing.assert_array_equal(r.pupil, full.pupil)
    pr.set_pupil(None)
    a = pr.step(5)[0]
    b = pr.step(7)[0]
    np.testing.assert_allclose(np.concatenate([a, b]), full.mse, rtol=1e-13)
    np.testing.assert_allclose(pr.pupil, full.pupil, rtol=0, atol=1e-13 * abs(full.pupil).max())
    # re-seeding with the retrieved pupil reproduces the continuation
    pr.set_pupil(full.pupil)
    c = pr.step(1)[0]
    more = retrieve_phase(data, dict(params), 13, 0, 0, precision="fp64")
    assert abs(c[0] / more.mse[12] - 1) <= 1e-10


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("nz,phase_only,mtol", [(25, False, 0.0), (19, True, 0.0), (25, False, 2e-3)])
def test_one_kernel_per_iteration_equals_two_kernel_loop(precision, nz, phase_only, mtol, monkeypatch):
    """The fused loop (plane kernel + in-kernel finalize behind a grid-wide arrival barrier: one launch per iteration,
    stacks whose plane-kernel grid is resident at once) sums in the order of the two-kernel loop: bit-identical
    per-iteration statistics, stop iteration and pupils (256^2, the fixture's shape)."""
    import torch

    from pyotf_b200 import _engine as eng

    data, z = _synthetic_stack(256, nz, seed=nz)
    rt = torch.float32 if precision == "fp32" else torch.float64
    dev = torch.as_tensor(data, device="cuda").to(rt)
    out = {}
    for label in ("fused", "fused1", "two"):
        # "fused": up to 8 iterations per launch behind in-kernel grid barriers; "fused1": one launch per iteration
        monkeypatch.delenv("PYOTF_B200_PR_NO_FUSE", raising=False)
        monkeypatch.delenv("PYOTF_B200_PR_ITERS_PER_LAUNCH", raising=False)
        if label == "two":
            monkeypatch.setenv("PYOTF_B200_PR_NO_FUSE", "1")
        elif label == "fused1":
            monkeypatch.setenv("PYOTF_B200_PR_ITERS_PER_LAUNCH", "1")
        st = eng.PRState(OPT["wl"], OPT["na"], OPT["ni"], OPT["res"], 256, z, dev, precision)
        assert st.kernels_per_iteration() == (1 if label != "two" and precision == "fp32" else 2)
        mse, mse_diff, pupil_diff = st.run(30, 0.0, mtol, phase_only=phase_only)
        out[label] = (np.array(mse), np.array(mse_diff), np.array(pupil_diff), st.get_pupil(0).cpu().numpy(), st.get_pupil(1).cpu().numpy())
        # a second run on the same state (PhaseRetrieval.step): the arrival counter restarts
        mse2, _, _ = st.run(3, 0.0, 0.0, phase_only=phase_only)
        out[label] += (np.array(mse2),)
    a, b = out["fused"], out["two"]
    if mtol:
        assert len(a[0]) == len(b[0]) < 30
    for i, (x, y) in enumerate(zip(a, b)):
        if i == 2:          # pupil_diff: the per-CTA sums of |old - new|^2 are grouped differently (124 vs 32 pixels per CTA)
            np.testing.assert_allclose(x[1:], y[1:], rtol=1e-12, atol=0)
        else:
            np.testing.assert_array_equal(x, y)
    for x, y in zip(out["fused"], out["fused1"]):      # the same kernel, 8 or 1 iterations per launch: everything identical
        np.testing.assert_array_equal(x, y)
