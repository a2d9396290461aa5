"""Seeded random sweep (GPU): optics, sizes, plane positions, polarisation / apodisation modes and pupils drawn at
random; HanserPSF and a few phase-retrieval iterations must match the oracle at the stated tolerances.  Covers the
combinations of support-box size (odd / even H, table pitch), rows-per-CTA choice and symmetry path that the
fixed-parameter tests do not enumerate."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _draw(rng):
    size = int(rng.choice([32, 64, 128, 256, 48, 96]))
    wl = float(rng.uniform(450, 700))
    ni = float(rng.choice([1.0, 1.33, 1.518]))
    na = float(rng.uniform(0.3, 0.95) * ni)
    abbe = wl / (2 * na)
    res = float(rng.uniform(0.25, 0.95) * abbe)
    nz = int(rng.integers(1, 7))
    z = np.sort(rng.uniform(-3000, 3000, nz))
    return size, wl, na, ni, res, z


@pytest.mark.parametrize("seed", range(24))
def test_hanser_random(seed):
    from oracle import pyotf_oracle as orc
    from pyotf_b200.otf import HanserPSF

    rng = np.random.default_rng(1000 + seed)
    size, wl, na, ni, res, z = _draw(rng)
    vc = str(rng.choice(["none", "none", "x", "y", "z", "total"]))
    cond = str(rng.choice(["none", "sine", "herschel"]))
    precision = "fp64" if seed % 2 else "fp32"
    tol = 1e-11 if precision == "fp64" else 1e-5
    m = HanserPSF(wl, na, ni, res, size, zrange=z, vec_corr=vc, condition=cond, precision=precision)
    ref = orc.hanser_psfa(wl, na, ni, res, size, z, vc, cond)
    assert rel_l2(m.PSFi, orc.psfi(ref)) <= tol, (size, wl, na, ni, res, vc, cond)
    assert rel_l2(m.PSFa, ref) <= tol
    if seed % 3 == 0:                                   # a random masked pupil through apply_pupil
        _, kr, _, _, _ = orc.hanser_kspace(wl, ni, res, size)
        pupil = orc.hanser_pupil(kr, na, wl) * (1 + 0.2 * rng.standard_normal((size, size))) * np.exp(
            1j * rng.standard_normal((size, size)))
        m.apply_pupil(pupil)
        assert rel_l2(m.PSFi, orc.psfi(orc.hanser_psfa(wl, na, ni, res, size, z, vc, cond, pupil))) <= tol


@pytest.mark.parametrize("seed", range(10))
def test_phase_retrieval_random(seed):
    from oracle import pyotf_oracle as orc
    from pyotf_b200.phaseretrieval import retrieve_phase

    rng = np.random.default_rng(2000 + seed)
    size, wl, na, ni, res, z = _draw(rng)
    z = z + 11.0 if len(z) > 1 else np.array([-400.0, 150.0, 700.0])
    pupil = orc.aberrated_pupil(wl, na, ni, res, size, rng.random(10) * 0.1, rng.random(10))
    data = orc.psfi(orc.hanser_psfa(wl, na, ni, res, size, z, "none", "none", pupil)) * float(rng.uniform(1e2, 1e5))
    phase_only = bool(seed % 4 == 3)
    o = orc.retrieve_phase_loop(data, wl, na, ni, res, zrange=z, max_iters=6, pupil_tol=0, mse_tol=0, phase_only=phase_only)
    precision = "fp64" if seed % 2 else "fp32"
    r = retrieve_phase(data, dict(wl=wl, na=na, ni=ni, res=res, zrange=z), 6, 0, 0, phase_only, precision=precision)
    tol = 1e-10 if precision == "fp64" else 1e-5
    assert rel_l2(r.mse, o["mse"]) <= tol, (size, wl, na, ni, res)
    assert rel_l2(r.pupil, o["pupil"]) <= 3 * tol
