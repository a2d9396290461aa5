"""BASELINE.json's FULL sizes on the GPU, checked through size-independent properties (the oracle cannot
run these in seconds): Parseval per plane, run-splitting of the phase-retrieval loop, shard invariance."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _plane_energy(handle, pupilc):
    """sum |pupil * amp|^2 / N^2 per batch entry: what every plane of PSFi must sum to (Parseval; the
    defocus factor has unit modulus)."""
    import torch

    _, amp, _ = handle.tables()
    p = pupilc[:, :handle.K].to(torch.complex128)
    return ((p.abs() ** 2) * (amp[0].to(torch.float64) ** 2)).sum((1, 2)) / float(handle.size) ** 2


def test_cfg5_full_library_parseval():
    """All 8192 entries x 64 planes x 256^2 (137 GB of fp32 PSFi, streamed 128 entries at a time): every
    plane of every entry carries the energy of its aberrated pupil."""
    import torch

    from pyotf_b200 import _engine as eng
    from pyotf_b200.otf import HanserPSF, _aberrated_pupils

    base = HanserPSF(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=256, zsize=64, precision="fp32")
    coefs = np.random.default_rng(0).standard_normal((8192, 120)) * 0.1
    z = eng.to_device_f64(base.zrange)
    B = 128
    out = torch.empty((B, 64, 256, 256), dtype=torch.float32, device="cuda")
    worst = 0.0
    for i in range(0, 8192, B):
        pupils = _aberrated_pupils(base, None, coefs[i:i + B])
        h = pupils.handle(base)
        h.psf(z, pupils.tensor, want_psfa=False, want_psfi=True, psfi_out=out)
        sums = out.sum((2, 3), dtype=torch.float64)                    # [B, 64]
        want = _plane_energy(h, pupils.tensor)[:, None]
        worst = max(worst, float(((sums - want).abs() / want).max()))
        assert torch.isfinite(out[::16, ::8]).all()
    assert worst <= 2e-6, worst
    # phase-only aberrations do not change |pupil|: every entry has the ideal pupil's energy
    assert abs(float(want[0, 0]) / (9289 / 65536 * float((_plane_amp2(base))) ) - 1) <= 1e-6


def _plane_amp2(base):
    """mean of amp^2 over the ideal pupil's 9289 pixels for the sine condition (host, fp64)."""
    from oracle import pyotf_oracle as orc

    _, kr, _, kmag, _ = orc.hanser_kspace(base.wl, base.ni, base.res, base.size)
    mask = kr <= base.na / base.wl
    theta = np.arcsin((kr < kmag) * kr / kmag)
    return (1.0 / np.cos(theta))[mask].mean()


def test_cfg4_full_stack_parseval_and_shard_invariance():
    """1024 planes of 1024^2 with 15 Zernike phase terms: Parseval on every plane; a z-shard computed on its
    own equals the same planes of a longer chunk bit for bit (what the 8-GPU z-sharding relies on)."""
    import torch

    from pyotf_b200 import _engine as eng
    from pyotf_b200.otf import HanserPSF, _aberrated_pupils

    big = HanserPSF(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=1024, zsize=1024, precision="fp32")
    pup = _aberrated_pupils(big, None, (np.random.default_rng(0).standard_normal(15) * 0.2)[None])
    h = pup.handle(big)
    want = float(_plane_energy(h, pup.tensor)[0])
    out = torch.empty((1, 64, 1024, 1024), dtype=torch.float32, device="cuda")
    worst = 0.0
    keep = None
    for c in range(0, 1024, 64):
        h.psf(eng.to_device_f64(big.zrange[c:c + 64]), pup.tensor, want_psfa=False, want_psfi=True, psfi_out=out)
        sums = out[0].sum((1, 2), dtype=torch.float64)
        worst = max(worst, float(((sums - want).abs() / want).max()))
        if c == 128:
            keep = out[0, 16:24].clone()
    assert worst <= 2e-6, worst
    _, part = h.psf(eng.to_device_f64(big.zrange[144:152]), pup.tensor)
    assert torch.equal(part[0], keep)


def test_cfg2_run_splitting_and_cfg1_parseval():
    """Phase retrieval on the fixture: 60 + 40 iterations continue to exactly the pupil of 100 in one go
    (same kernels, same order), and the mse history of the first leg is a prefix of the long run."""
    import torch

    from pyotf_b200 import _engine as eng
    from pyotf_b200.otf import HanserPSF
    from pyotf_b200.utils import prep_data_for_PR

    data = prep_data_for_PR(np.load(os.path.join(GOLDEN, "fixture_psf_520.npz"))["raw"], 256, 1.1)
    z = (np.arange(25) - 13) * 300.0
    for precision in ("fp32", "fp64"):
        dt = np.float32 if precision == "fp32" else np.float64
        d = torch.as_tensor(data.astype(dt), device="cuda")
        a = eng.PRState(520, 0.85, 1.0, 130, 256, z, d, precision)
        mse_a, _, _ = a.run(100, 0, 0)
        b = eng.PRState(520, 0.85, 1.0, 130, 256, z, d, precision)
        mse_b1, _, _ = b.run(60, 0, 0)
        mse_b2, _, _ = b.run(40, 0, 0)
        assert np.array_equal(mse_b1, mse_a[:60]) and np.array_equal(mse_b2, mse_a[60:])
        assert torch.equal(a.get_pupil(), b.get_pupil())
        assert torch.equal(a.get_pupil(applied=True), b.get_pupil(applied=True))
    # cfg1: every plane of the ideal 128-plane stack carries the pupil's energy
    m = HanserPSF(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=256, zsize=128, precision="fp32")
    sums = m.PSFi_dev.sum((1, 2), dtype=torch.float64)
    want = 9289 / 65536 * _plane_amp2(m)
    assert float(((sums - want).abs() / want).max()) <= 2e-6


def test_cfg3_sheppard_parseval():
    """256^3 shell: sum |PSFa|^2 = (number of shell voxels) / (N^2 NZ) for the scalar model."""
    import torch

    from pyotf_b200.otf import SheppardPSF

    s = SheppardPSF(wl=520, na=1.27, ni=1.33, res=100, zres=190, size=256, zsize=256, precision="fp32")
    total = float(s.PSFi_dev.sum(dtype=torch.float64))
    assert abs(total / (56986 / 256.0 ** 3) - 1) <= 2e-6
