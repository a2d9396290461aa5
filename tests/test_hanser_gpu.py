"""K1 parity (GPU): fused pupil -> defocus -> 2D IFFT -> |.|^2 against the CPU oracle and the
golden vectors generated from the live reference.  All calls go through the C ABI."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-5, "fp64": 1e-11}


def _run(precision, wl, na, ni, res, size, zrange, vec_corr="none", condition="sine", pupil=None,
         want_psfa=True):
    import torch

    from pyotf_b200 import _engine as eng

    support = -1
    pupilc = None
    if pupil is not None:
        pd = eng.to_device_c128(pupil)
        support = eng.pupil_support(pd)
    h = eng.hanser_handle(wl, na, ni, res, size, vec_corr, condition, precision, support)
    if pupil is not None:
        pupilc = h.pack_pupil(pd[None])
    z = eng.to_device_f64(np.atleast_1d(zrange))
    psfa, psfi = h.psf(z, pupilc, want_psfa=want_psfa, want_psfi=True)
    torch.cuda.synchronize()
    return (psfa[0].cpu().numpy() if psfa is not None else None), psfi[0].cpu().numpy()


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("vec_corr", ["none", "x", "y", "z", "total"])
@pytest.mark.parametrize("condition", ["none", "sine", "herschel"])
def test_small_all_modes_vs_oracle_and_golden(precision, vec_corr, condition):
    from oracle import pyotf_oracle as orc

    g = np.load(os.path.join(GOLDEN, "hanser_small.npz"))
    wl, na, ni, res, size, zres, zsize = g["params"]
    size, zsize = int(size), int(zsize)
    z = orc.default_zrange(zsize, zres)
    psfa, psfi = _run(precision, wl, na, ni, res, size, z, vec_corr, condition)
    ref_a = orc.hanser_psfa(wl, na, ni, res, size, z, vec_corr, condition)
    assert psfa.shape == ref_a.shape
    assert rel_l2(psfa, ref_a) <= TOL[precision]
    assert rel_l2(psfi, orc.psfi(ref_a)) <= TOL[precision]
    assert rel_l2(psfi, g[f"psfi_{vec_corr}_{condition}"]) <= TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("size", [32, 64, 128, 256])
def test_sizes_and_user_pupil(precision, size):
    from oracle import pyotf_oracle as orc

    rng = np.random.default_rng(size)
    wl, na, ni, res = 520, 0.85, 1.0, 130
    z = np.array([-2150.0, -300.0, 0.0, 777.7, 4100.0])
    _, kr, _, _, _ = orc.hanser_kspace(wl, ni, res, size)
    # masked random pupil (phase-retrieval like)
    pupil = orc.hanser_pupil(kr, na, wl) * (1 + 0.3 * rng.standard_normal((size, size))) * np.exp(
        1j * rng.standard_normal((size, size)))
    psfa, psfi = _run(precision, wl, na, ni, res, size, z, "none", "none", pupil)
    ref_a = orc.hanser_psfa(wl, na, ni, res, size, z, "none", "none", pupil)
    assert rel_l2(psfa, ref_a) <= TOL[precision]
    assert rel_l2(psfi, orc.psfi(ref_a)) <= TOL[precision]
    # dense random pupil (support = whole grid)
    pupil = rng.standard_normal((size, size)) + 1j * rng.standard_normal((size, size))
    psfa, psfi = _run(precision, wl, na, ni, res, size, z[:2], "none", "sine", pupil)
    ref_a = orc.hanser_psfa(wl, na, ni, res, size, z[:2], "none", "sine", pupil)
    assert rel_l2(psfa, ref_a) <= TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_cfg1_golden(precision):
    """BASELINE config 1: HanserPSF(520, 0.85, 1.0, 130, 256, zres=300, zsize=128) -> PSFi."""
    g = np.load(os.path.join(GOLDEN, "cfg1_hanser256.npz"))
    _, psfi = _run(precision, 520, 0.85, 1.0, 130, 256, g["zrange"], want_psfa=False)
    tol = TOL[precision]
    assert psfi.shape == (128, 256, 256)
    assert rel_l2(psfi.sum((1, 2)), g["psfi_plane_sums"]) <= tol
    assert rel_l2(psfi[g["psfi_plane_idx"]], g["psfi_planes"]) <= tol
    for i, p in zip(g["psfi_plane_idx"], g["psfi_planes"]):
        assert rel_l2(psfi[i], p) <= tol, f"plane {i}"
    assert rel_l2(psfi[:, 128, 128], g["psfi_center_line"]) <= tol
    assert abs(np.linalg.norm(psfi.ravel().astype(np.float64)) / g["psfi_l2"] - 1) <= tol
    # known-answer values from SURVEY.md section 8c
    assert abs(psfi.astype(np.float64).sum() / 23.75643331623896 - 1) <= tol
    assert abs(psfi[64, 128, 128] / 0.02609296492831893 - 1) <= 10 * tol
    assert (psfi >= 0).all()


def test_geometry_tables_match_oracle():
    """The fp64 edge predicate kr <= na/wl and kz must match numpy's bit for bit / to 1 ulp."""
    from oracle import pyotf_oracle as orc
    from pyotf_b200 import _engine as eng

    for (wl, na, ni, res, size) in [(520, 0.85, 1.0, 130, 256), (525, 1.27, 1.33, 100, 128),
                                    (488, 1.4, 1.518, 65, 256), (500, 0.5, 1.0, 250, 64)]:
        h = eng.hanser_handle(wl, na, ni, res, size, "none", "sine", "fp64", -1)
        kz, amp, mask = [t.cpu().numpy() for t in h.tables()]
        _, kr, _, kmag, kz_ref = orc.hanser_kspace(wl, ni, res, size)
        sel = np.arange(h.f0, h.f0 + h.K) % size
        sub = np.ix_(sel, sel)
        np.testing.assert_array_equal(mask.astype(bool), (kr <= na / wl)[sub])
        # nothing of the ideal pupil lies outside the compact box
        assert (kr <= na / wl).sum() == mask.sum()
        # kr is bit-equal to numpy.hypot (hypot_ref restates glibc's algorithm), hence so is kz
        np.testing.assert_array_equal(kz, kz_ref[sub])
        theta = np.arcsin((kr < kmag) * kr / kmag)
        a = (1 / np.sqrt(np.cos(theta))) * (kr <= na / wl)
        np.testing.assert_allclose(amp[0], a[sub], rtol=1e-14)


def test_streaming_mode_is_bit_identical():
    """pyotf_b200_hanser_set_overlap: back-to-back launches overlap (programmatic dependent launch) but every
    stack, including repeated launches into one buffer, equals the ordinary launch bit for bit."""
    import torch

    from pyotf_b200 import _engine as eng

    h = eng.hanser_handle(520, 0.85, 1.0, 130, 256, "none", "sine", "fp32", -1)
    zs = [eng.to_device_f64((np.arange(40) - 20 + 3 * k) * 300.0) for k in range(6)]
    ref = [h.psf(z)[1].clone() for z in zs]
    torch.cuda.synchronize()
    h.set_overlap(True)
    try:
        outs = [torch.empty_like(ref[0]) for _ in zs]
        same = torch.empty_like(ref[0])
        for k, z in enumerate(zs):
            h.psf(z, psfi_out=outs[k])
        for z in zs:                       # six launches into ONE buffer: the last one must win everywhere
            h.psf(z, psfi_out=same)
        torch.cuda.synchronize()
    finally:
        h.set_overlap(False)
    for k in range(6):
        assert torch.equal(outs[k], ref[k]), k
    assert torch.equal(same, ref[-1])


def test_host_zrange_entry_is_bit_identical():
    """pyotf_b200_hanser_psf_zhost (zrange on the host, as the reference holds it): short stacks travel in the kernel
    arguments and overlap back to back, long stacks / user pupils / vectorial models / other sizes are uploaded;
    every variant equals the device-z entry bit for bit, also for repeated launches into one buffer."""
    import torch

    from pyotf_b200 import _engine as eng

    h = eng.hanser_handle(520, 0.85, 1.0, 130, 256, "none", "sine", "fp32", -1)
    zs = [(np.arange(nz) - nz // 2 + 3 * k) * 150.0 + 7.0 for k, nz in enumerate((1, 40, 128, 256, 257, 300))]
    ref = [h.psf(eng.to_device_f64(z))[1].clone() for z in zs]
    torch.cuda.synchronize()
    got = [h.psf(z)[1] for z in zs]                       # numpy in -> zhost entry
    same = torch.empty_like(ref[2])
    for k in range(6):                                    # six overlapping launches into ONE buffer
        h.psf(zs[2] + 10.0 * k, psfi_out=same)
    last = h.psf(eng.to_device_f64(zs[2] + 50.0))[1]
    torch.cuda.synchronize()
    for k in range(len(zs)):
        assert torch.equal(got[k], ref[k]), k
    assert torch.equal(same, last)
    # fp64, vectorial model, PSFa kept
    h6 = eng.hanser_handle(520, 0.85, 1.0, 130, 64, "total", "herschel", "fp64", -1)
    z = np.array([-500.0, 0.0, 125.0])
    a0, i0 = h6.psf(eng.to_device_f64(z), want_psfa=True)
    a1, i1 = h6.psf(z, want_psfa=True)
    assert torch.equal(a0, a1) and torch.equal(i0, i1)
    # user pupil (precomputed-table path) and a size outside the fused kernel's set (two-pass path)
    rng = np.random.default_rng(3)
    for size in (128, 96):
        pupil = np.zeros((size, size), complex)
        f = np.fft.fftfreq(size, 1.0 / size)
        m = np.hypot(*np.meshgrid(f, f)) <= 20
        pupil[m] = np.exp(1j * rng.standard_normal(int(m.sum())))
        pd = eng.to_device_c128(pupil)
        hp = eng.hanser_handle(520, 0.85, 1.0, 130, size, "none", "sine", "fp32", eng.pupil_support(pd))
        pc = hp.pack_pupil(pd[None])
        z = np.linspace(-900, 900, 7)
        i0 = hp.psf(eng.to_device_f64(z), pc)[1]
        i1 = hp.psf(z, pc)[1]
        assert torch.equal(i0, i1), size


def test_rows128_variant_matches_rows64_and_oracle():
    """Stacks long enough that the 64-row grid would not fit one wave (4 nz > 2 x SMs) run the 128-row CTAs of the
    even-symmetric variant, also with PSFa kept and in streaming mode: same planes as the short (64-row) stacks
    within fp32 rounding (different column-FFT factorisation), and within 1e-5 of the oracle."""
    import torch

    from oracle import pyotf_oracle as orc
    from pyotf_b200 import _engine as eng

    wl, na, ni, res, size = 520, 0.85, 1.0, 130, 256
    h = eng.hanser_handle(wl, na, ni, res, size, "none", "sine", "fp32", -1)
    nz = 160
    zr = (np.arange(nz) - 77) * 125.0 + 11.0
    psfa_long, psfi_long = h.psf(eng.to_device_f64(zr), want_psfa=True, want_psfi=True)
    short_a, short_i = [], []
    for k in range(0, nz, 40):
        a, i = h.psf(eng.to_device_f64(zr[k:k + 40]), want_psfa=True, want_psfi=True)
        short_a.append(a[0].clone()); short_i.append(i[0].clone())
    h.set_overlap(True)
    try:
        stream_a, stream_i = h.psf(eng.to_device_f64(zr), want_psfa=True, want_psfi=True)
        stream_i = stream_i.clone()
    finally:
        h.set_overlap(False)
    torch.cuda.synchronize()
    assert torch.equal(stream_i, psfi_long)
    # PSFi alone goes through the whole-plane kernel (hanser_sym.cuh): same planes within fp32 rounding
    whole_i = h.psf(eng.to_device_f64(zr))[1]
    assert rel_l2(whole_i[0].cpu().numpy(), psfi_long[0].cpu().numpy()) <= 2e-6
    psfi_long, psfa_long = psfi_long[0].cpu().numpy(), psfa_long[0].cpu().numpy()
    psfi_short = torch.cat(short_i).cpu().numpy()
    psfa_short = torch.cat(short_a, dim=1).cpu().numpy()
    assert rel_l2(psfi_long, psfi_short) <= 2e-6
    assert rel_l2(psfa_long, psfa_short) <= 2e-6
    # even symmetry of the stored planes (mirror stores of the rows y > N/2)
    np.testing.assert_array_equal(psfi_long[:, 1:, :], psfi_long[:, :0:-1, :])
    sel = [0, 1, 77, 78, 121, 159]
    ref = orc.hanser_psfa(wl, na, ni, res, size, zr[sel])
    assert rel_l2(psfa_long[:, sel], ref) <= 1e-5
    assert rel_l2(psfi_long[sel], orc.psfi(ref)) <= 1e-5
    for j, k in enumerate(sel):                       # worst single plane
        assert rel_l2(psfi_long[k], orc.psfi(ref)[j]) <= 1e-5


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_wide_support_high_na(precision):
    """High-NA optics at coarse sampling: the pupil box (K = 149 at 256^2) is too large for the staged defocus
    table, so K1 / K2 evaluate the defocus on the fly and pick 32 rows per CTA; same parity bar."""
    from oracle import pyotf_oracle as orc
    from pyotf_b200 import _engine as eng
    from pyotf_b200.phaseretrieval import retrieve_phase

    wl, na, ni, res, size = 488, 1.4, 1.518, 100, 256
    z = np.array([-800.0, -150.0, 90.0, 610.0])
    h = eng.hanser_handle(wl, na, ni, res, size, "none", "sine", precision, -1)
    assert h.K >= 147
    for vc, cond in (("none", "sine"), ("total", "herschel")):
        psfa, psfi = _run(precision, wl, na, ni, res, size, z, vc, cond)
        ref = orc.hanser_psfa(wl, na, ni, res, size, z, vc, cond)
        assert rel_l2(psfa, ref) <= TOL[precision]
        assert rel_l2(psfi, orc.psfi(ref)) <= TOL[precision]
    rng = np.random.default_rng(1)
    pupil = orc.aberrated_pupil(wl, na, ni, res, size, rng.random(12) * 0.1, rng.random(12))
    data = orc.psfi(orc.hanser_psfa(wl, na, ni, res, size, z + 33.0, "none", "none", pupil)) * 1e4
    o = orc.retrieve_phase_loop(data, wl, na, ni, res, zrange=z + 33.0, max_iters=6, pupil_tol=0, mse_tol=0)
    r = retrieve_phase(data, dict(wl=wl, na=na, ni=ni, res=res, zrange=z + 33.0), 6, 0, 0, precision=precision)
    assert rel_l2(r.mse, o["mse"]) <= TOL[precision]
    assert rel_l2(r.pupil, o["pupil"]) <= (1e-11 if precision == "fp64" else 2e-5)


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_whole_plane_kernel_cfg1_and_overlapping_launches(precision):
    """K1s (hanser_sym.cuh: one CTA per plane, even / transpose symmetry, PSFi of the default model at 256^2):
    the full BASELINE config-1 stack against the oracle, exact mirror symmetry of the stored planes, and
    back-to-back launches -- into rotating buffers (they overlap: each only has to finish after its predecessor)
    and into one buffer (each waits before its first store) -- bit-identical to serialised launches."""
    import torch

    from oracle import pyotf_oracle as orc
    from pyotf_b200 import _engine as eng

    wl, na, ni, res, size, nz = 520, 0.85, 1.0, 130, 256, 128
    z = orc.default_zrange(nz, 300)
    h = eng.hanser_handle(wl, na, ni, res, size, "none", "sine", precision, -1)
    got = h.psf(z)[1]
    torch.cuda.synchronize()
    p = got[0].cpu().numpy()
    ref = orc.psfi(orc.hanser_psfa(wl, na, ni, res, size, z))
    assert rel_l2(p, ref) <= TOL[precision]
    assert max(rel_l2(p[i], ref[i]) for i in range(nz)) <= TOL[precision]
    np.testing.assert_array_equal(p[:, 1:, :], p[:, :0:-1, :])          # y <-> -y
    np.testing.assert_array_equal(p[:, :, 1:], p[:, :, :0:-1])          # x <-> -x
    # row 0 (y = -128) is stored from the transposed column x = -128
    np.testing.assert_array_equal(p[:, 0, :], p[:, :, 0])
    # different stacks back to back: rotating outputs (no wait before the stores) and a single output buffer
    stacks = [z + 37.0 * k for k in range(12)]
    serial = []
    for zz in stacks:
        serial.append(h.psf(zz)[1].clone())
        torch.cuda.synchronize()
    ring = [torch.empty_like(got) for _ in range(5)]
    keep = []
    for k, zz in enumerate(stacks):
        h.psf(zz, psfi_out=ring[k % 5])
        if k % 5 == 4 or k == len(stacks) - 1:
            torch.cuda.synchronize()
            for j in range(k - k % 5, k + 1):
                keep.append(ring[j % 5].clone())
    for a, b in zip(keep, serial):
        assert torch.equal(a, b)
    one = torch.empty_like(got)
    for zz in stacks:
        h.psf(zz, psfi_out=one)
    torch.cuda.synchronize()
    assert torch.equal(one, serial[-1])
    # odd plane counts, planes far from focus, a device-resident z (the launch then waits for its predecessor first)
    z2 = np.array([-20000.0, -1234.5, 0.0, 17.0, 5000.0])
    r2 = orc.psfi(orc.hanser_psfa(wl, na, ni, res, size, z2))
    assert rel_l2(h.psf(z2)[1][0].cpu().numpy(), r2) <= TOL[precision]
    assert rel_l2(h.psf(eng.to_device_f64(z2))[1][0].cpu().numpy(), r2) <= TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("paired", [True, False])
def test_whole_plane_kernel_z_pairing(precision, paired, monkeypatch):
    """K1s can compute planes at +-z once and store them twice (PSFi(-z) == PSFi(z) for the default model; the default
    in the fp64 mode, PYOTF_B200_K1_ZSYM=1/0 forces it on / off): mixed stacks with pairs, singles, z = 0, duplicated
    positions and unsorted order, every plane against the oracle, planes of equal |z| equal to rounding, and against the
    same positions computed one plane per call (no partner: the ordinary path)."""
    import torch

    from oracle import pyotf_oracle as orc
    from pyotf_b200 import _engine as eng

    monkeypatch.setenv("PYOTF_B200_K1_ZSYM", "1" if paired else "0")
    wl, na, ni, res, size = 520, 0.85, 1.0, 130, 256
    h = eng.hanser_handle(wl, na, ni, res, size, "none", "sine", precision, -1)
    for z in (np.array([-600.0, -300.0, 0.0, 300.0, 450.0, 600.0, 300.0]),
              np.array([1500.0, -40.0, 40.0, -1500.0, 0.0, 0.0, 7.5]),
              np.array([-250.0, 250.0])):
        ref = orc.psfi(orc.hanser_psfa(wl, na, ni, res, size, z))
        p = h.psf(z)[1][0].cpu().numpy()
        for i in range(len(z)):
            assert rel_l2(p[i], ref[i]) <= TOL[precision], (z, i)
        for i in range(len(z)):
            for j in range(i + 1, len(z)):
                if abs(z[i]) == abs(z[j]):
                    assert rel_l2(p[i], p[j]) <= TOL[precision] / 10
        if paired:
            np.testing.assert_array_equal(p[0], p[list(z).index(-z[0])])      # the first pair shares one computation
        single = np.stack([h.psf(np.array([zi]))[1][0, 0].cpu().numpy() for zi in z])
        for i in range(len(z)):
            assert rel_l2(single[i], p[i]) <= TOL[precision] / 10
    torch.cuda.synchronize()
