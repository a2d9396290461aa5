"""TEST INFRASTRUCTURE ONLY -- regenerates tests/golden/*.npz from the LIVE upstream pyotf.

Run in the build container (needs /root/reference):  python oracle/make_golden.py
Every array written here is an output of the *unmodified* reference (imported through
oracle/refshim.py); nothing comes from oracle/pyotf_oracle.py or from the CUDA path.
The only non-output is ``fixture_psf_520.npz``: the reference's measured 25x256x256 uint16
z-stack (fixtures/psf_wl520nm_z300nm_x130nm_na0.85_n1.0.tif) re-encoded losslessly, because
BASELINE.json config 2 is defined on it and the GPU box has no /root/reference.
"""

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def save(name, **arrays):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB  keys={list(arrays)}")


def main():
    assert refshim.reference_available(), "needs the live reference"
    refshim.load_reference()
    from pyotf.otf import HanserPSF, SheppardPSF, apply_aberration
    from pyotf.phaseretrieval import retrieve_phase
    from pyotf.utils import prep_data_for_PR
    from pyotf.zernike import noll2degrees, zernike

    os.makedirs(OUT, exist_ok=True)

    # ---- 1. small Hanser cases, every vec_corr x condition ---------------------------------
    out = {}
    kw = dict(wl=525, na=1.27, ni=1.33, res=100, size=64, zres=250, zsize=5)
    for vc in ("none", "x", "y", "z", "total"):
        for cond in ("none", "sine", "herschel"):
            m = HanserPSF(vec_corr=vc, condition=cond, **kw)
            out[f"psfi_{vc}_{cond}"] = m.PSFi
            if vc in ("none", "y"):
                out[f"psfa_{vc}_{cond}"] = m.PSFa
    m = HanserPSF(**kw)
    out["otfi_none_sine"] = m.OTFi
    out["otfa_none_sine"] = m.OTFa
    out["params"] = np.array([kw[k] for k in ("wl", "na", "ni", "res", "size", "zres", "zsize")], float)
    save("hanser_small.npz", **out)

    # ---- 2. config 1 (HanserPSF 256^2 x 128) ------------------------------------------------
    m = HanserPSF(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=256, zsize=128)
    p, o = m.PSFi, m.OTFi
    save(
        "cfg1_hanser256.npz",
        psfi_plane_sums=p.sum((1, 2)),
        psfi_planes=p[[0, 37, 64, 100, 127]],
        psfi_plane_idx=np.array([0, 37, 64, 100, 127]),
        psfi_center_line=p[:, 128, 128],
        psfi_sum=p.sum(),
        psfi_l2=np.linalg.norm(p.ravel()),
        otfi_planes=o[[64, 70]],
        otfi_plane_idx=np.array([64, 70]),
        otfi_zx=o[:, 128, :],
        otfi_l2=np.linalg.norm(o.ravel()),
        zrange=m.zrange,
    )

    # ---- 3. fixture stack + config 2 (phase retrieval) -------------------------------------
    raw = refshim.read_fixture_stack()
    assert raw.shape == (25, 256, 256) and raw.dtype == np.uint16
    save("fixture_psf_520.npz", raw=raw)
    data = prep_data_for_PR(raw, 256, 1.1)
    params = dict(wl=520, na=0.85, ni=1.0, res=130, zres=300)
    t = time.time()
    r1 = retrieve_phase(data, dict(params), 100, 1e-4, 1e-4)
    t1 = time.time() - t
    t = time.time()
    r2 = retrieve_phase(data, dict(params), 100, 0, 0)
    t2 = time.time() - t
    print(f"cfg2: early-stop {len(r1.mse)} iters {t1:.1f}s; full {len(r2.mse)} iters {t2:.1f}s")
    # single-step pins: pupil after 1, 2 and 10 updates (tol=0 runs, max_iters = k+... see below)
    steps = {}
    for k in (1, 2, 10):
        # max_iters=k with tol 0 performs k updates; complex pupil = mag*exp(i*phase) (unwrap
        # stub is identity so phase = angle)
        rk = retrieve_phase(data, dict(params), k, 0, 0)
        steps[f"pupil_after_{k}"] = np.fft.ifftshift(rk.mag * np.exp(1j * rk.phase))
    save(
        "cfg2_phase_retrieval.npz",
        data_sum=data.sum(),
        data_max=data.max(),
        es_mse=r1.mse,
        es_mse_diff=r1.mse_diff,
        es_pupil_diff=r1.pupil_diff,
        es_mag=r1.mag,
        es_angle=r1.phase,
        es_model_psfi_plane13=r1.model.PSFi[13],
        full_mse=r2.mse,
        full_mse_diff=r2.mse_diff,
        full_pupil_diff=r2.pupil_diff,
        full_mag=r2.mag,
        full_angle=r2.phase,
        ref_seconds=np.array([t1, t2]),
        **steps,
    )

    # ---- 4. integration-test twin (tests/integration_test.py:25-63), 400 iterations ---------
    np.random.seed(12345)
    mk = dict(wl=525, na=1.27, ni=1.33, res=100, size=128,
              zrange=[-1000, -500, 0, 250, 1000, 3000], vec_corr="none", condition="none")
    model = HanserPSF(**mk)
    model._gen_kr()
    r = model._kr * model.wl / model.na
    zerns = zernike(r, model._phi, *noll2degrees(np.arange(5, 16)))
    pcoefs = np.random.rand(zerns.shape[0])
    mcoefs = np.random.rand(zerns.shape[0])
    pupil_phase = (zerns * pcoefs[:, None, None]).sum(0)
    pupil_mag = (zerns * mcoefs[:, None, None]).sum(0)
    pupil_mag = pupil_mag + model._gen_pupil() * (2.0 - pupil_mag.min())
    model.apply_pupil(pupil_mag * np.exp(1j * pupil_phase) * model._gen_pupil())
    psfi = model.PSFi
    res = retrieve_phase(psfi, dict(mk), max_iters=400, pupil_tol=0, mse_tol=0)
    res60 = retrieve_phase(psfi, dict(mk), max_iters=60, pupil_tol=0, mse_tol=0)
    save(
        "integration_twin.npz",
        pcoefs=pcoefs, mcoefs=mcoefs, pupil_mag=pupil_mag.real, pupil_phase=pupil_phase,
        psfi=psfi, mse400=res.mse, pupil_diff400=res.pupil_diff, mag400=res.mag,
        angle400=res.phase, mse60=res60.mse, pupil_diff60=res60.pupil_diff, mse_diff60=res60.mse_diff,
        mag60=res60.mag, angle60=res60.phase,
    )

    # ---- 5. Zernike modes, Noll 1..120 ------------------------------------------------------
    x = np.linspace(-1.1, 1.1, 41)
    xx, yy = np.meshgrid(x, x)
    rr, th = np.hypot(yy, xx), np.arctan2(yy, xx)
    n, mm = noll2degrees(np.arange(1, 121))
    save("zernike_noll120.npz", r=rr, theta=th, n=n, m=mm, zern=zernike(rr, th, n, mm),
         zern_unnormed=zernike(rr, th, n, mm, norm=False)[:36])

    # ---- 6. config 5 unit: one library entry (120 Noll phase coefs, 256^2 x 64) -------------
    base = HanserPSF(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=256, zsize=64)
    coefs = np.random.default_rng(0).standard_normal((8192, 120)) * 0.1
    idx = np.random.default_rng(1).choice(8192, 32, replace=False)
    lib = {}
    one = np.random.default_rng(0).standard_normal(120) * 0.1
    mdl = apply_aberration(base, None, one)
    lib["unit_psfi_plane32"] = mdl.PSFi[32]
    lib["unit_psfi_plane_sums"] = mdl.PSFi.sum((1, 2))
    lib["unit_psfi_sum"] = mdl.PSFi.sum()
    sums, centers, p32 = [], [], []
    for j in idx[:8]:
        mdl = apply_aberration(base, None, coefs[j])
        sums.append(mdl.PSFi.sum((1, 2)))
        centers.append(mdl.PSFi[:, 120:136, 120:136])
        p32.append(mdl.PSFi[32])
    save("cfg5_library.npz", sample_idx=idx, entry_plane_sums=np.array(sums),
         entry_center_crops=np.array(centers), entry_plane32=np.array(p32[:2]), **lib)

    # ---- 7. small aberration case incl. magnitude coefs --------------------------------------
    rng = np.random.default_rng(7)
    pc, mc = rng.standard_normal(37) * 0.2, rng.standard_normal(37) * 0.05
    b = HanserPSF(wl=520, na=0.85, ni=1.0, res=130, size=64, zres=300, zsize=4)
    save("aberration_small.npz", pcoefs=pc, mcoefs=mc,
         psfi_both=apply_aberration(b, mc, pc).PSFi, psfi_phase=apply_aberration(b, None, pc).PSFi,
         psfi_mag=apply_aberration(b, mc, None).PSFi)

    # ---- 8. Sheppard: small exhaustive + config 3 (zres=190: 200 raises ValueError) ---------
    sk = dict(wl=520, na=1.27, ni=1.33, res=100, size=32, zres=190, zsize=24)
    out = {}
    for vc, cond, dual in (("none", "sine", False), ("total", "sine", False),
                           ("x", "herschel", False), ("z", "none", False), ("none", "none", True)):
        s = SheppardPSF(vec_corr=vc, condition=cond, dual=dual, **sk)
        tag = f"{vc}_{cond}_{int(dual)}"
        out[f"psfi_{tag}"] = s.PSFi
        out[f"nshell_{tag}"] = np.array(s.valid_points.sum())
        if vc == "none":
            out[f"psfa_{tag}"] = s.PSFa
            out[f"otfa_{tag}"] = s.OTFa
    save("sheppard_small.npz", **out)
    s = SheppardPSF(wl=520, na=1.27, ni=1.33, res=100, zres=190, size=256, zsize=256)
    p = s.PSFi
    save("cfg3_sheppard256.npz", nshell=np.array(s.valid_points.sum()), dk=s.dk, dkz=s.dkz,
         psfi_plane_sums=p.sum((1, 2)), psfi_planes=p[[0, 100, 128, 129, 200]],
         psfi_plane_idx=np.array([0, 100, 128, 129, 200]), psfi_l2=np.linalg.norm(p.ravel()),
         psfi_zx=p[:, 128, :], otfi_zx=s.OTFi[:, 128, :])

    # ---- 9. config 4 sample: 1024^2 + 15 Zernikes, 4 planes of the 1024-plane stack ---------
    big = HanserPSF(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=1024, zsize=1024)
    zr = big.zrange
    pc15 = np.random.default_rng(0).standard_normal(15) * 0.2
    sel = np.array([0, 511, 512, 1023])
    chunk = HanserPSF(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=1024, zrange=zr[sel])
    mdl = apply_aberration(chunk, None, pc15)
    p = mdl.PSFi
    save("cfg4_hanser1024.npz", pcoefs=pc15, plane_idx=sel, zvals=zr[sel],
         psfi_plane_sums=p.sum((1, 2)), psfi_l2=np.sqrt((p**2).sum((1, 2))),
         psfi_center_crops=p[:, 448:576, 448:576], psfi_row512=p[:, 512, :])


if __name__ == "__main__":
    main()
