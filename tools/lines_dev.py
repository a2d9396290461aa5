#!/usr/bin/env python
"""Timing of the paths built on the line-FFT passes: SheppardPSF cfg3, HanserPSF 1024^2 (cfg4, 64-plane chunk), OTFi cfg1.
PYOTF_B200_NO_REG_LINES=1 selects the staged shared-memory kernel for the strided passes (A/B)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench
    from pyotf_b200 import _engine as eng
    from pyotf_b200 import _fft
    from pyotf_b200.otf import HanserPSF, SheppardPSF, _aberrated_pupils

    prec = "fp64" if "--precision=fp64" in sys.argv else "fp32"
    rt = torch.float32 if prec == "fp32" else torch.float64
    es = 4 if prec == "fp32" else 8
    tag = "staged" if os.environ.get("PYOTF_B200_NO_REG_LINES") else "register"
    sh = SheppardPSF(wl=520, na=1.27, ni=1.33, res=100, zres=190, size=256, zsize=256, precision=prec)

    def shep():
        sh._attribute_changed()
        return sh.PSFi_dev

    ms = bench._time_gpu(shep, reps=6)
    print(f"[{tag} {prec}] sheppard cfg3: {1e3 * ms:8.1f} us / volume  ({256 ** 3 * es / ms / 1e6:7.1f} GB/s)", flush=True)
    big = HanserPSF(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=1024, zsize=1024, precision=prec)
    zc = eng.to_device_f64(big.zrange[:64])
    pup = _aberrated_pupils(big, None, (np.random.default_rng(0).standard_normal(15) * 0.2)[None])
    hb = pup.handle(big)
    outb = [torch.empty((1, 64, 1024, 1024), dtype=rt, device="cuda") for _ in range(2)]
    st = {"i": 0}

    def bigstep():
        st["i"] += 1
        hb.psf(zc, pup.tensor, want_psfa=False, want_psfi=True, psfi_out=outb[st["i"] % 2])

    ms = bench._time_gpu(bigstep, reps=6)
    print(f"[{tag} {prec}] hanser 1024^2: {1e3 * ms / 64:8.2f} us / plane  ({64 / ms / 1e3 * 1e3:9.0f} planes/s, {64 * 1024 * 1024 * es / ms / 1e6:7.1f} GB/s)", flush=True)
    hs, _ = big._handle_and_pupil(None)

    def symstep():
        st["i"] += 1
        hs.psf(zc, None, want_psfa=False, want_psfi=True, psfi_out=outb[st["i"] % 2])

    ms = bench._time_gpu(symstep, reps=6)
    print(f"[{tag} {prec}] hanser 1024^2 default model: {1e3 * ms / 64:8.2f} us / plane  ({64 / ms * 1e3:9.0f} planes/s, {64 * 1024 * 1024 * es / ms / 1e6:7.1f} GB/s)", flush=True)
    del outb
    m = HanserPSF(precision=prec, **bench.CFG1)
    psfi = m.PSFi_dev
    ms = bench._time_gpu(lambda: _fft.shifted_fft_dev(psfi, axes=None, inverse=False), reps=6)
    print(f"[{tag} {prec}] OTFi cfg1:     {1e3 * ms:8.1f} us  ({128 * 65536 * 3 * es / ms / 1e6:7.1f} GB/s)", flush=True)


if __name__ == "__main__":
    main()
