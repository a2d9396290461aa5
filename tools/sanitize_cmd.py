#!/usr/bin/env python
"""Small invocations of the hot kernels for compute-sanitizer (memcheck / racecheck / initcheck):

    compute-sanitizer --tool memcheck  python tools/sanitize_cmd.py
    compute-sanitizer --tool racecheck python tools/sanitize_cmd.py

K1: even-symmetric variant with 64-row and 128-row CTAs (device z and host z), a user pupil, a vectorial model;
K2: three phase-retrieval iterations on a small synthetic stack."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from pyotf_b200 import _engine as eng
    from pyotf_b200.otf import HanserPSF, apply_aberration
    from pyotf_b200.phaseretrieval import retrieve_phase

    h = eng.hanser_handle(520, 0.85, 1.0, 130, 256, "none", "sine", "fp32", -1)
    for nz in (3, 80):                                   # 64-row and 128-row CTAs
        z = (np.arange(nz) - nz // 2) * 300.0
        a = h.psf(eng.to_device_f64(z))[1]
        b = h.psf(z)[1]
        b2 = h.psf(z + 1.0)[1]                           # overlapping launches
        torch.cuda.synchronize()
        assert torch.equal(a, b)
        print("k1 sym", nz, float(a.sum()), float(b2.sum()))
    m = HanserPSF(520, 0.85, 1.0, 130, 128, zres=300, zsize=6, vec_corr="total", condition="herschel", precision="fp32")
    print("k1 vectorial", float(m.PSFi.sum()))
    m = HanserPSF(520, 0.85, 1.0, 130, 128, zres=300, zsize=6, precision="fp32")
    ab = apply_aberration(m, None, np.array([0, 0, 0, 0, 0.3, 0.2, -0.4, 0.1]))
    data = ab.PSFi * 1e4
    print("k1 user pupil", float(data.sum()))
    res = retrieve_phase(data, dict(wl=520, na=0.85, ni=1.0, res=130, zres=300), 3, 0, 0, precision="fp32")
    print("k2", res.mse)


if __name__ == "__main__":
    main()
