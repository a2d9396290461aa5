#!/usr/bin/env python
"""K1 development check: cfg1 stack (256^2 x 128 planes -> PSFi) against the oracle, then the pipelined launch rate
through pyotf_b200_hanser_psf_zhost (the call HanserPSF.PSFi makes).  PYOTF_B200_K1_SYM_OLD=1 selects the previous
half-plane kernel for an A/B comparison.

    python tools/k1_dev.py [--precision=fp64] [--nocheck]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from pyotf_b200 import _engine as eng
    from pyotf_b200 import _lib

    prec = "fp64" if "--precision=fp64" in sys.argv else "fp32"
    dt = torch.float64 if prec == "fp64" else torch.float32
    tol = 1e-11 if prec == "fp64" else 1e-5
    nz = 128
    z = (np.arange(nz) - (nz + 1) // 2) * 300.0
    h = eng.hanser_handle(520, 0.85, 1.0, 130, 256, "none", "sine", prec, -1)
    _, psfi = h.psf(z)
    torch.cuda.synchronize()
    if "--nocheck" not in sys.argv:
        from oracle import pyotf_oracle as orc

        ref = orc.psfi(orc.hanser_psfa(520, 0.85, 1.0, 130, 256, z))
        got = psfi[0].cpu().numpy()
        err = np.linalg.norm((got - ref).ravel()) / np.linalg.norm(ref.ravel())
        worst = max(np.linalg.norm(got[i] - ref[i]) / np.linalg.norm(ref[i]) for i in range(nz))
        print(f"{prec}: rel L2 vs oracle {err:.3e} (worst plane {worst:.3e}), tol {tol:g} -> {'OK' if err <= tol else 'FAIL'}",
              flush=True)
        # odd plane counts / offsets from focus
        z2 = np.array([-1234.5, 0.0, 17.0, 5000.0, 20000.0])
        _, p2 = h.psf(z2)
        ref2 = orc.psfi(orc.hanser_psfa(520, 0.85, 1.0, 130, 256, z2))
        err2 = np.linalg.norm((p2[0].cpu().numpy() - ref2).ravel()) / np.linalg.norm(ref2.ravel())
        print(f"{prec}: 5 odd planes rel L2 {err2:.3e} -> {'OK' if err2 <= tol else 'FAIL'}", flush=True)
    lib = _lib.load()
    outs = [torch.empty((1, nz, 256, 256), dtype=dt, device="cuda") for _ in range(8)]
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    zc = np.ascontiguousarray(z)
    calls = [(h._h, C.c_void_p(zc.ctypes.data), nz, C.c_void_p(0), 1, 0, C.c_void_p(0), C.c_void_p(o.data_ptr()), sp) for o in outs]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for steps in (1, 50):
        for i in range(5):
            lib.pyotf_b200_hanser_psf_zhost(*calls[i % 8])
        best = 1e9
        for rep in range(20):
            torch.cuda.synchronize()
            e0.record()
            for i in range(steps):
                lib.pyotf_b200_hanser_psf_zhost(*calls[i % 8])
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / steps)
        es = 4 if prec == "fp32" else 8
        print(f"{prec}: {steps:3d} back-to-back launches: {1e3 * best:7.2f} us per 128-plane stack = "
              f"{nz / best / 1e3:6.2f} Mplanes/s = {nz * 65536 * es / best / 1e6:7.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
