#!/usr/bin/env python
"""K1 timing vs stack length (tuning aid): reveals wave quantisation (148 SMs x 2 CTAs, 4 CTAs per plane)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from pyotf_b200 import _engine as eng

    prec = "fp64" if "--precision=fp64" in sys.argv else "fp32"
    dt = torch.float64 if prec == "fp64" else torch.float32
    h = eng.hanser_handle(520, 0.85, 1.0, 130, 256, "none", "sine", prec, -1)
    for nz in (37, 74, 111, 128, 148, 222, 296, 592, 1184):
        z = eng.to_device_f64((np.arange(nz) - nz // 2) * 300.0)
        outs = [torch.empty((1, nz, 256, 256), dtype=dt, device="cuda") for _ in range(max(2, 256 // nz + 1))]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(5):
            h.psf(z, psfi_out=outs[i % len(outs)])
        best = 1e9
        for rep in range(5):
            torch.cuda.synchronize()
            e0.record()
            for i in range(20):
                h.psf(z, psfi_out=outs[i % len(outs)])
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 20)
        print(f"nz={nz:5d} ctas={4 * nz:5d} waves={4 * nz / 296:5.2f}: {1e3 * best:8.2f} us  {1e3 * best / nz:6.3f} us/plane "
              f"{nz / best / 1e3:6.2f} Mplanes/s", flush=True)


if __name__ == "__main__":
    main()
