#!/usr/bin/env python
"""Short driver for ncu captures: a few launches of each hot kernel on the BASELINE shapes.

    python tools/prof_cmd.py [k1] [pr] [otf] [sheppard] [--precision fp32|fp64]

k1: HanserPSF cfg1 (256^2 x 128 planes -> PSFi), 4 launches.   pr: retrieve_phase loop on the fixture
stack, 6 iterations.  Numbers printed by this script under ncu are NOT bench values.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from pyotf_b200 import _engine as eng

    what = [a for a in sys.argv[1:] if not a.startswith("--")] or ["k1", "pr"]
    precision = "fp64" if "--precision=fp64" in sys.argv else "fp32"
    torch.cuda.set_device(0)
    if "k1" in what:
        h = eng.hanser_handle(520, 0.85, 1.0, 130, 256, "none", "sine", precision, -1)
        z = (np.arange(128) - 64) * 300.0          # host zrange: the call HanserPSF.PSFi makes
        outs = [None] * 4
        for i in range(4):
            _, outs[i] = h.psf(z)
        psfi = outs[-1]
        torch.cuda.synchronize()
        print("k1 ok", float(psfi.sum()))
    if "pr" in what:
        sys.path.insert(0, ROOT)
        import bench

        data = bench.fixture_stack()
        dt = np.float32 if precision == "fp32" else np.float64
        zall = (np.arange(25) - 13) * 300.0
        st = eng.PRState(520, 0.85, 1.0, 130, 256, zall, torch.as_tensor(data.astype(dt), device="cuda"), precision)
        mse, _, _ = st.run(6, 0.0, 0.0)
        print("pr ok", mse[-1])
    if "lib" in what:
        from pyotf_b200.otf import HanserPSF, _aberrated_pupils

        base = HanserPSF(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=256, zsize=64, precision=precision)
        coefs = np.random.default_rng(0).standard_normal((32, 120)) * 0.1
        z = eng.to_device_f64(base.zrange)
        for _ in range(2):
            pupils = _aberrated_pupils(base, None, coefs)
            _, psfi = pupils.handle(base).psf(z, pupils.tensor, want_psfa=False, want_psfi=True)
        torch.cuda.synchronize()
        print("lib ok", float(psfi.sum()))
    if "big" in what:
        from pyotf_b200.otf import HanserPSF, _aberrated_pupils

        big = HanserPSF(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=1024, zsize=1024, precision=precision)
        pup = _aberrated_pupils(big, None, (np.random.default_rng(0).standard_normal(15) * 0.2)[None])
        zc = eng.to_device_f64(big.zrange[:26])
        for _ in range(2):
            _, psfi = pup.handle(big).psf(zc, pup.tensor)
        torch.cuda.synchronize()
        print("big ok", float(psfi.sum()))
    if "bigsym" in what:
        from pyotf_b200.otf import HanserPSF

        big = HanserPSF(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=1024, zsize=1024, precision=precision)
        hs, _ = big._handle_and_pupil(None)
        zc = eng.to_device_f64(big.zrange[:26])
        for _ in range(2):
            _, psfi = hs.psf(zc, None)
        torch.cuda.synchronize()
        print("bigsym ok", float(psfi.sum()))
    if "otf" in what:
        from pyotf_b200.otf import HanserPSF

        m = HanserPSF(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=256, zsize=128, precision=precision)
        o = m.OTFi_dev
        torch.cuda.synchronize()
        print("otf ok", float(o.abs().max()))
    if "sheppard" in what:
        from pyotf_b200.otf import SheppardPSF

        m = SheppardPSF(wl=520, na=1.27, ni=1.33, res=100, zres=190, size=256, zsize=256, precision=precision)
        p = m.PSFi_dev
        torch.cuda.synchronize()
        print("sheppard ok", float(p.sum()))


if __name__ == "__main__":
    main()
