#!/usr/bin/env python
"""Time the phase-retrieval iteration on the fixture stack for several launch shapes (tuning aid).
rows_per_cta: 0 = auto, 64 / 32 = narrow (256-thread) CTAs, -64 = wide (512-thread) CTAs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench
    from pyotf_b200 import _engine as eng

    data = bench.fixture_stack()
    zall = (np.arange(25) - 13) * 300.0
    for nrep in (1, 4):                       # 25 planes and a 100-plane stack (the fixture tiled)
        d = np.concatenate([data] * nrep)
        z = np.concatenate([zall + 7500.0 * k for k in range(nrep)])
        for rows in (0, 64, 32, -64):
            st = eng.PRState(520, 0.85, 1.0, 130, 256, z, torch.as_tensor(d.astype(np.float32), device="cuda"), "fp32",
                             rows_per_cta=rows)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e9
            for rep in range(5):
                st.set_pupil(None)
                torch.cuda.synchronize()
                e0.record()
                mse, _, _ = st.run(100, 0.0, 0.0)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(f"nz={25 * nrep:4d} rows_per_cta={rows:4d} -> M={st.rows_per_cta} parts={st.nparts}: "
                  f"{10 * best:7.2f} us/iter  mse[-1]={mse[-1]:.6g}", flush=True)


if __name__ == "__main__":
    main()
