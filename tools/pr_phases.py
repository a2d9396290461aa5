#!/usr/bin/env python
"""Per-phase timestamps of the fused phase-retrieval iteration (pr_plane_kernel + pr_fused_finalize), debug library
from `tools/build_dbg.sh pr_f32`.  globaltimer marks of thread 0 of every CTA, last iteration of a 40-iteration run:
0 entry, 1 past griddepcontrol.wait, 2 fold + inverse columns done, 3 rows done, 4 forward columns + unfold done,
5 squared-error sum stored, 6 past the grid barrier, 7 scalars known, 8 new pupil written, 9 exit."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import pyotf_b200._lib as L

    L.LIB_PATH = os.path.join(ROOT, "tools", "_dbg", "libpyotf_b200_dbg.so")
    import bench
    from pyotf_b200 import _engine as eng

    lib = L.load()
    lib.pyotf_b200_debug_pr_marks.restype = C.c_int
    lib.pyotf_b200_debug_pr_marks.argtypes = [C.c_void_p, C.c_int]
    data = bench.fixture_stack()
    zall = (np.arange(25) - 13) * 300.0
    st = eng.PRState(520, 0.85, 1.0, 130, 256, zall, torch.as_tensor(data.astype(np.float32), device="cuda"), "fp32")
    for _ in range(3):
        st.set_pupil(None)
        st.run(40, 0.0, 0.0)
    torch.cuda.synchronize()
    m = np.zeros(256 * 16, dtype=np.uint64)
    assert lib.pyotf_b200_debug_pr_marks(m.ctypes.data, m.size) == 0
    m = m.reshape(256, 16).astype(np.int64)[: st.nparts]
    t0 = m[:, 0].min()
    names = ["entry", "past pdl wait", "fold + inverse columns", "rows", "forward columns + unfold", "mse stored",
             "past grid barrier", "scalars known", "new pupil written", "exit"]
    print(f"{st.nparts} CTAs; ns relative to the first CTA's entry (min / median / max over CTAs)")
    for i, n in enumerate(names):
        if not m[:, i].any():
            continue
        d = m[:, i] - t0
        print(f"  {i} {n:28s} {d.min():8d} {int(np.median(d)):8d} {d.max():8d}")


if __name__ == "__main__":
    main()
