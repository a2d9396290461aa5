#!/usr/bin/env python
"""Per-phase timestamps of K1 (debug build with -DPYOTF_K1_TIMING in tools/_dbg/, see tools/README.md):
where a CTA's time goes (table / fold / columns / rows) and when the CTAs of a launch start and end."""
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
    from pyotf_b200 import _engine as eng

    lib = L.load()
    lib.pyotf_b200_debug_k1_marks.restype = C.c_int
    lib.pyotf_b200_debug_k1_marks.argtypes = [C.c_void_p, C.c_int]
    h = eng.hanser_handle(520, 0.85, 1.0, 130, 256, "none", "sine", "fp32", -1)
    for nz in [int(a) for a in sys.argv[1:]] or [128]:
        z = eng.to_device_f64((np.arange(nz) - nz // 2) * 300.0)
        out = torch.empty((1, nz, 256, 256), dtype=torch.float32, device="cuda")
        for _ in range(5):
            h.psf(z, psfi_out=out)
        torch.cuda.synchronize()
        h.psf(z, psfi_out=out)
        torch.cuda.synchronize()
        m = np.zeros(4096 * 8, dtype=np.uint64)
        rc = lib.pyotf_b200_debug_k1_marks(m.ctypes.data, m.size)
        assert rc == 0, rc
        m = m.reshape(4096, 8).astype(np.int64)
        m = m[m[:, 6] > 0]
        print(f"nz={nz}: {len(m)} CTAs recorded")
        names = ["start->prologue", "prologue->table", "table->fold", "fold->columns", "columns->rows"]
        pairs = [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5)]
        for nm, (a, b) in zip(names, pairs):
            d = m[:, b] - m[:, a]
            print(f"  {nm:18s} cycles: median {np.median(d):8.0f}  mean {d.mean():8.0f}  min {d.min():8.0f}  max {d.max():8.0f}")
        tot = m[:, 5] - m[:, 0]
        print(f"  {'total':18s} cycles: median {np.median(tot):8.0f}  mean {tot.mean():8.0f}  min {tot.min():8.0f}  max {tot.max():8.0f}")
        t0 = m[:, 6].min()
        st, en = m[:, 6] - t0, m[:, 7] - t0
        print(f"  globaltimer ns: CTA starts  min {st.min()} median {np.median(st):.0f} max {st.max()};  ends min {en.min()} median {np.median(en):.0f} max {en.max()}")
        hist, edges = np.histogram(st, bins=8)
        print("  start histogram:", list(zip(edges[:-1].astype(int).tolist(), hist.tolist())))
        hist, edges = np.histogram(en, bins=8)
        print("  end histogram:  ", list(zip(edges[:-1].astype(int).tolist(), hist.tolist())))


if __name__ == "__main__":
    main()
