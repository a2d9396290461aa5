#!/usr/bin/env python
"""Per-phase timestamps of the whole-plane K1 (hanser_sym_kernel), debug library from tools/build_dbg.sh.
Marks (SM cycles; warp 0 and the store warp of each 9-warp group): 0 start, 1 table built, 3 column stage 1 done,
4 past its barrier, 5 column stage 2 done, 2 columns complete, 6 / 7 / 8 first row block: stage 1 done, past the
barrier, stage 2 done; 13 transform warps done; store warp: 10 first block ready, 11 first block stored, 12 done."""
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
    lib.pyotf_b200_debug_sym_marks.restype = C.c_int
    lib.pyotf_b200_debug_sym_marks.argtypes = [C.c_void_p, C.c_int]
    h = eng.hanser_handle(520, 0.85, 1.0, 130, 256, "none", "sine", "fp32", -1)
    nz = 128
    z = (np.arange(nz) - nz // 2) * 300.0
    outs = [torch.empty((1, nz, 256, 256), dtype=torch.float32, device="cuda") for _ in range(8)]
    for mode, nlaunch in (("isolated", 1), ("streaming", 12)):
        for i in range(5):
            h.psf(z, psfi_out=outs[i % 8])
        torch.cuda.synchronize()
        for i in range(nlaunch):
            h.psf(z, psfi_out=outs[i % 8])
        torch.cuda.synchronize()
        m = np.zeros(1024 * 2 * 16, dtype=np.uint64)
        assert lib.pyotf_b200_debug_sym_marks(m.ctypes.data, m.size) == 0
        m = m.reshape(1024, 2, 16).astype(np.int64)[:nz]
        print(f"--- {mode} ({nlaunch} launch(es)), {nz} CTAs")
        def show(g, name, a, b):
            d = m[:, g, b] - m[:, g, a]
            print(f"   {name:34s} median {np.median(d):8.0f}  min {d.min():8.0f}  max {d.max():8.0f}")

        for g in range(2):
            if not m[:, g, 13].any():
                continue
            print(f" group {g}")
            show(g, "table", 0, 1)
            show(g, "col stage 1", 1, 3)
            show(g, "  barrier", 3, 4)
            show(g, "col stage 2", 4, 5)
            show(g, "  barrier", 5, 2)
            show(g, "row stage 1 (block 0)", 2, 6)
            show(g, "  barrier", 6, 7)
            show(g, "row stage 2 (block 0)", 7, 8)
            show(g, "remaining blocks", 8, 13)
            show(g, "transform warps total", 0, 13)
            show(g, "store warp: first block ready at", 0, 10)
            show(g, "store warp: first block stored in", 10, 11)
            show(g, "store warp: all stored at", 0, 12)
        t0 = m[:, 0, 14].min()
        st, en = m[:, 0, 14] - t0, m[:, 0, 15] - t0
        print(f"  globaltimer ns: starts min {st.min()} med {np.median(st):.0f} max {st.max()}; ends min {en.min()} med {np.median(en):.0f} max {en.max()}")


if __name__ == "__main__":
    main()
