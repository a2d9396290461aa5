import cProfile, pstats, io, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from pyotf_b200.otf import HanserPSF
model = HanserPSF(precision="fp32", **bench.CFG1)
zr = np.asarray(model.zrange, dtype=np.float64).copy()
for _ in range(6):
    model.zrange = zr; p = model.PSFi
torch.cuda.synchronize()
n = 40
t0 = time.perf_counter()
for _ in range(n):
    model.zrange = zr; p = model.PSFi
dt = (time.perf_counter() - t0) / n
print(f"e2e step {1e6*dt:.1f} us  ({128/dt:.0f} planes/s)")
pr = cProfile.Profile(); pr.enable()
for _ in range(n):
    model.zrange = zr; p = model.PSFi
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:4500])
