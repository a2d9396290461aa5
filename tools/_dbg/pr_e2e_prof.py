import cProfile, pstats, io, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from pyotf_b200.phaseretrieval import retrieve_phase
data = bench.fixture_stack()
params = dict(bench.PR_PARAMS)
for _ in range(2):
    retrieve_phase(data, dict(params), 100, 0, 0, precision="fp32")
t0 = time.perf_counter(); r = retrieve_phase(data, dict(params), 100, 0, 0, precision="fp32"); print("total ms", 1e3 * (time.perf_counter() - t0))
pr = cProfile.Profile(); pr.enable(); retrieve_phase(data, dict(params), 100, 0, 0, precision="fp32"); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
