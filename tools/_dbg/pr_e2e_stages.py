"""Where the end-to-end time of retrieve_phase(host stack) goes: wall-clock per host stage + cProfile."""
import cProfile, pstats, io, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from pyotf_b200 import _engine as eng
from pyotf_b200.phaseretrieval import retrieve_phase, PhaseRetrieval, unwrap_phase
data = bench.fixture_stack()
print("data", data.dtype, data.shape)
params = dict(bench.PR_PARAMS)
for _ in range(3):
    retrieve_phase(data, dict(params), 100, 0, 0, precision="fp32")
def T(label, fn, n=5):
    best = 1e9
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    print(f"{label:40s} {1e3 * best:8.3f} ms", flush=True)
    return r
T("retrieve_phase total", lambda: retrieve_phase(data, dict(params), 100, 0, 0, precision="fp32"))
T("as_tensor f64 -> cuda", lambda: torch.as_tensor(data, device="cuda"))
T("astype f32 host", lambda: data.astype(np.float32))
d32 = data.astype(np.float32)
T("as_tensor f32 -> cuda", lambda: torch.as_tensor(d32, device="cuda"))
pr = T("PhaseRetrieval.__init__", lambda: PhaseRetrieval(data, dict(params), precision="fp32"))
T("state.set_pupil + run(100)", lambda: (pr.state.set_pupil(None), pr.state.run(100, 0.0, 0.0, False, None)))
T("pr.run (run + result)", lambda: pr.run(100, 0, 0))
new_pupil = T("to_host(get_pupil)", lambda: eng.to_host(pr.state.get_pupil(applied=False)))
T("model.apply_pupil", lambda: pr.model.apply_pupil(pr.state.get_pupil(applied=True)))
T("model._gen_kr", lambda: pr.model._gen_kr())
ang = np.fft.fftshift(np.angle(new_pupil))
T("np.angle + fftshift", lambda: np.fft.fftshift(np.angle(new_pupil)))
os.environ["PYOTF_UNWRAP_PROFILE"] = "1"
T("unwrap_phase", lambda: unwrap_phase(ang), n=2)
del os.environ["PYOTF_UNWRAP_PROFILE"]
T("abs + fftshift", lambda: np.fft.fftshift(abs(new_pupil)))
prof = cProfile.Profile(); prof.enable(); retrieve_phase(data, dict(params), 100, 0, 0, precision="fp32"); prof.disable()
s = io.StringIO(); pstats.Stats(prof, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:5000])
