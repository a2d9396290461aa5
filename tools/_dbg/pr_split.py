import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from pyotf_b200 import _engine as eng
data = bench.fixture_stack()
zall = (np.arange(25) - 13) * 300.0
st = eng.PRState(520, 0.85, 1.0, 130, 256, zall, torch.as_tensor(data.astype(np.float32), device="cuda"), "fp32")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(5):
    st.set_pupil(None); torch.cuda.synchronize(); e0.record()
    try:
        st.run(100, 0.0, 0.0)
    except Exception as e:
        pass
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"skip={os.environ.get('PYOTF_B200_PR_DEBUG_SKIP')} pdl_off={os.environ.get('PYOTF_B200_NO_PDL')}: {10*best:7.2f} us/iter", flush=True)
