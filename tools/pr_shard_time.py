#!/usr/bin/env python
"""Under torchrun: z-sharded phase retrieval on the fixture stack -- us per iteration with the peer-window reduction,
with the NCCL all-reduce, and of this rank's shard alone (no exchange: what the exchange costs on top)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bench
    from pyotf_b200 import _engine as eng

    data = bench.fixture_stack().astype(np.float32)
    nz = data.shape[0]
    zall = (np.arange(nz) - (nz + 1) // 2) * 300.0
    lo, hi = eng.shard_bounds(nz, rank, world)
    comm = eng.Comm(rank, world)
    dloc = torch.as_tensor(data[lo:hi], device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for label, use_comm, nccl in (("window", True, False), ("nccl", True, True), ("shard alone", False, False),
                                  ("self window", True, False)):
        if nccl:
            os.environ["PYOTF_B200_PR_NCCL"] = "1"
        else:
            os.environ.pop("PYOTF_B200_PR_NCCL", None)
        if label == "self window":
            os.environ["PYOTF_B200_PR_SELFWIN"] = "1"
        st = eng.PRState(520, 0.85, 1.0, 130, 256, zall[lo:hi], dloc, "fp32", nz_total=nz)
        best = 1e9
        for rep in range(6):
            st.set_pupil(None)
            dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            st.run(100, 0.0, 0.0, comm=comm if use_comm else None)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rep:
                best = min(best, float(t.item()))
        if rank == 0:
            print(f"world {world}: {label:12s} {10 * best:7.2f} us / iteration  ({hi - lo} planes on rank 0, rows per CTA {st.rows_per_cta}, "
                  f"{st.nparts} CTAs)", flush=True)
        del st
    comm.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
