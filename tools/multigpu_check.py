#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multigpu_check.py

1. z-sharded phase retrieval must reproduce the single-GPU run (same iteration count, mse, pupil), both with the
   peer-window reduction inside the finalize kernel (default) and with the NCCL all-reduce (PYOTF_B200_PR_NCCL=1).
2. z-sharded HanserPSF stack (no collective): every rank's block equals the matching slice.
3. OTFi of the z-sharded stack (all-gather and all-to-all schemes) equals the matching z block of the single-GPU OTFi.
"""
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
    from pyotf_b200.otf import HanserPSF
    from pyotf_b200.phaseretrieval import retrieve_phase

    data = bench.fixture_stack()
    nz = data.shape[0]
    zall = (np.arange(nz) - (nz + 1) // 2) * 300.0
    lo, hi = eng.shard_bounds(nz, rank, world)
    comm = eng.Comm(rank, world)
    ok = True
    for nccl in (False, True):
        if nccl:
            os.environ["PYOTF_B200_PR_NCCL"] = "1"
        else:
            os.environ.pop("PYOTF_B200_PR_NCCL", None)
        for precision, tol in (("fp64", 1e-12), ("fp32", 1e-5)):
            for ptol in (0.0, 1e-4):
                full = retrieve_phase(data, dict(bench.PR_PARAMS), 60, ptol, ptol, precision=precision)
                part = retrieve_phase(data[lo:hi], dict(bench.PR_PARAMS, zrange=zall[lo:hi]), 60, ptol, ptol,
                                      precision=precision, comm=comm, nz_total=nz)
                e_mse = float(np.abs(part.mse / full.mse - 1).max()) if len(part.mse) == len(full.mse) else np.inf
                e_pup = float(np.linalg.norm(part.pupil - full.pupil) / np.linalg.norm(full.pupil))
                good = len(part.mse) == len(full.mse) and e_mse <= tol and e_pup <= max(tol, 1e-12) * 50
                ok &= good
                print(f"[rank {rank}] PR {'nccl' if nccl else 'window'} {precision} tol={ptol}: iters {len(part.mse)}/{len(full.mse)} "
                      f"mse err {e_mse:.2e} pupil err {e_pup:.2e} {'ok' if good else 'FAIL'}", flush=True)
    os.environ.pop("PYOTF_B200_PR_NCCL", None)
    # z-sharded Hanser stack
    kw = dict(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=256)
    whole = HanserPSF(zsize=32, precision="fp32", **kw)
    lo, hi = eng.shard_bounds(32, rank, world)
    mine = HanserPSF(zrange=whole.zrange[lo:hi], precision="fp32", **kw)
    good = np.array_equal(mine.PSFi, whole.PSFi[lo:hi])
    ok &= good
    print(f"[rank {rank}] Hanser z-shard planes [{lo},{hi}): {'ok' if good else 'FAIL'}", flush=True)
    # OTFi of the z-sharded stack: 27 planes (uneven blocks), both exchange schemes, both precisions
    from pyotf_b200.otf import otfi_zsharded

    for precision, tol in (("fp64", 1e-11), ("fp32", 1e-5)):
        whole = HanserPSF(zsize=27, precision=precision, **kw)
        ref = whole.OTFi
        lo, hi = eng.shard_bounds(27, rank, world)
        mine = HanserPSF(zrange=whole.zrange[lo:hi], precision=precision, **kw)
        for mode in ("gather", "alltoall"):
            got = otfi_zsharded(mine, 27, mode=mode).cpu().numpy()
            err = float(np.linalg.norm(got - ref[lo:hi]) / np.linalg.norm(ref[lo:hi]))
            good = got.shape == ref[lo:hi].shape and err <= tol
            ok &= good
            print(f"[rank {rank}] OTFi z-sharded {precision} {mode}: rel L2 vs single GPU {err:.2e} {'ok' if good else 'FAIL'}", flush=True)
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    comm.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIGPU CHECK", "PASSED" if t.item() == 1.0 else "FAILED", flush=True)
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
