"""world_size-2 gloo test (CPU) of the multi-GPU host logic: plane sharding + the one sum all-reduce per
phase-retrieval iteration.  Each rank computes its planes' contribution with the
oracle's arithmetic (standing in for the CUDA plane kernel, which cannot run here), all-reduces the
[sum of new pupils, sum of squared errors] buffer exactly like pyotf_b200_pr_run does, and must reproduce
the unsharded loop."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyotf_oracle as orc

OPT = dict(wl=520, na=0.85, ni=1.0, res=130)


def _stack(size=32, nz=7):
    rng = np.random.default_rng(5)
    z = orc.default_zrange(nz, 300)
    pupil = orc.aberrated_pupil(OPT["wl"], OPT["na"], OPT["ni"], OPT["res"], size, rng.random(10) * 0.1, rng.random(10))
    return orc.psfi(orc.hanser_psfa(OPT["wl"], OPT["na"], OPT["ni"], OPT["res"], size, z, "none", "none", pupil)) * 1e3, z


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyotf_b200._engine import shard_bounds

    data, z = _stack()
    nz, size = data.shape[0], data.shape[-1]
    lo, hi = shard_bounds(nz, rank, world)
    _, kr, _, _, kz = orc.hanser_kspace(OPT["wl"], OPT["ni"], OPT["res"], size)
    mask = orc.hanser_pupil(kr, OPT["na"], OPT["wl"]).real
    pupil = mask.astype(complex)
    defocus = orc.hanser_defocus(kz, z[lo:hi])
    mag = orc.psqrt(data[lo:hi])
    mses = []
    for _ in range(6):
        psfa = np.fft.fftshift(np.fft.ifftn(pupil * defocus, axes=(1, 2)), axes=(1, 2))
        sq = ((data[lo:hi] - abs(psfa) ** 2) ** 2).sum()
        new = np.fft.fftn(np.fft.ifftshift(mag * np.exp(1j * np.angle(psfa)), axes=(1, 2)), axes=(1, 2)) / defocus
        buf = torch.from_numpy(np.concatenate([new.sum(0).view(np.float64).ravel(), [sq]]))
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)                      # the one collective of an iteration
        b = buf.numpy()
        mses.append(b[-1] / (nz * size * size))
        pupil = b[:-1].view(np.complex128).reshape(size, size) / nz * mask
    if rank == 0:
        np.save(out, np.concatenate([np.array(mses), pupil.view(np.float64).ravel()]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_z_sharded_iteration_matches_unsharded(tmp_path):
    out = str(tmp_path / "r.npy")
    port = 29400 + os.getpid() % 500
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    data, z = _stack()
    o = orc.retrieve_phase_loop(data, OPT["wl"], OPT["na"], OPT["ni"], OPT["res"], zrange=z, max_iters=6, pupil_tol=0,
                                mse_tol=0)
    np.testing.assert_allclose(got[:6], o["mse"], rtol=1e-12)
    pupil = got[6:].view(np.complex128).reshape(32, 32)
    assert np.linalg.norm(pupil - o["pupil"]) / np.linalg.norm(o["pupil"]) <= 1e-12


def _otfi_worker(rank, world, port, out, mode):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyotf_b200 import _dist
    from pyotf_b200._engine import shard_bounds

    z = orc.default_zrange(9, 300)                         # 9 planes over 2 ranks: uneven z blocks (5 + 4)
    full = orc.psfi(orc.hanser_psfa(OPT["wl"], OPT["na"], OPT["ni"], OPT["res"], 20, z))     # 20 rows: y blocks 10 + 10
    lo, hi = shard_bounds(9, rank, world)
    mine = torch.from_numpy(np.ascontiguousarray(full[lo:hi]))
    got = _dist.otfi_zsharded(mine, 9, mode=mode, fft=_dist.numpy_centred_fft).numpy()
    np.save(f"{out}.{rank}.npy", got)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("mode", ["gather", "alltoall"])
def test_z_sharded_otfi_exchange_matches_whole_stack(tmp_path, mode):
    """OTFi of a z-sharded stack: both exchange schemes of pyotf_b200._dist return every rank's
    z block of easy_fft(PSFi) (pyotf/otf.py:209-212); the local transforms are numpy here (gloo, CPU), the CUDA
    kernel on the GPU box (tools/multigpu_check.py)."""
    out = str(tmp_path / "o")
    port = 29400 + (os.getpid() + (7 if mode == "gather" else 13)) % 500
    mp.spawn(_otfi_worker, args=(2, port, out, mode), nprocs=2, join=True)
    z = orc.default_zrange(9, 300)
    ref = orc.otfi(orc.psfi(orc.hanser_psfa(OPT["wl"], OPT["na"], OPT["ni"], OPT["res"], 20, z)))
    got = np.concatenate([np.load(f"{out}.{r}.npy") for r in range(2)])
    assert got.shape == ref.shape
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= 1e-13
