"""Multi-GPU paths that need an exchange (one process per GPU, torch.distributed for the plumbing).

OTFi of a z-sharded stack: BasePSF.OTFi is the centred 3D FFT of PSFi
(pyotf/otf.py:209-212, pyotf/utils.py:15-17).  The transform separates per axis (the shifts commute with the
other axes' transforms), so the y / x passes run on each rank's own planes and only the z pass needs every plane of
a (ky, kx) line:

* ``gather``   -- all-gather PSFi (real, nz*N*N*s_r bytes: 32 MiB at BASELINE config 1 in fp32), transform the whole
                  stack locally, keep this rank's z block.  One collective, 3x redundant arithmetic on a tiny FFT.
* ``alltoall`` -- 2D transform of the local planes, all-to-all transpose z-blocks <-> y-blocks, z transform of the
                  local y-block, all-to-all back.  Two collectives of the complex stack, no redundant work, no rank
                  ever holds the whole stack (BASELINE config 4: 1024^2 x 1024 planes = 8 GiB complex64).

Both return this rank's z block of OTFi, i.e. the same sharding as the input.  The local transforms are
parameters so that the exchange logic runs under gloo on CPU tensors in the test-suite (tests/test_gloo_sharding.py);
on the GPU they default to the CUDA shifted FFT (pyotf_b200_fftn) and the backend is NCCL over NVLink.
"""
import numpy as np


def shard_sizes(n, world):
    base, extra = divmod(int(n), int(world))
    return [base + (1 if r < extra else 0) for r in range(world)]


def _default_fft(x, axes):
    from . import _fft

    return _fft.shifted_fft_dev(x, axes=axes, inverse=False)


def _all_to_all(recv, send, group=None):
    """dist.all_to_all (NCCL); point-to-point sends where the backend has no all-to-all (gloo, CPU tests)."""
    import torch.distributed as dist

    if dist.get_backend(group) != "gloo":
        dist.all_to_all(recv, send, group=group)
        return
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    reqs = []
    for r in range(world):
        if r == rank:
            recv[r].copy_(send[r])
            continue
        peer = dist.get_global_rank(group, r) if group is not None else r
        reqs.append(dist.isend(send[r], peer, group=group))
        reqs.append(dist.irecv(recv[r], peer, group=group))
    for q in reqs:
        q.wait()


def otfi_zsharded(psfi_local, nz_total, group=None, mode="auto", fft=None, gather_limit_bytes=64 << 20):
    """Centred 3D FFT of a z-sharded real stack.

    psfi_local : [nz_local, N, N] real tensor, this rank's contiguous z block (np.array_split sizes over the group)
    nz_total   : planes over all ranks
    mode       : "gather", "alltoall" or "auto" (gather while the whole real stack is at most gather_limit_bytes)
    fft        : callable (tensor, axes) -> centred FFT along `axes` (default: the CUDA kernel)
    Returns this rank's [nz_local, N, N] complex block of OTFi."""
    import torch
    import torch.distributed as dist

    fft = fft or _default_fft
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    zs = shard_sizes(nz_total, world)
    nzl, ny, nx = psfi_local.shape
    if nzl != zs[rank]:
        raise ValueError(f"rank {rank} holds {nzl} planes, expected {zs[rank]} of {nz_total}")
    if mode == "auto":
        mode = "gather" if nz_total * ny * nx * psfi_local.element_size() <= gather_limit_bytes else "alltoall"
    lo = sum(zs[:rank])
    if world == 1:
        return fft(psfi_local, (0, 1, 2))
    if mode == "gather":
        zmax = max(zs)
        mine = psfi_local
        if nzl < zmax:                                    # all_gather wants equal shapes: pad the short blocks
            mine = torch.cat([psfi_local, psfi_local.new_zeros((zmax - nzl, ny, nx))])
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine.contiguous(), group=group)
        full = torch.cat([p[:zs[r]] for r, p in enumerate(parts)])
        return fft(full, (0, 1, 2))[lo:lo + nzl].contiguous()
    if mode != "alltoall":
        raise ValueError("mode must be 'gather', 'alltoall' or 'auto'")
    ys = shard_sizes(ny, world)
    ylo = [sum(ys[:r]) for r in range(world)]
    a = fft(psfi_local, (1, 2))                            # [nzl, ny, nx] complex, y / x done
    send = [a[:, ylo[r]:ylo[r] + ys[r], :].contiguous() for r in range(world)]
    recv = [a.new_empty((zs[r], ys[rank], nx)) for r in range(world)]
    _all_to_all(recv, send, group)                         # z-blocks of my y-block from every rank
    b = fft(torch.cat(recv), (0,))                         # [nz_total, ys[rank], nx]: z done
    zlo = [sum(zs[:r]) for r in range(world)]
    send = [b[zlo[r]:zlo[r] + zs[r]].contiguous() for r in range(world)]
    recv = [b.new_empty((nzl, ys[r], nx)) for r in range(world)]
    _all_to_all(recv, send, group)                         # back: my z block, every y-block
    return torch.cat(recv, dim=1).contiguous()


def numpy_centred_fft(x, axes):
    """CPU stand-in for the CUDA shifted FFT (gloo tests): fftshift(fftn(ifftshift(x))) along `axes`."""
    import torch

    a = x.numpy()
    return torch.from_numpy(np.ascontiguousarray(np.fft.fftshift(np.fft.fftn(np.fft.ifftshift(a, axes=axes), axes=axes), axes=axes)))
