"""Host-side plumbing between the Python API and the C ABI: precision mode, device tensors,
streams and handle caches.  torch is used for device memory / streams only."""
import ctypes as C
import os
import threading

import numpy as np

from . import _lib

_VEC = {"none": 0, "x": 1, "y": 2, "z": 3, "total": 4}
_COND = {"none": 0, "sine": 1, "herschel": 2}

_state = threading.local()
_default_precision = os.environ.get("PYOTF_B200_PRECISION", "fp64")


def set_precision(mode):
    """Select "fp32" (complex64 kernels, rel. L2 <= 1e-5 vs pyotf) or "fp64" (<= 1e-11)."""
    global _default_precision
    if mode not in ("fp32", "fp64"):
        raise ValueError("precision must be 'fp32' or 'fp64'")
    _default_precision = mode


def get_precision():
    return _default_precision


_cuda_checked = False


def _torch():
    """torch, after checking once that a CUDA device is there (every launch path calls this several times)."""
    global _cuda_checked
    import torch

    if not _cuda_checked:
        if not torch.cuda.is_available():
            raise _lib.Pyotf_b200Error("pyotf_b200 needs a CUDA device (B200); there is no CPU fallback")
        _cuda_checked = True
    return torch


def dtype_code(precision):
    return 0 if precision == "fp32" else 1


def torch_dtypes(precision):
    torch = _torch()
    return (torch.float32, torch.complex64) if precision == "fp32" else (torch.float64, torch.complex128)


def stream_ptr():
    return C.c_void_p(_torch().cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class HanserHandle:
    """Owns one pyotf_hanser_t (device geometry tables for fixed optics / size / dtype / support)."""

    def __init__(self, wl, na, ni, res, size, vec_corr, condition, precision, support):
        lib = _lib.load()
        _torch()
        p = _lib.HanserParams(float(wl), float(na), float(ni), float(res), int(size), _VEC[vec_corr],
                              _COND[condition], dtype_code(precision), int(support))
        h = C.c_void_p()
        _lib.check(lib.pyotf_b200_hanser_create(C.byref(p), stream_ptr(), C.byref(h)))
        self._h, self.precision, self.size = h, precision, int(size)
        K, KR, f0, nvec = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _lib.check(lib.pyotf_b200_hanser_info(h, C.byref(K), C.byref(KR), C.byref(f0), C.byref(nvec)))
        self.K, self.KR, self.f0, self.nvec = K.value, KR.value, f0.value, nvec.value

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.load().pyotf_b200_hanser_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def set_overlap(self, enable=True):
        """Streaming mode: back-to-back `psf` launches overlap (see pyotf_b200_hanser_set_overlap)."""
        _lib.check(_lib.load().pyotf_b200_hanser_set_overlap(self._h, int(bool(enable))))

    def tables(self):
        """(kz [K,K] f64, amp [nvec,K,K], mask [K,K] u8) device tensors -- for tests."""
        torch = _torch()
        rt, _ = torch_dtypes(self.precision)
        kz = torch.empty((self.KR, self.K), dtype=torch.float64, device="cuda")
        amp = torch.empty((self.nvec, self.KR, self.K), dtype=rt, device="cuda")
        mask = torch.empty((self.KR, self.K), dtype=torch.uint8, device="cuda")
        _lib.check(_lib.load().pyotf_b200_hanser_tables(self._h, ptr(kz), ptr(amp), ptr(mask), stream_ptr()))
        return kz[:self.K], amp[:, :self.K], mask[:self.K]

    def pack_pupil(self, pupil_dev):
        """[B,N,N] complex128 unshifted device tensor -> compact [B,KR,K] (zero-padded rows) in the handle's dtype."""
        torch = _torch()
        _, ct = torch_dtypes(self.precision)
        B = pupil_dev.shape[0]
        out = torch.empty((B, self.KR, self.K), dtype=ct, device="cuda")
        _lib.check(_lib.load().pyotf_b200_hanser_pack_pupil(self._h, ptr(pupil_dev), B, ptr(out), stream_ptr()))
        return out

    def psf(self, z_dev, pupilc=None, batch=1, want_psfa=False, want_psfi=True, psfa_out=None, psfi_out=None):
        """Run K1.  Returns (psfa or None, psfi or None) as device tensors.  ``z_dev`` is a float64 device tensor or
        a host array of plane positions (pyotf_b200_hanser_psf_zhost: short stacks travel in the kernel arguments)."""
        torch = _torch()
        rt, ct = torch_dtypes(self.precision)
        z_host = None
        if not torch.is_tensor(z_dev):
            z_host = np.ascontiguousarray(np.atleast_1d(np.asarray(z_dev, dtype=np.float64)))
            nz = int(z_host.size)
        else:
            nz = int(z_dev.numel())
        N = self.size
        stride = 0
        if pupilc is not None:
            if pupilc.dim() == 2:
                pupilc = pupilc[None]
            stride = self.KR * self.K if pupilc.shape[0] > 1 else 0
            if pupilc.shape[0] > 1:
                batch = pupilc.shape[0]
        for name, out, dt, shape in (("psfa_out", psfa_out, ct, (batch, self.nvec, nz, N, N)),
                                     ("psfi_out", psfi_out, rt, (batch, nz, N, N))):
            # the C side writes through the raw pointer: anything but a matching contiguous CUDA tensor would be an
            # out-of-bounds or garbled device write
            if out is not None and not (torch.is_tensor(out) and out.is_cuda and out.is_contiguous() and out.dtype == dt
                                        and tuple(out.shape) == shape):
                raise ValueError(f"{name} must be a contiguous CUDA tensor of dtype {dt} and shape {shape}")
        if want_psfa and psfa_out is None:
            psfa_out = torch.empty((batch, self.nvec, nz, N, N), dtype=ct, device="cuda")
        if want_psfi and psfi_out is None:
            psfi_out = torch.empty((batch, nz, N, N), dtype=rt, device="cuda")
        if z_host is not None:
            _lib.check(_lib.load().pyotf_b200_hanser_psf_zhost(
                self._h, z_host.ctypes.data, nz, ptr(pupilc), batch, stride, ptr(psfa_out), ptr(psfi_out), stream_ptr()))
        else:
            _lib.check(_lib.load().pyotf_b200_hanser_psf(
                self._h, ptr(z_dev), nz, ptr(pupilc), batch, stride, ptr(psfa_out), ptr(psfi_out), stream_ptr()))
        return psfa_out, psfi_out


class PRState:
    """Owns one pyotf_pr_t: the device state of a phase retrieval on one measured stack (K2)."""

    def __init__(self, wl, na, ni, res, size, zrange, data_dev, precision, nz_total=None, rows_per_cta=0):
        lib = _lib.load()
        torch = _torch()
        if data_dev.dtype == torch.float32:
            dcode = 0
        elif data_dev.dtype == torch.float64:
            dcode = 1
        else:
            raise TypeError(f"data must be float32 or float64 on the device, not {data_dev.dtype}")
        data_dev = data_dev.contiguous()
        nz = int(data_dev.shape[0])
        z = to_device_f64(np.asarray(zrange, dtype=np.float64))
        if z.numel() != nz:
            raise ValueError(f"{nz} planes of data but {z.numel()} z positions")
        p = _lib.HanserParams(float(wl), float(na), float(ni), float(res), int(size), 0, 0, dtype_code(precision), -2)
        h = C.c_void_p()
        _lib.check(lib.pyotf_b200_pr_create(C.byref(p), ptr(z), nz, int(nz if nz_total is None else nz_total),
                                            ptr(data_dev), dcode, int(rows_per_cta), stream_ptr(), C.byref(h)))
        self._h, self.precision, self.size, self.nz = h, precision, int(size), nz
        K, KR, f0, M, nparts = (C.c_int() for _ in range(5))
        _lib.check(lib.pyotf_b200_pr_info(h, C.byref(K), C.byref(KR), C.byref(f0), C.byref(M), C.byref(nparts)))
        self.K, self.KR, self.f0, self.rows_per_cta, self.nparts = K.value, KR.value, f0.value, M.value, nparts.value

    def kernels_per_iteration(self, sharded=False):
        """Kernel launches per iteration of run() (1 = the plane kernel finalizes the iteration itself)."""
        n = C.c_int()
        _lib.check(_lib.load().pyotf_b200_pr_kernels_per_iteration(self._h, int(bool(sharded)), C.byref(n)))
        return n.value

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.load().pyotf_b200_pr_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def set_pupil(self, pupil_dev=None):
        """Re-seed the estimate from an unshifted [N, N] complex128 device tensor (None: ideal pupil)."""
        if pupil_dev is not None:
            pupil_dev = pupil_dev.contiguous()
        _lib.check(_lib.load().pyotf_b200_pr_set_pupil(self._h, ptr(pupil_dev), stream_ptr()))

    def get_pupil(self, applied=False):
        """Unshifted [N, N] complex128 device tensor: the final pupil, or the last applied one."""
        torch = _torch()
        out = torch.empty((self.size, self.size), dtype=torch.complex128, device="cuda")
        _lib.check(_lib.load().pyotf_b200_pr_get_pupil(self._h, int(bool(applied)), ptr(out), stream_ptr()))
        return out

    def open_window(self, comm):
        """z-sharded runs: export this state's partial-sum window and map every other rank's (CUDA IPC; the
        handles travel through torch.distributed).  Collective over the ranks of `comm`; idempotent."""
        if getattr(self, "_window", False) or comm is None or comm.world < 2:
            return
        import torch.distributed as dist

        lib = _lib.load()
        mine = (C.c_ubyte * 64)()
        if os.environ.get("PYOTF_B200_PR_SELFWIN"):     # timing experiment: the window kernel without any peer (wrong sums)
            _lib.check(lib.pyotf_b200_pr_window_alloc(self._h, 1, C.cast(mine, C.c_void_p)))
            _lib.check(lib.pyotf_b200_pr_window_open(self._h, C.cast(mine, C.c_void_p), 0, 1))
            self._window = True
            return
        _lib.check(lib.pyotf_b200_pr_window_alloc(self._h, int(comm.world), C.cast(mine, C.c_void_p)))
        allh = [None] * comm.world
        dist.all_gather_object(allh, bytes(mine))
        buf = (C.c_ubyte * (64 * comm.world)).from_buffer_copy(b"".join(allh))
        _lib.check(lib.pyotf_b200_pr_window_open(self._h, C.cast(buf, C.c_void_p), int(comm.rank), int(comm.world)))
        dist.barrier()
        self._window = True

    def run(self, max_iters, pupil_tol, mse_tol, phase_only=False, comm=None):
        """Iterate on the device; returns (mse, mse_diff, pupil_diff) trimmed to the iterations run.
        With `comm` (one rank per GPU) the stack is z-sharded: the per-iteration sum over ranks runs inside the
        finalize kernel through peer-mapped windows (`open_window`, done on first use unless
        PYOTF_B200_PR_NCCL is set, which keeps the NCCL all-reduce)."""
        if comm is not None and not os.environ.get("PYOTF_B200_PR_NCCL"):
            self.open_window(comm)
        mse, msed, pd = (np.empty(int(max_iters), dtype=np.float64) for _ in range(3))
        n = C.c_int(0)
        _lib.check(_lib.load().pyotf_b200_pr_run(
            self._h, int(max_iters), float(pupil_tol), float(mse_tol), int(bool(phase_only)),
            comm._h if comm is not None else C.c_void_p(0), mse.ctypes.data_as(C.c_void_p),
            msed.ctypes.data_as(C.c_void_p), pd.ctypes.data_as(C.c_void_p), C.byref(n), stream_ptr()))
        k = n.value
        return mse[:k], msed[:k], pd[:k]


class Comm:
    """NCCL communicator owned by the C library (one process per GPU).

    Rank 0 creates the unique id; it travels through ``torch.distributed`` (any backend)."""

    def __init__(self, rank=None, world=None):
        torch = _torch()
        import torch.distributed as dist

        lib = _lib.load()
        rank = dist.get_rank() if rank is None else rank
        world = dist.get_world_size() if world is None else world
        _lib.check(lib.pyotf_b200_comm_load(nccl_library_path().encode()))
        idbuf = (C.c_ubyte * 128)()
        if rank == 0:
            _lib.check(lib.pyotf_b200_comm_unique_id(C.cast(idbuf, C.c_void_p)))
        obj = [bytes(idbuf)]
        dist.broadcast_object_list(obj, src=0)
        idbuf = (C.c_ubyte * 128).from_buffer_copy(obj[0])
        h = C.c_void_p()
        _lib.check(lib.pyotf_b200_comm_create(C.cast(idbuf, C.c_void_p), int(rank), int(world), C.byref(h)))
        self._h, self.rank, self.world = h, rank, world
        torch.cuda.synchronize()

    def allreduce_f64(self, t):
        _lib.check(_lib.load().pyotf_b200_comm_allreduce_f64(self._h, ptr(t), t.numel(), stream_ptr()))
        return t

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().pyotf_b200_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_library_path():
    """Path of the libnccl.so.2 this process (torch) already has mapped, else the default soname."""
    try:
        with open("/proc/self/maps") as f:
            for line in f:
                if "libnccl.so" in line:
                    return line.split()[-1]
    except OSError:
        pass
    try:
        import nvidia.nccl

        cand = os.path.join(os.path.dirname(nvidia.nccl.__path__[0] + "/"), "lib", "libnccl.so.2")
        if os.path.exists(cand):
            return cand
    except Exception:
        pass
    return "libnccl.so.2"


def bind_to_gpu_numa_node(device=None):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (one process per GPU).

    Pinned staging buffers are then allocated on, and filled by, the memory controller next to the GPU's
    PCIe root; with 8 ranks on a two-socket host this is what keeps the device-to-host copies of the
    e2e path from crossing the socket interconnect.  Returns the node id, or None when the topology is
    not exposed (VMs) -- in which case nothing is changed."""
    torch = _torch()
    try:
        dev = torch.cuda.current_device() if device is None else int(device)
        p = torch.cuda.get_device_properties(dev)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def shard_bounds(n, rank, world):
    """Contiguous block [lo, hi) of n items owned by `rank` (np.array_split sizes)."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


_handles = {}
_handles_lock = threading.Lock()


def hanser_handle(wl, na, ni, res, size, vec_corr, condition, precision, support):
    """Cached pyotf_hanser handle.  A handle owns per-launch scratch (pre-multiplied pupils, per-plane defocus
    tables, uploaded plane positions) that the C side overwrites and regrows on every launch, so it must never be
    shared between host threads or CUDA streams: both are part of the cache key."""
    torch = _torch()
    key = (float(wl), float(na), float(ni), float(res), int(size), vec_corr, condition, precision, int(support),
           torch.cuda.current_device(), int(torch.cuda.current_stream().cuda_stream), threading.get_ident())
    with _handles_lock:
        h = _handles.get(key)
        if h is None:
            if len(_handles) > 64:
                _handles.clear()
            h = _handles[key] = HanserHandle(wl, na, ni, res, size, vec_corr, condition, precision, support)
    return h


def pupil_support(pupil_dev):
    """Largest |frequency bin| where the unshifted [N,N] complex128 device pupil is non-zero (syncs)."""
    out = C.c_int(0)
    _lib.check(_lib.load().pyotf_b200_pupil_support(ptr(pupil_dev), int(pupil_dev.shape[-1]), C.byref(out), stream_ptr()))
    return out.value


def to_device_f64(a):
    torch = _torch()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64)), device="cuda")


def to_device_c128(a):
    torch = _torch()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.complex128)), device="cuda")


def to_device_i32(a):
    torch = _torch()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.int32)), device="cuda")


def zernike_eval(r, theta, n, m, norm):
    """Host arrays in, host [nmodes, npts] float64 out; evaluated by the CUDA kernel."""
    torch = _torch()
    rd, td = to_device_f64(np.ravel(r)), to_device_f64(np.ravel(theta))
    nd, md = to_device_i32(n), to_device_i32(m)
    npts, nmodes = rd.numel(), nd.numel()
    out = torch.empty((nmodes, npts), dtype=torch.float64, device="cuda")
    if npts and nmodes:
        _lib.check(_lib.load().pyotf_b200_zernike(ptr(rd), ptr(td), npts, ptr(nd), ptr(md), nmodes, int(norm),
                                                  ptr(out), stream_ptr()))
    return out.cpu().numpy()


def zernike_lstsq(r, theta, n, m, data):
    """Least-squares coefficients of `data` rows ([nrhs, npts]) on the Zernike modes (n, m) evaluated at the points
    (r, theta): modes and normal equations on the GPU (fp64), the nmodes x nmodes solve on the host.
    Returns [nrhs, nmodes]."""
    torch = _torch()
    rd, td = to_device_f64(np.ravel(r)), to_device_f64(np.ravel(theta))
    nd, md = to_device_i32(n), to_device_i32(m)
    b = to_device_f64(np.atleast_2d(data))
    npts, nmodes, nrhs = rd.numel(), nd.numel(), b.shape[0]
    Z = torch.empty((nmodes, npts), dtype=torch.float64, device="cuda")
    lib = _lib.load()
    _lib.check(lib.pyotf_b200_zernike(ptr(rd), ptr(td), npts, ptr(nd), ptr(md), nmodes, 1, ptr(Z), stream_ptr()))
    G = torch.empty((nmodes, nmodes + nrhs), dtype=torch.float64, device="cuda")
    _lib.check(lib.pyotf_b200_zernike_gram(ptr(Z), ptr(b), npts, nmodes, nrhs, ptr(G), stream_ptr()))
    g = G.cpu().numpy()
    A, rhs = g[:, :nmodes], g[:, nmodes:]
    A = 0.5 * (A + A.T)
    try:
        coefs = np.linalg.solve(A, rhs)
    except np.linalg.LinAlgError:       # rank deficient (e.g. more modes than pixels): minimum-norm solution
        coefs = np.linalg.lstsq(A, rhs, rcond=None)[0]
    return coefs.T


def hanser_kspace(wl, ni, res, size):
    """(kr, phi, kz) float64 [N, N] device tensors, unshifted (HanserPSF._gen_kr)."""
    torch = _torch()
    kr, phi, kz = (torch.empty((size, size), dtype=torch.float64, device="cuda") for _ in range(3))
    _lib.check(_lib.load().pyotf_b200_hanser_kspace(float(wl), float(ni), float(res), int(size), ptr(kr), ptr(phi),
                                                    ptr(kz), stream_ptr()))
    return kr, phi, kz


def hanser_defocus(kz_dev, z_dev):
    torch = _torch()
    nz, npix = int(z_dev.numel()), int(kz_dev.numel())
    out = torch.empty((nz,) + tuple(kz_dev.shape), dtype=torch.complex128, device="cuda")
    _lib.check(_lib.load().pyotf_b200_hanser_defocus(ptr(kz_dev), npix, ptr(z_dev), nz, ptr(out), stream_ptr()))
    return out


def _coef_tensor(c):
    """[B, nmodes] float64 CUDA tensor from a host array or a CUDA tensor (None stays None)."""
    if c is None:
        return None
    torch = _torch()
    if not isinstance(c, np.ndarray) and hasattr(c, "is_cuda"):
        t = c.to(device="cuda", dtype=torch.float64)
        return (t[None] if t.dim() == 1 else t).contiguous()
    return to_device_f64(np.atleast_2d(np.asarray(c, dtype=np.float64)))


def aberration_pupil(handle, n, m, mcoefs, pcoefs):
    """Compact pupils [B, K, K] for [B, nmodes] coefficient arrays / CUDA tensors (either may be None)."""
    torch = _torch()
    _, ct = torch_dtypes(handle.precision)
    mc, pc = _coef_tensor(mcoefs), _coef_tensor(pcoefs)
    ref = mc if mc is not None else pc
    B, nmodes = int(ref.shape[0]), int(ref.shape[1])
    nd, md = to_device_i32(n), to_device_i32(m)
    out = torch.empty((B, handle.KR, handle.K), dtype=ct, device="cuda")
    _lib.check(_lib.load().pyotf_b200_hanser_aberration_pupil(handle._h, ptr(nd), ptr(md), nmodes, ptr(mc), ptr(pc), B,
                                                              ptr(out), stream_ptr()))
    return out


_pinned_pool = []          # [pinned tensor, weakref to the numpy view handed out | _CLAIMED]
_PINNED_POOL_MAX = 8
_PINNED_POOL_BYTES = 256 << 20      # at most this much pinned host memory is retained for reuse
_CLAIMED = object()                 # slot handed to a thread whose copy is still in flight


def _slot_free(ref):
    return ref is None or (ref is not _CLAIMED and ref() is None)


def _pinned_like(t):
    """(slot) with a pinned host tensor of t's shape/dtype.  cudaHostAlloc costs milliseconds for a 32 MiB
    stack, so a buffer is recycled once the numpy array handed out for it (and every view of it) is gone.
    The slot is claimed while the lock is held (two threads never get the same buffer); buffers beyond the
    byte cap are one-off allocations that are not retained."""
    torch = _torch()
    nbytes = t.numel() * t.element_size()
    with _handles_lock:
        for i, slot in enumerate(_pinned_pool):
            buf, ref = slot
            if buf.dtype == t.dtype and buf.shape == t.shape and _slot_free(ref):
                slot[1] = _CLAIMED
                _pinned_pool.append(_pinned_pool.pop(i))
                return slot
        slot = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True), _CLAIMED]
        if nbytes > _PINNED_POOL_BYTES:
            return slot                                   # too large to keep around
        _pinned_pool.append(slot)
        total = sum(b.numel() * b.element_size() for b, _ in _pinned_pool)
        i = 0
        while (len(_pinned_pool) > _PINNED_POOL_MAX or total > _PINNED_POOL_BYTES) and i < len(_pinned_pool) - 1:
            b, ref = _pinned_pool[i]
            if _slot_free(ref):                           # drop idle buffers only, oldest first
                total -= b.numel() * b.element_size()
                _pinned_pool.pop(i)
            else:
                i += 1
        return slot


def to_host(t):
    """CUDA tensor -> numpy array through a pinned staging buffer (the array keeps the buffer alive)."""
    import weakref

    torch = _torch()
    t = t.detach()
    if not t.is_cuda:
        return t.numpy()
    slot = _pinned_like(t)
    slot[0].copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    arr = slot[0].numpy()
    slot[1] = weakref.ref(arr)
    return arr
