#!/usr/bin/env python
"""Benchmark of the PSF/OTF hot path on B200 (driver contract: see README / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload hanser256|pr|library|sheppard] [--precision fp32|fp64]

Default workload = BASELINE.json config 1, the one the metric "HanserPSF z-planes/s at 256^2"
is quoted on: HanserPSF(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=256, zsize=128) -> PSFi.
One step = one pass of K1 over that 128-plane stack (one kernel launch per GPU).  With N GPUs
every rank owns its own contiguous 128-plane block of a 128*N-plane stack (weak scaling, no
data-path collective).  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG1 = dict(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=256, zsize=128)
# identical in both arms' config.workload (the driver compares the strings)
WORKLOAD = ("hanser256: HanserPSF(wl=520,na=0.85,ni=1.0,res=130,zres=300,size=256,zsize=128) -> PSFi "
            "(BASELINE config 1)")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU arms (the ONLY place bench.py executes oracle/)
# ----------------------------------------------------------------------------------------------
def _cpu_hanser_chunk(args):
    from oracle import pyotf_oracle as orc

    kw, z = args
    a = orc.hanser_psfa(kw["wl"], kw["na"], kw["ni"], kw["res"], kw["size"], z)
    return float(orc.psfi(a).sum())


def cpu_baseline_hanser(budget_s=12.0):
    """Oracle port (numpy, single thread) on the full cfg1 stack, repeated for ~budget_s."""
    from oracle import pyotf_oracle as orc

    z = orc.default_zrange(CFG1["zsize"], CFG1["zres"])
    _cpu_hanser_chunk((CFG1, z))  # warm-up
    reps, t0 = 0, time.perf_counter()
    best = None
    while True:
        t = time.perf_counter()
        _cpu_hanser_chunk((CFG1, z))
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
        reps += 1
        if time.perf_counter() - t0 > budget_s or reps >= 30:
            break
    total = time.perf_counter() - t0
    return {"value": CFG1["zsize"] * reps / total, "unit": "planes/s", "cores": 1, "kind": "port",
            "sample": f"oracle/pyotf_oracle.py (numpy/pocketfft, complex128) on the full cfg1 stack "
                      f"(128 planes of 256^2), {reps} repetitions in {total:.1f} s; best single = "
                      f"{CFG1['zsize'] / best:.1f} planes/s; {os.cpu_count()} cores available, 1 used"}


def run_reference(args):
    """--impl reference: the reference algorithm's CPU implementation (oracle port) using every host core:
    the stack is z-chunked over a process pool (HanserPSF(zrange=chunk) is the reference's own shard API)."""
    import multiprocessing as mp

    from oracle import pyotf_oracle as orc

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = max(1, min(os.cpu_count() or 1, 32))
    z = orc.default_zrange(CFG1["zsize"], CFG1["zres"])
    chunks = [c for c in np.array_split(z, ncores) if len(c)]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(chunks)) as pool:
        for _ in range(args.warmup):
            pool.map(_cpu_hanser_chunk, [(CFG1, c) for c in chunks])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_hanser_chunk, [(CFG1, c) for c in chunks])
        dt = time.perf_counter() - t0
    value = CFG1["zsize"] * args.steps / dt
    line = {
        "impl": "reference", "metric": "HanserPSF z-planes/s at 256^2", "value": value, "unit": "planes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "step": "one 128-plane stack on the host cores, z-chunked over a process pool (numpy complex128)"},
        "cpu_baseline": {"value": value, "unit": "planes/s", "cores": len(chunks), "kind": "port",
                         "sample": f"oracle port (numpy complex128), full cfg1 stack per step over {len(chunks)} "
                                   f"processes"},
        "e2e": {"value": value, "unit": "planes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def fixture_stack():
    """BASELINE config 2 input: the reference's measured 25 x 256^2 stack (committed losslessly under
    tests/golden/), background-subtracted exactly like the reference's demo (prep_data_for_PR(data, 256, 1.1))."""
    from pyotf_b200.utils import prep_data_for_PR

    raw = np.load(os.path.join(ROOT, "tests", "golden", "fixture_psf_520.npz"))["raw"]
    return prep_data_for_PR(raw, 256, 1.1)


PR_PARAMS = dict(wl=520, na=0.85, ni=1.0, res=130, zres=300)


def cpu_baseline_pr(iters=8):
    """Oracle port of retrieve_phase's loop on the fixture stack, `iters` iterations, one core."""
    from oracle import pyotf_oracle as orc

    data = fixture_stack()
    t0 = time.perf_counter()
    orc.retrieve_phase_loop(data, PR_PARAMS["wl"], PR_PARAMS["na"], PR_PARAMS["ni"], PR_PARAMS["res"],
                            zres=PR_PARAMS["zres"], max_iters=iters, pupil_tol=0, mse_tol=0)
    dt = time.perf_counter() - t0
    return {"value": iters / dt, "unit": "iters/s", "cores": 1, "kind": "port",
            "sample": f"oracle/pyotf_oracle.py retrieve_phase_loop on the 25 x 256^2 fixture stack, {iters} iterations "
                      f"in {dt:.1f} s (numpy/pocketfft complex128, 1 of {os.cpu_count()} cores)"}


def bench_pr(precision, rank, world, dist, reps=5, iters=100, extras=True):
    """Phase-retrieval iterations/s on BASELINE config 2 (fixture stack, 100 iterations, tolerances 0 so
    that exactly 100 run).  With N ranks the 25 planes are z-sharded (strong scaling) and every iteration
    all-reduces the partial pupil sums over NCCL."""
    import torch

    from pyotf_b200 import _engine as eng
    from pyotf_b200.phaseretrieval import retrieve_phase

    data = fixture_stack()
    nz_total = data.shape[0]
    lo, hi = eng.shard_bounds(nz_total, rank, world)
    zall = (np.arange(nz_total) - (nz_total + 1) // 2) * PR_PARAMS["zres"]
    comm = eng.Comm(rank, world) if world > 1 else None
    dt = np.float32 if precision == "fp32" else np.float64
    dloc = torch.as_tensor(np.ascontiguousarray(data[lo:hi].astype(dt)), device="cuda")
    st = eng.PRState(PR_PARAMS["wl"], PR_PARAMS["na"], PR_PARAMS["ni"], PR_PARAMS["res"], 256, zall[lo:hi], dloc,
                     precision, nz_total=nz_total)
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for rep in range(reps + 2):
        st.set_pupil(None)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        mse, _, _ = st.run(iters, 0.0, 0.0, comm=comm)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        if rep >= 2:
            times.append(ms)
    assert len(mse) == iters
    best = min(times)
    out = {"iters_per_s": iters / (best * 1e-3), "us_per_iter": 1e3 * best / iters,
           "us_per_iter_median": 1e3 * float(np.median(times)) / iters, "iters": iters, "n_gpus": world,
           "scaling": "strong", "dtype": "f32" if precision == "fp32" else "f64",
           "workload": "retrieve_phase on the 25 x 256^2 fixture stack, 100 iterations, tol 0 (BASELINE config 2)",
           "rows_per_cta": st.rows_per_cta, "ctas_per_iteration": st.nparts,
           "kernels_per_iteration": st.kernels_per_iteration(world > 1), "final_mse": float(mse[-1]),
           # the fused single-GPU loop runs up to 8 iterations per launch behind in-kernel grid barriers
           "iterations_per_launch": (int(os.environ.get("PYOTF_B200_PR_ITERS_PER_LAUNCH", 8))
                                     if st.kernels_per_iteration(world > 1) == 1 else 1),
           "parallelism": f"z-shard x{world}" + ("" if world == 1 else " + per-iteration sum over ranks inside the finalize "
                                                                      "kernel (peer windows over NVLink, no NCCL call)")}
    # roofline of one iteration: nz*N^2*s_r of measured data + 2 compact... full-grid pupils (in/out)
    es = 4 if precision == "fp32" else 8
    alg = nz_total * 256 * 256 * es + 2 * 256 * 256 * 2 * es
    peak, peak_src = measured_peaks()
    out["roofline"] = {"bound": "hbm", "achieved": alg / (best * 1e-3 / iters) / 1e9, "peak": peak, "unit": "GB/s",
                       "frac": alg / (best * 1e-3 / iters) / 1e9 / peak, "algorithmic_bytes_per_iteration": alg,
                       "kernel": "pr_plane_kernel (fused finalize)" if st.kernels_per_iteration(world > 1) == 1 else "pr_plane_kernel + pr_finalize_kernel",
                       "peak_source": peak_src,
                       "note": "the 7.6 MB working set of an iteration is L2-resident (ncu: 7.4 MB of DRAM reads per "
                               "plane-kernel launch only because ncu flushes caches), so HBM is not the binding limit: "
                               "the plane kernel is instruction-issue / latency bound on the 100 SMs its 100 CTAs occupy"}
    # what does bind: instruction issue on the SMs the grid occupies, and the FP32 work of the transforms
    _, winst, _ = k1_traffic("pr_" + precision)
    us = 1e3 * best / iters
    props = torch.cuda.get_device_properties(torch.cuda.current_device())
    sms_used = min(st.nparts, props.multi_processor_count)
    # 2 x nz x (K + N) transforms of N = 256 points, 5 N log2 N flops each, + ~20 flops per pixel of magnitude replacement
    K = 2 * 55 + 1
    nzl = hi - lo
    flops = 2 * nzl * (K + 256) * 5 * 256 * 8 + nzl * 256 * 256 * 20
    fp32_floor_us = flops / (props.multi_processor_count * 128 * 2 * 1.965e3) if precision == "fp32" else None
    out["roofline"]["binding_limit"] = {
        "fp32_floor_us": fp32_floor_us, "fp32_flops_per_iteration": flops, "frac_of_fp32_floor": (fp32_floor_us / us) if fp32_floor_us else None,
        "warp_instructions_per_iteration": winst, "sms_used": sms_used,
        "min_us_at_one_issue_per_cycle_on_the_sms_used": (winst / (sms_used * 4) / 1965.0) if winst else None,
        "frac_of_issue_bound": (winst / (sms_used * 4) / 1965.0 / us) if winst else None,
        "note": "FP32 floor: all 148 SMs at 128 FMA lanes per clock; issue bound: smsp__inst_executed.sum of the stamped ncu "
                "capture of pr_plane_kernel on the SMs its grid occupies (one CTA per SM)"}
    if world == 1:
        # end to end through the public API: host stack in, PhaseRetrievalResult (host arrays) out
        params = dict(PR_PARAMS)
        retrieve_phase(data, dict(params), iters, 0, 0, precision=precision)
        t0 = time.perf_counter()
        r = retrieve_phase(data, dict(params), iters, 0, 0, precision=precision)
        dt_s = time.perf_counter() - t0
        out["e2e"] = {"iters_per_s": iters / dt_s, "ms_total": 1e3 * dt_s, "h2d_bytes": int(data.nbytes),
                      "d2h_bytes": int(r.pupil.nbytes + 3 * 8 * iters),
                      "path": "retrieve_phase(host ndarray) -> PhaseRetrievalResult incl. host unwrap"}
    if world == 1 and extras:
        # the same from the raw camera stack (uint16): background removal / clip on the device (N3), then the loop
        from pyotf_b200.utils import prep_data_for_PR, prep_data_for_PR_dev

        raw = np.load(os.path.join(ROOT, "tests", "golden", "fixture_psf_520.npz"))["raw"]
        retrieve_phase(prep_data_for_PR_dev(raw, 256, 1.1, precision=precision), dict(params), iters, 0, 0, precision=precision)
        t0 = time.perf_counter()
        retrieve_phase(prep_data_for_PR_dev(raw, 256, 1.1, precision=precision), dict(params), iters, 0, 0, precision=precision)
        dt_raw = time.perf_counter() - t0
        t0 = time.perf_counter()
        prep_data_for_PR(raw, 256, 1.1)
        dt_prep_cpu = time.perf_counter() - t0
        out["e2e_from_raw_u16"] = {"iters_per_s": iters / dt_raw, "ms_total": 1e3 * dt_raw, "h2d_bytes": int(raw.nbytes),
                                   "host_prep_ms_for_comparison": 1e3 * dt_prep_cpu,
                                   "path": "prep_data_for_PR_dev(uint16 stack) -> retrieve_phase(device tensor)"}
    if world > 1 and extras:
        # replicas: every rank retrieves the full stack on its own GPU at the same time (no collective) --
        # the batched-independent-retrievals mode SURVEY.md section 8e recommends for short stacks
        dfull = torch.as_tensor(np.ascontiguousarray(data.astype(dt)), device="cuda")
        st2 = eng.PRState(PR_PARAMS["wl"], PR_PARAMS["na"], PR_PARAMS["ni"], PR_PARAMS["res"], 256, zall, dfull, precision)
        rt = []
        for rep in range(reps + 1):
            st2.set_pupil(None)
            dist.barrier()
            torch.cuda.synchronize()
            e0.record(stream)
            st2.run(iters, 0.0, 0.0)
            e1.record(stream)
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rep >= 1:
                rt.append(float(t.item()))
        out["replicas"] = {"iters_per_s": world * iters / (min(rt) * 1e-3), "us_per_iter": 1e3 * min(rt) / iters,
                           "scaling": "weak", "note": f"{world} independent retrievals of the fixture stack, one per GPU"}
    if comm is not None:
        comm.close()
    return out


def _time_gpu(fn, reps, warmup=2):
    """Best CUDA-event time (ms) of fn() over `reps` repetitions on the current stream."""
    import torch

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(warmup):
        fn()
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best


def _time_gpu_ranks(fn, reps, dist, warmup=1):
    """Best over `reps` of the max-over-ranks CUDA-event time (ms) of fn(), every repetition behind a barrier."""
    import torch

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(warmup):
        fn()
    best = None
    for _ in range(reps):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        best = ms if best is None else min(best, ms)
    return best


def bench_other_configs(precision, rank, world, dist=None):
    """Device-resident throughput of BASELINE configs 3, 4 and 5.  Configs 4 and 5 are multi-GPU workloads: every rank
    runs its WHOLE shard (1024 / world planes of the 1024^2 stack; 8192 / world library entries), the time is the max over
    ranks and the rates are whole-job aggregates.  (Parity of each config is covered by tests/.)"""
    import torch

    from pyotf_b200 import _engine as eng
    from pyotf_b200 import _fft
    from pyotf_b200.otf import HanserPSF, SheppardPSF, _aberrated_pupils

    out = {}
    rt = torch.float32 if precision == "fp32" else torch.float64
    es = 4 if precision == "fp32" else 8
    # ---- config 5: PSF library, 8192 x (120 Noll phase coefs -> 256^2 x 64 PSFi), batch-sharded -------------
    base = HanserPSF(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=256, zsize=64, precision=precision)
    coefs = np.random.default_rng(0).standard_normal((8192, 120)) * 0.1
    lo, hi = eng.shard_bounds(8192, rank, world)
    B = 128                                                   # entries per launch: 2.1 GiB of fp32 PSFi
    mine = coefs[lo:hi]
    ring = [torch.empty((B, 64, 256, 256), dtype=rt, device="cuda") for _ in range(2)]
    z = eng.to_device_f64(base.zrange)
    pc = [eng.to_device_f64(mine[i:i + B]) for i in range(0, len(mine), B)]

    def lib_shard():
        for i, c in enumerate(pc):
            pupils = _aberrated_pupils(base, None, c)
            o = ring[i % 2] if c.shape[0] == B else ring[i % 2][:c.shape[0]]
            pupils.handle(base).psf(z, pupils.tensor, want_psfa=False, want_psfi=True, psfi_out=o)

    ms = _time_gpu_ranks(lib_shard, 2, dist)
    out["library_cfg5"] = {"entries_per_s": 8192 / (ms * 1e-3), "planes_per_s": 8192 * 64 / (ms * 1e-3), "ms_whole_library": ms,
                           "entries": 8192, "entries_per_rank": len(mine), "entries_per_launch": B, "n_gpus": world,
                           "achieved_GBps_per_gpu": len(mine) * 64 * 256 * 256 * es / (ms * 1e-3) / 1e9,
                           "scaling": "strong (the whole 8192-entry library, batch-sharded; max over ranks)",
                           "note": "120 Noll coefs -> aberrated pupil kernel + K1 per 128-entry launch; outputs alternate over "
                                   "2 x 2.1 GiB buffers (the library is 137 GB of PSFi: streamed, never resident)"}
    del ring, pc
    torch.cuda.empty_cache()
    # ---- config 4: 1024^2 x 1024 planes with 15 Zernike phase terms, z-sharded --------------------------------
    big = HanserPSF(wl=520, na=0.85, ni=1.0, res=130, zres=300, size=1024, zsize=1024, precision=precision)
    lo, hi = eng.shard_bounds(1024, rank, world)
    chunk = 64
    zcs = [eng.to_device_f64(big.zrange[a:min(a + chunk, hi)]) for a in range(lo, hi, chunk)]
    pup = _aberrated_pupils(big, None, (np.random.default_rng(0).standard_normal(15) * 0.2)[None])
    hb = pup.handle(big)
    outb = [torch.empty((1, chunk, 1024, 1024), dtype=rt, device="cuda") for _ in range(2)]

    def big_shard():
        for i, zc in enumerate(zcs):
            o = outb[i % 2] if zc.numel() == chunk else outb[i % 2][:, :zc.numel()].contiguous()
            hb.psf(zc, pup.tensor, want_psfa=False, want_psfi=True, psfi_out=o)

    ms = _time_gpu_ranks(big_shard, 2, dist)
    out["hanser1024_cfg4"] = {"planes_per_s": 1024 / (ms * 1e-3), "ms_whole_stack": ms, "planes": 1024, "planes_per_rank": hi - lo,
                              "n_gpus": world, "achieved_GBps_per_gpu": (hi - lo) * 1024 * 1024 * es / (ms * 1e-3) / 1e9,
                              "scaling": "strong (the whole 1024-plane stack, z-sharded; max over ranks)",
                              "note": "K1br: two register-resident passes per plane (csrc/hanser_bigsym.cuh), 64-plane calls "
                                      "(outputs alternate over 2 x 256 MiB buffers)"}
    del outb
    torch.cuda.empty_cache()
    if rank == 0:
        # ---- config 3: SheppardPSF 256^3 (zres=190; the 200 of BASELINE.json raises ValueError in the reference) ----
        sh = SheppardPSF(wl=520, na=1.27, ni=1.33, res=100, zres=190, size=256, zsize=256, precision=precision)

        NB = 8                                                # back-to-back calls per repetition (steady state)

        def shep_step():
            for _ in range(NB):
                sh._attribute_changed()
                sh.PSFi_dev

        ms = _time_gpu(shep_step, reps=4) / NB
        out["sheppard_cfg3"] = {"volumes_per_s": 1 / (ms * 1e-3), "ms_per_volume": ms,
                                "achieved_GBps": 256 ** 3 * es / (ms * 1e-3) / 1e9,
                                "note": "shell generator + shifted 3D inverse FFT + |.|^2 -> PSFi (replicas only across GPUs); "
                                        "8 back-to-back volumes per timed repetition"}
        # ---- config 1 epilogue: OTFi from the resident PSFi --------------------------------------------------------
        m = HanserPSF(precision=precision, **CFG1)
        psfi = m.PSFi_dev
        def otf_step():
            for _ in range(NB):
                _fft.shifted_fft_dev(psfi, axes=None, inverse=False)

        ms = _time_gpu(otf_step, reps=4) / NB
        out["otfi_cfg1"] = {"ms": ms, "achieved_GBps": 128 * 256 * 256 * 3 * es / (ms * 1e-3) / 1e9,
                            "note": "shifted 3D FFT of the resident 128 x 256^2 PSFi through the Hermitian half-spectrum passes "
                                    "(algorithmic bytes: real read + complex write); 8 back-to-back transforms per timed repetition"}
    if dist is not None:
        # OTFi of a z-sharded stack (cfg1 planes per rank, all-gather or all-to-all, pyotf_b200/_dist.py)
        from pyotf_b200 import _dist

        nzl = CFG1["zsize"]
        zall = (np.arange(nzl * world) - (nzl * world + 1) // 2) * CFG1["zres"]
        mloc = HanserPSF(wl=CFG1["wl"], na=CFG1["na"], ni=CFG1["ni"], res=CFG1["res"], size=CFG1["size"],
                         zrange=zall[rank * nzl:(rank + 1) * nzl], precision=precision)
        pl = mloc.PSFi_dev
        for mode in ("gather", "alltoall"):
            ms = _time_gpu_ranks(lambda: _dist.otfi_zsharded(pl, nzl * world, mode=mode), 3, dist)
            out[f"otfi_zsharded_{mode}"] = {"ms": ms, "planes": nzl * world, "planes_per_rank": nzl, "n_gpus": world,
                                            "note": "PSFi resident and z-sharded -> this rank's z block of OTFi of the whole stack"}
    torch.cuda.empty_cache()
    return out


def bench_pr_weak(precision, rank, world, dist, planes_per_rank=128, iters=20, reps=3):
    """Phase retrieval, weak scaling: a synthetic stack of planes_per_rank * world planes (the PSFi of a pupil with
    random Noll 5-15 magnitude / phase terms, the recipe of the reference's tests/integration_test.py:50-58), z-sharded.
    Every iteration sums the per-rank partial pupils across the ranks inside the finalize kernel (peer windows)."""
    import torch

    from pyotf_b200 import _engine as eng
    from pyotf_b200.otf import HanserPSF, apply_aberration

    nzt = planes_per_rank * world
    zall = (np.arange(nzt) - (nzt + 1) // 2) * PR_PARAMS["zres"]
    zloc = zall[rank * planes_per_rank:(rank + 1) * planes_per_rank].astype(np.float64)
    rng = np.random.default_rng(12345)
    pc, mc = np.zeros(15), np.zeros(15)
    pc[4:], mc[4:] = rng.random(11), rng.random(11)
    model = HanserPSF(wl=PR_PARAMS["wl"], na=PR_PARAMS["na"], ni=PR_PARAMS["ni"], res=PR_PARAMS["res"], size=256,
                      zrange=zloc, vec_corr="none", condition="none", precision=precision)
    data = (apply_aberration(model, mc, pc).PSFi_dev * 1e4).contiguous()
    st = eng.PRState(PR_PARAMS["wl"], PR_PARAMS["na"], PR_PARAMS["ni"], PR_PARAMS["res"], 256, zloc, data, precision,
                     nz_total=nzt)
    comm = eng.Comm(rank, world) if world > 1 else None
    state = {}

    def run():
        st.set_pupil(None)
        state["mse"] = st.run(iters, 0.0, 0.0, comm=comm)[0]

    ms = _time_gpu_ranks(run, reps, dist)
    if comm is not None:
        comm.close()
    return {"us_per_iter": 1e3 * ms / iters, "iters_per_s": iters / (ms * 1e-3), "plane_iters_per_s": nzt * iters / (ms * 1e-3),
            "planes_per_rank": planes_per_rank, "planes_total": nzt, "n_gpus": world, "iters": iters, "scaling": "weak",
            "dtype": "f32" if precision == "fp32" else "f64", "final_mse": float(state["mse"][-1]),
            "note": "synthetic stack (Noll 5-15 pupil), z-sharded; set_pupil + the 20-iteration loop inside the timed region; "
                    "weak-scaling efficiency = us_per_iter at 1 GPU / us_per_iter at N"}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def k1_traffic(precision):
    """DRAM bytes and warp instructions per K1 launch from the committed ncu capture -- only if that capture was
    taken from the kernel sources this run uses (tools/ncu_traffic.py stamps the source digest)."""
    try:
        from pyotf_b200 import build as _build

        with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("source_digest") != _build.hot_digest():
            return None, None, "profiles/k1_traffic.json was captured from other kernel sources than this run's: not reported"
        rec = tj.get(precision) or {}
        return rec.get("dram_bytes"), rec.get("warp_instructions"), tj.get("note")
    except Exception as e:                                  # no capture committed
        return None, None, f"no ncu capture ({e.__class__.__name__})"


def measure_k1(precision, args, rank, world, local, dist, sample_clocks):
    """cfg1 through pyotf_b200_hanser_psf_zhost (host zrange, the call HanserPSF.PSFi makes), through
    pyotf_b200_hanser_psf (device zrange) and end to end through the Python API, in one precision."""
    import torch

    from pyotf_b200 import _engine as eng
    from pyotf_b200 import _lib
    from pyotf_b200.otf import HanserPSF

    rt = torch.float32 if precision == "fp32" else torch.float64
    esize = 4 if precision == "fp32" else 8
    N, nz = CFG1["size"], CFG1["zsize"]
    # weak scaling: rank r owns planes [r*nz, (r+1)*nz) of an nz*world-plane stack
    zall = (np.arange(nz * world) - (nz * world + 1) // 2) * CFG1["zres"]
    zr = np.ascontiguousarray(zall[rank * nz:(rank + 1) * nz].astype(np.float64))
    h = eng.hanser_handle(CFG1["wl"], CFG1["na"], CFG1["ni"], CFG1["res"], N, "none", "sine", precision, -1)
    zd = eng.to_device_f64(zr)
    # 8 x 32 MiB (fp32) / 5 x 64 MiB (fp64) outputs > 126 MB L2: every step's stores must reach HBM
    ring = 8 if precision == "fp32" else 5
    outs = [torch.empty((1, nz, N, N), dtype=rt, device="cuda") for _ in range(ring)]
    lib = _lib.load()
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    calls = [(h._h, C.c_void_p(zr.ctypes.data), nz, C.c_void_p(0), 1, 0, C.c_void_p(0), C.c_void_p(o.data_ptr()), sp)
             for o in outs]
    calls_dev = [(h._h, C.c_void_p(zd.data_ptr()), nz, C.c_void_p(0), 1, 0, C.c_void_p(0), C.c_void_p(o.data_ptr()), sp)
                 for o in outs]

    def step(i):
        rc = lib.pyotf_b200_hanser_psf_zhost(*calls[i % ring])
        if rc:
            _lib.check(rc)

    def step_dev(i):
        rc = lib.pyotf_b200_hanser_psf(*calls_dev[i % ring])
        if rc:
            _lib.check(rc)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local) if (sample_clocks and rank == 0) else None
    if sampler:
        sampler.start()
    # keep the GPU busy long enough for nvidia-smi to see clocks under load: repeat the K-step
    # measurement until >= min_seconds of wall time, report the best repetition (each is exactly K steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps, best_ms, t_wall = 0, None, time.perf_counter()
    min_s = args.min_seconds if sample_clocks else min(args.min_seconds, 0.3)
    while True:
        barrier()
        e0.record(stream)
        for i in range(args.steps):
            step(i)
        e1.record(stream)
        barrier()
        ms = allmax(e0.elapsed_time(e1))
        best_ms = ms if best_ms is None else min(best_ms, ms)
        reps += 1
        if time.perf_counter() - t_wall > min_s or reps >= 200:
            break
    clocks = sampler.stop() if sampler else None
    ms_per_step = best_ms / args.steps
    # z on the device (pyotf_b200_hanser_psf, no promise about who wrote it): every launch waits for its
    # predecessor before it reads anything, so consecutive stacks do not overlap
    for i in range(3):
        step_dev(i)
    serial_ms = None
    for _ in range(20):
        barrier()
        e0.record(stream)
        for i in range(args.steps):
            step_dev(i)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        serial_ms = ms if serial_ms is None else min(serial_ms, ms)
    serial_ms = allmax(serial_ms) / args.steps
    # context for the roofline: what a plain memset of the very same output buffers costs (cudaMemsetAsync through
    # torch's zero_(), same rotation over the ring, same K launches per repetition) -- a write-only stream of this
    # granularity does not reach the copy bandwidth MEASURED_PEAKS.json quotes
    memset_ms = None
    for _ in range(20):
        e0.record(stream)
        for i in range(args.steps):
            outs[i % ring].zero_()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        memset_ms = ms if memset_ms is None else min(memset_ms, ms)
    memset_ms /= args.steps

    # ---- e2e through the public API: host zrange in, host PSFi (numpy) out, every step ----
    model = HanserPSF(precision=precision, **CFG1)
    z_host = torch.from_numpy(zr.copy()).pin_memory()
    p = None
    for _ in range(4):   # warm-up: lets the pinned staging pool reach its steady state (two buffers)
        model.zrange = z_host.numpy()
        p = model.PSFi
    barrier()
    ksteps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(ksteps):
        model.zrange = z_host.numpy()
        p = model.PSFi
    torch.cuda.synchronize()
    e2e_s = allmax((time.perf_counter() - t0) / ksteps)
    assert p.shape == (nz, N, N) and p.dtype == (np.float32 if precision == "fp32" else np.float64)
    del outs
    torch.cuda.empty_cache()
    peak, peak_src = measured_peaks()
    alg_bytes = nz * N * N * esize  # N^2 * s_r per plane -> PSFi only
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
    traffic, winst, traffic_note = k1_traffic(precision)
    binding = None
    if winst:
        props = torch.cuda.get_device_properties(torch.cuda.current_device())
        mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
        min_us = winst / (props.multi_processor_count * 4) / mhz
        binding = {"kind": "instruction issue", "warp_instructions_per_launch": winst,
                   "schedulers": props.multi_processor_count * 4, "sm_mhz": mhz, "min_us_at_one_issue_per_cycle": min_us,
                   "frac": min_us / (ms_per_step * 1e3), "source": "smsp__inst_executed.sum of the stamped ncu capture"}
    return {
        "precision": precision, "reps": reps, "ring": ring, "clocks": clocks,
        "ms_per_step": ms_per_step, "planes_per_s": nz * world / (ms_per_step * 1e-3),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                     "kernel": "hanser_sym_kernel", "algorithmic_bytes_per_launch": alg_bytes,
                     "binding_limit": binding,
                     "memset_of_the_same_buffers": {
                         "ms_per_step": memset_ms, "achieved": alg_bytes / (memset_ms * 1e-3) / 1e9,
                         "note": "torch zero_() (cudaMemsetAsync) of the same output ring, K back-to-back launches: the "
                                 "write-only stream a kernel that computes nothing would produce"},
                     "device_z_serialized": {
                         "ms_per_step": serial_ms, "planes_per_s": nz * world / (serial_ms * 1e-3),
                         "achieved": alg_bytes / (serial_ms * 1e-3) / 1e9,
                         "frac": alg_bytes / (serial_ms * 1e-3) / 1e9 / peak,
                         "note": "pyotf_b200_hanser_psf with zrange in device memory: each launch waits for its "
                                 "predecessor before reading anything (consecutive stacks do not overlap)"}},
        "e2e": {"value": nz * world / e2e_s, "unit": "planes/s", "h2d_bytes_per_step": int(zr.nbytes),
                "d2h_bytes_per_step": int(nz * N * N * esize), "ms_per_step": 1e3 * e2e_s,
                "path": "HanserPSF.zrange = host array; HanserPSF.PSFi -> numpy (pinned D2H)"},
    }


def run_ours(args):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from pyotf_b200 import _engine as eng

    numa_node = eng.bind_to_gpu_numa_node(local) if world > 1 else None   # before any pinned allocation
    precision = args.precision
    other = "fp64" if precision == "fp32" else "fp32"
    k1 = measure_k1(precision, args, rank, world, local, dist, True)
    # the same measurement in the other precision (fp64 = the reference's own arithmetic, pyotf/otf.py:426)
    k1_other = measure_k1(other, args, rank, world, local, dist, False) if not args.single_precision else None

    pr = pr_other = None
    if not args.no_pr:
        pr = bench_pr(precision, rank, world, dist)
        if not args.single_precision:
            pr_other = bench_pr(other, rank, world, dist, reps=3, extras=False)
    pr_weak = None
    if not args.no_pr:
        pr_weak = bench_pr_weak(precision, rank, world, dist)
    others = None
    if not args.no_others:
        others = bench_other_configs(precision, rank, world, dist)

    if rank == 0:
        nz = CFG1["zsize"]
        line = {
            "metric": "HanserPSF z-planes/s at 256^2", "value": k1["planes_per_s"], "unit": "planes/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": k1["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if precision == "fp32" else "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "step": "one fused K1 launch over the 128-plane stack per GPU (one CTA per plane) through "
                               "pyotf_b200_hanser_psf_zhost, the call HanserPSF.PSFi makes: zrange by value in the kernel "
                               "arguments; consecutive stacks overlap by programmatic dependent launch (a launch whose "
                               "output is disjoint from the launches still in flight only has to finish after them, "
                               "one that overwrites a recent output waits before its first store)",
                       "l2": f"outputs rotate over {k1['ring']} buffers of {nz * 65536 * (4 if precision == 'fp32' else 8) >> 20} MiB "
                             f"(> 126 MB L2)",
                       "timing": f"best of {k1['reps']} repetitions of exactly {args.steps} steps, CUDA events on the "
                                 f"launch stream, max over ranks",
                       "parallelism": f"z-shard x{world} (no collective)"},
            "roofline": k1["roofline"],
            "clocks": k1["clocks"],
            "e2e": dict(k1["e2e"], numa_node_rank0=numa_node),
            # per step: one hanser_sym_kernel launch (128 CTAs, one per plane; every CTA builds its plane's defocus table)
            "gpu_launches": args.steps,
        }
        if k1_other is not None:
            # the other precision mode, same workload / timing protocol (fp64: the reference arm's arithmetic)
            line[other] = {"value": k1_other["planes_per_s"], "unit": "planes/s", "ms_per_step": k1_other["ms_per_step"],
                           "dtype": "f64" if other == "fp64" else "f32", "roofline": k1_other["roofline"],
                           "e2e": k1_other["e2e"],
                           "l2": f"outputs rotate over {k1_other['ring']} buffers (> 126 MB L2)"}
        if pr_weak is not None:
            line["phase_retrieval_weak"] = pr_weak
        if others is not None:
            line["other_configs"] = others       # whole-job aggregates over all ranks (max-over-ranks time)
        if pr is not None:
            line["phase_retrieval"] = pr
            if pr_other is not None:
                pr[other] = pr_other
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_hanser()
            if pr is not None:
                pr["cpu_baseline"] = cpu_baseline_pr()
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hanser256")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-pr", action="store_true", help="skip the phase-retrieval leg")
    ap.add_argument("--no-others", action="store_true", help="skip the config 3/4/5 legs")
    ap.add_argument("--single-precision", action="store_true", help="skip the sub-records in the other precision mode")
    ap.add_argument("--min-seconds", type=float, default=1.0,
                    help="repeat the K-step measurement for at least this long (clock sampling); 0 = once")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
