#!/usr/bin/env python
"""Writes profiles/k1_traffic.json from ncu --set full captures of K1 (one report per precision):

    python tools/ncu_traffic.py fp32=gpurun_out/k1_TAG.ncu-rep [fp64=gpurun_out/k1_f64_TAG.ncu-rep] [pr_fp32=gpurun_out/pr_TAG.ncu-rep]

DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and warp instructions (smsp__inst_executed.sum) of the first
hanser_sym_kernel launch in each report, stamped with the digest of the kernel sources they were built from
(pyotf_b200.build.hot_digest: K1s, the phase-retrieval kernels and their headers) and the git commit, so that bench.py only quotes them for the very same sources."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def grab(rep, kernel="hanser_sym_kernel"):
    # a .ncu-rep, or the `ncu -i REPORT --page raw --csv` dump of one (big reports are condensed on the GPU box)
    out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    for r in body:
        if kernel in r[hdr.index("Kernel Name")]:
            def val(key):
                i = hdr.index(key)
                v, u = float(r[i].replace(",", "")), units[i].lower()
                scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
                return v * scale
            return {"dram_bytes": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")),
                    "dram_read_bytes": int(val("dram__bytes_read.sum")), "dram_write_bytes": int(val("dram__bytes_write.sum")),
                    "warp_instructions": int(val("smsp__inst_executed.sum")),
                    "duration_us_under_ncu": val("gpu__time_duration.sum") * ({"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[hdr.index("gpu__time_duration.sum")], 1)),
                    "kernel": r[hdr.index("Kernel Name")], "report": os.path.basename(rep)}
    raise SystemExit(f"{kernel} not found in {rep}")


def main():
    from pyotf_b200 import build

    out = {"source_digest": build.hot_digest(),
           "commit": subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip(),
           "note": "dram__bytes_read.sum + dram__bytes_write.sum and smsp__inst_executed.sum of one K1 launch over the cfg1 stack, "
                   "ncu --set full (cold-ish caches, the launch alone on the GPU).  The PSFi store is absorbed by the 126 MB L2 "
                   "inside ncu's window and written back afterwards, so the per-launch DRAM traffic reads far BELOW the "
                   "algorithmic bytes: nothing is re-read from HBM (inputs are the 0.2 MB compact tables)."}
    for a in sys.argv[1:]:
        prec, rep = a.split("=", 1)
        # pr_fp32=...: the phase-retrieval plane kernel (fixture stack), same counters (bench.py: binding_limit of K2)
        out[prec] = grab(rep, "pr_plane_kernel") if prec.startswith("pr_") else grab(rep)
    with open(os.path.join(ROOT, "profiles", "k1_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
