#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS-instruction counters by CUDA source line.

    python tools/ncu_by_line.py REPORT.ncu-rep KERNEL_MANGLED_SUBSTR CUBIN_BASENAME [--top 40]

ncu's CSV source page only exports the SASS view, so the line mapping comes from
`nvdisasm -g` on the cubin extracted from pyotf_b200/libpyotf_b200.so (instruction order is
the join key).  Prints instructions executed, stall samples and shared-memory wavefront
excess per source line -- what DESIGN.md / profiles/*.md quote.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUTER = "--outer" in sys.argv   # attribute inlined code to the outermost call site


def sass_lines(cubin_base, kernel_substr):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "pyotf_b200", "libpyotf_b200.so")], cwd=tmp,
                   check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.startswith(cubin_base)][0]
    out = subprocess.run(["nvdisasm", "-gi" if OUTER else "-g", "-c", os.path.join(tmp, cubin)],
                         capture_output=True, text=True).stdout
    lines, cur, active, fresh = [], None, False, True
    for ln in out.splitlines():
        if ln.startswith("//--------------------- .text."):
            active = kernel_substr in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            # with -gi a chain of "X inlined at Y" comments precedes an instruction: keep the outermost
            if OUTER and m.group(3):
                cur = (os.path.basename(m.group(3)), int(m.group(4)))
            elif fresh or not OUTER:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
            fresh = False
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines.append((int(m.group(1), 16), cur, m.group(2).strip()))
            fresh = True
    return lines


def main():
    rep, ksub, cubin = sys.argv[1:4]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    # the report may hold several launches: keep the first block whose kernel name matches
    blocks = re.split(r'(?m)^"Kernel Name",', raw)
    blk = [b for b in blocks if b.strip()][0]
    body = blk.split("\n", 1)[1]
    rows = list(csv.DictReader(io.StringIO(body)))
    sass = sass_lines(cubin, ksub)
    if len(sass) != len(rows):
        print(f"warning: {len(sass)} disassembled vs {len(rows)} profiled instructions", file=sys.stderr)
    agg = collections.defaultdict(lambda: collections.Counter())
    tot = collections.Counter()
    for (addr, loc, text), r in zip(sass, rows):
        inst = int(float(r["Instructions Executed"] or 0))
        samp = int(float(r["# Samples"] or 0))
        exc = int(float(r.get("L1 Wavefronts Shared Excessive") or 0))
        a = agg[loc]
        a["inst"] += inst
        a["samples"] += samp
        a["smem_excess"] += exc
        for k in ("stall_long_sb", "stall_short_sb", "stall_barrier", "stall_mio", "stall_math", "stall_wait",
                  "stall_lg", "stall_not_selected", "stall_dispatch"):
            a[k] += int(float(r.get(k) or 0))
        tot["inst"] += inst
        tot["samples"] += samp
    print(f"total warp-instructions {tot['inst']}, samples {tot['samples']}")
    print(f"{'line':>22} {'inst%':>6} {'samp%':>6} {'long_sb':>7} {'short':>6} {'barr':>6} {'mio':>5} {'math':>5} "
          f"{'wait':>5} {'lg':>5} {'nsel':>5} {'smemX':>7}")
    order = (lambda kv: (kv[0] or ("", 0))) if "--by-line" in sys.argv else (lambda kv: -kv[1]["samples"])
    for loc, a in sorted(agg.items(), key=order)[:top]:
        name = f"{loc[0]}:{loc[1]}" if loc else "?"
        print(f"{name:>22} {100 * a['inst'] / max(tot['inst'], 1):6.2f} {100 * a['samples'] / max(tot['samples'], 1):6.2f} "
              f"{a['stall_long_sb']:7d} {a['stall_short_sb']:6d} {a['stall_barrier']:6d} {a['stall_mio']:5d} "
              f"{a['stall_math']:5d} {a['stall_wait']:5d} {a['stall_lg']:5d} {a['stall_not_selected']:5d} "
              f"{a['smem_excess']:7d}")


if __name__ == "__main__":
    main()
