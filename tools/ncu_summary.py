#!/usr/bin/env python
"""Summarise an ncu report (--set full) as a markdown table of the counters DESIGN.md / profiles/ quote.

    python tools/ncu_summary.py gpurun_out/hot_r1.ncu-rep > profiles/r1_hot_kernels.md
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_registers", "occ limit regs (CTAs/SM)"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem (CTAs/SM)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fma.sum", "FMA-pipe warp inst"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__inst_executed_pipe_lsu.sum", "LSU warp inst"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__average_warp_latency_issue_stalled_barrier.pct", "stall barrier %"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier (warps/issue)"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall MIO throttle (warps/issue)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math throttle (warps/issue)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait (warps/issue)"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not selected (warps/issue)"),
]


def main():
    rep = sys.argv[1]
    out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    print(f"# ncu summary of `{rep}` (ncu --set full --clock-control none; per launch, cold-ish caches, serialised)\n")
    seen = {}
    for r in body:
        name = r[hdr.index("Kernel Name")]
        short = name.split("(")[0].replace("void ", "").replace("pyotf::", "")
        seen[short] = seen.get(short, 0) + 1
        if seen[short] > 2:
            continue
        print(f"## {short}  (launch #{seen[short]}, id {r[hdr.index('ID')]})\n")
        print("| counter | value | unit |\n|---|---|---|")
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                print(f"| {label} (`{key}`) | {r[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    main()
