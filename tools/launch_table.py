#!/usr/bin/env python
"""Print an ncu --csv launch list (any --metrics set) as one line per launch:  python tools/launch_table.py file.csv [filter]"""
import csv
import sys
from collections import OrderedDict


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ks = OrderedDict()
    for r in rows[hdr + 1:]:
        d = dict(zip(h, r))
        ks.setdefault((d["ID"], d["Kernel Name"].replace("void pyotf::", "")[:72], d["Grid Size"], d["Block Size"]), {})[d["Metric Name"]] = d["Metric Value"]
    for k, v in ks.items():
        if flt in k[1]:
            print(f"{k[1]:72s} {k[2]:>14s} {k[3]:>12s}  " + "  ".join(f"{m.split('__')[-1]}={x}" for m, x in v.items()))


if __name__ == "__main__":
    main()
