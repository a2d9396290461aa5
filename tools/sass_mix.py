#!/usr/bin/env python
"""Static SASS instruction mix of the kernels of an object file whose (demangled) name contains a pattern:
    python tools/sass_mix.py pyotf_b200/csrc/_obj/fftnd_f32.o fft_lines_reg_kernel"""
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    name, counts = None, {}
    classes = [("fp", r"\b(FFMA|FADD|FMUL|FFMA2|FADD2|FMUL2)\b"), ("dp", r"\b(DFMA|DADD|DMUL)\b"), ("lds", r"\bLDS"), ("sts", r"\bSTS"),
               ("ldg", r"\bLDG"), ("stg", r"\bSTG"), ("int", r"\b(IMAD|IADD3|IADD|LEA|LOP3|SHF|ISETP|SEL|PRMT|VIADD)\b"), ("mov", r"\bMOV\b"),
               ("bra", r"\bBRA\b"), ("bar", r"\bBAR\b")]
    res = []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            counts = {"total": 0}
            res.append((name, counts))
            continue
        if name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            counts["total"] += 1
            for k, rx in classes:
                if re.search(rx, line):
                    counts[k] = counts.get(k, 0) + 1
    for name, c in res:
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        if pat in dem:
            print(dem[:110])
            print("   " + "  ".join(f"{k}={v}" for k, v in c.items()))


if __name__ == "__main__":
    main()
