"""Builds pyotf_b200/libpyotf_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m pyotf_b200.build [--force] [--verbose]

Each translation unit is compiled in parallel to an object under pyotf_b200/csrc/_obj and
linked into one shared library next to this file.  cudart is linked statically so the
library has no load-time dependency beyond libstdc++/libc (the CUDA driver comes from the
process -- torch -- that loads it).
"""
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libpyotf_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-Xptxas", "-v"] + os.environ.get("PYOTF_B200_BUILD_DEFS", "").split()   # e.g. -DPYOTF_K1_TIMING (debug marks)


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the CUDA extension cannot be built")
    return exe


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(HERE, "..", "include", "pyotf_b200.h"), "rb").read())
    h.update(" ".join(ARCH + FLAGS).encode())
    return h.hexdigest()


def _deps(path, seen=None):
    """The file and every local header it includes (transitively)."""
    import re

    seen = set() if seen is None else seen
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    for inc in re.findall(r'^\s*#\s*include\s+"([^"]+)"', open(path).read(), flags=re.M):
        _deps(os.path.normpath(os.path.join(os.path.dirname(path), inc)), seen)
    return seen


def _tu_digest(src):
    h = hashlib.sha256()
    for f in sorted(_deps(os.path.join(CSRC, src))):
        h.update(f.encode())
        h.update(open(f, "rb").read())
    h.update(" ".join(ARCH + FLAGS).encode())
    return h.hexdigest()


def hot_digest():
    """Digest of everything the two headline kernels are compiled from (K1s and the phase-retrieval kernels with their
    transitive headers, the build flags): what tools/ncu_traffic.py stamps its ncu counters with, and what bench.py
    compares before it quotes them.  Sources of other kernels do not enter."""
    h = hashlib.sha256()
    files = set()
    for src in ("hanser_sym_f32.cu", "hanser_sym_f64.cu", "pr_f32.cu"):
        files |= _deps(os.path.join(CSRC, src))
    for f in sorted(files):
        if f.endswith(".h"):
            continue                                     # the C-ABI declarations do not enter the kernels
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    h.update(" ".join(ARCH + FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp = LIB + ".digest"
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    os.makedirs(OBJ, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        tu_stamp, tu_dig = obj + ".digest", _tu_digest(src)
        if not force and os.path.exists(obj) and os.path.exists(tu_stamp) and open(tu_stamp).read() == tu_dig:
            return obj                                   # this translation unit and its headers are unchanged
        cmd = [nvcc(), *ARCH, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(OBJ, src[:-3] + ".ptxas.log")
        # register / spill report of every kernel (kept in the tree as evidence); compile times would churn the file
        open(log, "w").write("".join(l for l in r.stderr.splitlines(True) if "Compile time" not in l))
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stderr[-4000:]}")
        open(tu_stamp, "w").write(tu_dig)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc(), *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    open(stamp, "w").write(dig)
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
