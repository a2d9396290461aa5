#!/bin/bash
# Debug library with phase timestamps (tools/_dbg/libpyotf_b200_dbg.so): recompiles the K1 translation units with
# -DPYOTF_K1_TIMING and links them with the release objects of csrc/_obj (run `python -m pyotf_b200.build` first).
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_dbg
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -DPYOTF_K1_TIMING -DPYOTF_PR_TIMING"
for src in ${@:-hanser_sym_f32}; do
  nvcc $F -c pyotf_b200/csrc/$src.cu -o tools/_dbg/$src.o &
done
wait
objs=""
for o in pyotf_b200/csrc/_obj/*.o; do
  b=$(basename $o)
  if [ -f tools/_dbg/$b ]; then objs="$objs tools/_dbg/$b"; else objs="$objs $o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/_dbg/libpyotf_b200_dbg.so $objs -cudart static
echo built tools/_dbg/libpyotf_b200_dbg.so
