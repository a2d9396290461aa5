#!/bin/bash
# One GPU-box visit: parity tests, the bench line (both arms), the ncu launch list and full captures of the hot
# kernels.  Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag>
# Afterwards, here:  python tools/ncu_traffic.py fp32=gpurun_out/k1_<tag>.ncu-rep fp64=gpurun_out/k1f64_<tag>.ncu-rep
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; head -c 1500 gpurun_out/bench_$TAG.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_ref_$TAG.json
python tools/prof_cmd.py k1 pr > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python tools/prof_cmd.py k1 pr > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
python tools/prof_cmd.py k1 > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'hanser_sym_kernel' -s 1 -c 1 -f \
    -o gpurun_out/k1_$TAG python tools/prof_cmd.py k1 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "K1 capture rc=$?"
python tools/prof_cmd.py k1 --precision=fp64 > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'hanser_sym_kernel' -s 1 -c 1 -f \
    -o gpurun_out/k1f64_$TAG python tools/prof_cmd.py k1 --precision=fp64 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "K1 fp64 capture rc=$?"
python tools/prof_cmd.py pr > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'pr_' -s 2 -c 3 -f \
    -o gpurun_out/pr_$TAG python tools/prof_cmd.py pr > gpurun_out/ncu_full_$TAG.log 2>&1
echo "K2 capture rc=$?"
