#!/bin/bash
# One GPU-box visit: parity tests, the bench line (both arms), the ncu launch list and full captures of the hot
# kernels.  Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag>
# Afterwards, here:
#   python tools/ncu_traffic.py fp32=gpurun_out/k1_<tag>.ncu-rep fp64=gpurun_out/k1f64_<tag>.ncu-rep pr_fp32=gpurun_out/pr_<tag>.ncu-rep
#   python tools/ncu_summary.py gpurun_out/<report> > profiles/...
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; head -c 600 gpurun_out/bench_$TAG.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; head -c 400 gpurun_out/bench_ref_$TAG.json; echo
LEGS="k1 pr big bigsym otf sheppard lib"
python tools/prof_cmd.py $LEGS > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python tools/prof_cmd.py $LEGS > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
python tools/prof_cmd.py k1 > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'hanser_sym_kernel' -s 1 -c 1 -f \
    -o gpurun_out/k1_$TAG python tools/prof_cmd.py k1 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "K1 capture rc=$?"
python tools/prof_cmd.py k1 --precision=fp64 > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'hanser_sym_kernel' -s 1 -c 1 -f \
    -o gpurun_out/k1f64_$TAG python tools/prof_cmd.py k1 --precision=fp64 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "K1 fp64 capture rc=$?"
# one iteration per launch for the capture (the shipped loop runs up to 8 behind in-kernel barriers: same kernel code),
# so that the per-launch counters are per-iteration counters
export PYOTF_B200_PR_ITERS_PER_LAUNCH=1
python tools/prof_cmd.py pr > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'pr_plane' -s 2 -c 1 -f \
    -o gpurun_out/pr_$TAG python tools/prof_cmd.py pr > gpurun_out/ncu_full_$TAG.log 2>&1
echo "K2 capture rc=$?"
unset PYOTF_B200_PR_ITERS_PER_LAUNCH
python tools/prof_cmd.py big bigsym otf sheppard > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none -k regex:'hanser_big|rfft_z|cfft_rows|fft_lines|sheppard_zplanes|hanser_psf_kernel' -c 24 -f \
    -o gpurun_out/other_$TAG python tools/prof_cmd.py big bigsym otf sheppard > gpurun_out/ncu_full_$TAG.log 2>&1
echo "other kernels capture rc=$?"
# gpurun brings back at most 64 MiB: condense the big reports here (ncu is on the box) and drop them
python tools/ncu_summary.py gpurun_out/other_$TAG.ncu-rep > gpurun_out/other_kernels_$TAG.md 2>/dev/null; rm -f gpurun_out/other_$TAG.ncu-rep
K2=$(cuobjdump -elf pyotf_b200/libpyotf_b200.so 2>/dev/null | grep -o "_ZN5pyotf15pr_plane_kernelIfLi256ELi64ELb1ELi512ELb1E[A-Za-z0-9_]*" | head -1)
python tools/ncu_by_line.py gpurun_out/pr_$TAG.ncu-rep "$K2" pr_f32 --top 60 > gpurun_out/k2_by_line_$TAG.txt 2>gpurun_out/k2_by_line_$TAG.err
ncu -i gpurun_out/pr_$TAG.ncu-rep --page raw --csv > gpurun_out/pr_raw_$TAG.csv 2>/dev/null; rm -f gpurun_out/pr_$TAG.ncu-rep
du -sh gpurun_out
