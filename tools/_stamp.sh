# ncu captures for the traffic stamp only (K1s fp32 / fp64, the fused PR kernel); see tools/gpu_round.sh
set -u
TAG=${1:-r2s}
python tools/prof_cmd.py k1 > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'hanser_sym_kernel' -s 1 -c 1 -f \
    -o gpurun_out/k1_$TAG python tools/prof_cmd.py k1 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "K1 capture rc=$?"
python tools/prof_cmd.py k1 --precision=fp64 > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'hanser_sym_kernel' -s 1 -c 1 -f \
    -o gpurun_out/k1f64_$TAG python tools/prof_cmd.py k1 --precision=fp64 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "K1 fp64 capture rc=$?"
export PYOTF_B200_PR_ITERS_PER_LAUNCH=1
python tools/prof_cmd.py pr > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'pr_plane' -s 2 -c 1 -f \
    -o gpurun_out/pr_$TAG python tools/prof_cmd.py pr > gpurun_out/ncu_full_$TAG.log 2>&1
echo "K2 capture rc=$?"
ncu -i gpurun_out/pr_$TAG.ncu-rep --page raw --csv > gpurun_out/pr_raw_$TAG.csv 2>/dev/null; rm -f gpurun_out/pr_$TAG.ncu-rep
