set -u
TAG=r2f
SECONDS=0; python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$? wall ${SECONDS}s"
SECONDS=0; python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref wall ${SECONDS}s"
LEGS="k1 pr big bigsym otf sheppard lib"
python tools/prof_cmd.py $LEGS > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python tools/prof_cmd.py $LEGS > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
