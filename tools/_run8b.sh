nvidia-smi -L | wc -l
for N in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench${N}_r2h.json 2> gpurun_out/bench${N}_r2h.err; echo "bench$N rc=$?"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/bench8_ref_r2h.json 2>/dev/null; echo "ref8 rc=$?"; head -c 300 gpurun_out/bench8_ref_r2h.json
