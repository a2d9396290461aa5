nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/multigpu_check.py 2>&1 | grep -E "rank 0\]|CHECK" | tail -16
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tools/pr_shard_time.py 2>&1 | grep "world"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench8_r2e.json 2> gpurun_out/bench8_r2e.err; echo "bench8 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 50 --warmup 5 --no-others > gpurun_out/bench4_r2e.json 2> gpurun_out/bench4_r2e.err; echo "bench4 rc=$?"
