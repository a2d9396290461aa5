nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/pr_shard_time.py 2>&1 | grep "world"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench2_r2k.json 2> gpurun_out/bench2_r2k.err; echo "bench2 rc=$?"
