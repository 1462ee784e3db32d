cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | grep "^{" | cut -c1-330
