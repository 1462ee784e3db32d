cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 | cut -c1-400
