cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${1:-2}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r11_bench_n$N.json 2> gpurun_out/r11_bench_n$N.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/r11_bench_n$N.json; tail -5 gpurun_out/r11_bench_n$N.err
