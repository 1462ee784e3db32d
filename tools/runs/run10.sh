cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r10_bench.json 2> gpurun_out/r10_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r10_bench.json
timeout 1200 python bench.py --workload seq --steps 4 --warmup 2 > gpurun_out/r10_seq.json 2> gpurun_out/r10_seq.err; echo "seq rc=$?"; tail -c 1500 gpurun_out/r10_seq.json
