cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python bench.py --workload seq --scans 4541 --queries 256 --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/r02b_seq_4541.json 2> gpurun_out/r02b_seq_4541.err; echo "seq rc=$?"; tail -c 1600 gpurun_out/r02b_seq_4541.json; tail -2 gpurun_out/r02b_seq_4541.err
