set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q > gpurun_out/r1_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r1_pytest.log
tools/ubench/fma_probe > gpurun_out/r1_fma.log 2>&1; cat gpurun_out/r1_fma.log
timeout 900 python tools/join_probe.py 100000 1024 > gpurun_out/r1_join_probe.log 2>&1; echo "probe rc=$?"
cat gpurun_out/r1_join_probe.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_vote_join|k_verify|k_collect_inv|k_probe_emit|k_query_index|k_topk" --launch-skip 6 --launch-count 6 -o gpurun_out/r1_search python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r1_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r1_ncu.log
ls -la gpurun_out
