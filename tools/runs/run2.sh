set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest.log
timeout 900 python tools/join_probe.py 100000 1024 "" "collect_mode=2" "" > gpurun_out/r2_join_probe.log 2>&1; echo "probe rc=$?"
cat gpurun_out/r2_join_probe.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_collect_inv|k_query_index|k_vote_join" --launch-skip 3 --launch-count 3 -o gpurun_out/r2_search python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r2_ncu.log | cut -c1-300
