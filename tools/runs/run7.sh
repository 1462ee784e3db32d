cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity_full.py > gpurun_out/r7_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r7_pytest.log
S1_TRACE=1 timeout 600 python tools/s1_probe.py 128 6 > gpurun_out/r7_s1.log 2>&1; grep "rep \|\] [a-z]" gpurun_out/r7_s1.log | grep -v "task "
timeout 900 python tools/join_probe.py 100000 1024 "" "collect_unroll=1" "" "collect_unroll=1" > gpurun_out/r7_join_probe.log 2>&1; cat gpurun_out/r7_join_probe.log
