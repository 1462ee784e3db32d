cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_instances.py tests/test_submap.py -x -q > gpurun_out/r5_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r5_pytest.log
S1_TRACE=2 S1_REPLAY=0 timeout 600 python tools/s1_probe.py 128 6 > gpurun_out/r5_s1_cc.log 2>&1; grep "rep \|\] [a-z]" gpurun_out/r5_s1_cc.log | grep -v "task "
