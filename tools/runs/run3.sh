set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_instances.py tests/test_submap.py -x -q > gpurun_out/r3_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r3_pytest.log
S1_TRACE=1 S1_REPLAY=1 timeout 600 python tools/s1_probe.py 128 6 > gpurun_out/r3_s1_seq.log 2>&1; tail -22 gpurun_out/r3_s1_seq.log
S1_TRACE=1 S1_REPLAY=0 timeout 600 python tools/s1_probe.py 128 6 > gpurun_out/r3_s1_cc.log 2>&1; tail -22 gpurun_out/r3_s1_cc.log
timeout 900 python tools/join_probe.py 100000 1024 "" "collect_unroll=1" "collect_unroll=4" "" > gpurun_out/r3_join_probe.log 2>&1; echo "probe rc=$?"
cat gpurun_out/r3_join_probe.log
