cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r16_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r16_pytest.log
timeout 900 python tools/join_probe.py 100000 1024 "" "" > gpurun_out/r16_probe.log 2>&1; cat gpurun_out/r16_probe.log
