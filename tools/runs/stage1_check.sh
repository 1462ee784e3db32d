cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_instances.py tests/test_submap.py -x -q > gpurun_out/r9_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r9_pytest.log
S1_TRACE=1 timeout 600 python tools/s1_probe.py 128 6 > gpurun_out/r9_s1.log 2>&1; grep "rep \|\] [a-z]" gpurun_out/r9_s1.log | grep -v "task "
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_s1|k_dcvc|k_fill|Radix|radix" --csv --log-file gpurun_out/r9_s1_launches.csv python tools/s1_probe.py 128 2 > /dev/null 2>&1
