cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_instances.py tests/test_submap.py -x -q > gpurun_out/r22_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r22_pytest.log
( timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_instances.py -x -q -k "components and stage1_golden" ) > gpurun_out/r21_racecheck_s1.log 2>&1; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/r21_racecheck_s1.log | tail -2
grep -E "Write access|Read access" gpurun_out/r21_racecheck_s1.log | sed -E 's/\+0x[0-9a-f]+//; s/=========//; s/\[[0-9]+ hazards\]//' | awk '{$1=$1};1' | sort | uniq -c | sort -rn | head -8
python tools/s1_probe.py 128 5 2>&1 | grep "rep "
