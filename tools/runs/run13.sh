cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r13_pytest_full.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r13_pytest_full.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r13_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r13_smoke.log
