cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r26_pytest_full.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r26_pytest_full.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
