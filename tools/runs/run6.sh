cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
S1_TRACE=2 S1_REPLAY=0 timeout 600 python tools/s1_probe.py 128 6 > gpurun_out/r6_s1_cc.log 2>&1; grep "rep \|\] [a-z]" gpurun_out/r6_s1_cc.log | grep -v "task "
