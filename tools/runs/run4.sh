cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
S1_TRACE=2 S1_REPLAY=1 timeout 600 python tools/s1_probe.py 128 4 > gpurun_out/r4_s1_seq.log 2>&1; grep "rep " gpurun_out/r4_s1_seq.log
S1_TRACE=2 S1_REPLAY=0 timeout 600 python tools/s1_probe.py 128 4 > gpurun_out/r4_s1_cc.log 2>&1; grep "rep " gpurun_out/r4_s1_cc.log
