cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_s1|k_dcvc|k_fill|Radix|radix" --csv --log-file gpurun_out/r8_s1_launches.csv python tools/s1_probe.py 128 2 > gpurun_out/r8_s1.log 2>&1
tail -3 gpurun_out/r8_s1.log
