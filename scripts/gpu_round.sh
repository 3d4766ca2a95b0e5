cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python scripts/prof_ops.py car,notch,bandpass,zscore 256 7200000 1 > gpurun_out/ops49.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sos_warm|car_fused|zscore_apply|row_stats_partial" -s 12 -c 7 -o gpurun_out/prof_r01_streaming -f python scripts/prof_ops.py car,notch,bandpass,zscore 256 7200000 1 > gpurun_out/ncu49.log 2>&1
cat gpurun_out/ops49.log; tail -1 gpurun_out/ncu49.log
