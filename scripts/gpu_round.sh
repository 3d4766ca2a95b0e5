cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "resample or fir or halfband" > gpurun_out/r6_t2.log 2>&1; tail -2 gpurun_out/r6_t2.log
timeout 300 python scripts/prof_ops.py fir,fir 256 7200000 10 > gpurun_out/r6_ops2.log 2>&1; cat gpurun_out/r6_ops2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sos_pair_ws --launch-skip 4 -c 2 -o gpurun_out/prof_r6_pair -f python scripts/prof_ops.py pair 256 7200000 1 > gpurun_out/r6_ncu_pair.log 2>&1; tail -1 gpurun_out/r6_ncu_pair.log
