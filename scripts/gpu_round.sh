cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "resample or fir or halfband" > gpurun_out/r5_t3.log 2>&1; tail -2 gpurun_out/r5_t3.log
timeout 300 python scripts/prof_ops.py fir,fir 256 7200000 10 > gpurun_out/r5_ops3.log 2>&1; cat gpurun_out/r5_ops3.log
