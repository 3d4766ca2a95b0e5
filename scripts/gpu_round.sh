cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "hilbert_golden or hilbert_shapes or hilbert_pruned or resample or butter_filtfilt_golden or ragged" > gpurun_out/memcheck48.log 2>&1; echo rc=$?; tail -6 gpurun_out/memcheck48.log; grep -c "Invalid\|ERROR SUMMARY" gpurun_out/memcheck48.log
