cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu -k "full_size" > gpurun_out/t16.log 2>&1; tail -12 gpurun_out/t16.log
