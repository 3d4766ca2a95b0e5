cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu > gpurun_out/t18.log 2>&1; tail -4 gpurun_out/t18.log
