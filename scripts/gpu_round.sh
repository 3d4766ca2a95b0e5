cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t51.log 2>&1; tail -3 gpurun_out/t51.log
timeout 300 python scripts/prof_ops.py car,resample,resample,resample1 256 7200000 5 > gpurun_out/ops51.log 2>&1
cat gpurun_out/ops51.log
