cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t54.log 2>&1; tail -2 gpurun_out/t54.log
timeout 60 scripts/micro/fp32_pipes > gpurun_out/fp32_pipes.txt 2>&1; cat gpurun_out/fp32_pipes.txt
