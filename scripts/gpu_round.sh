cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "hilbert or chain or full6 or smoke or pipeline" > gpurun_out/t50.log 2>&1; tail -3 gpurun_out/t50.log
timeout 300 python scripts/prof_ops.py car,hilbert,hilbert 256 7200000 5 > gpurun_out/ops50.log 2>&1
cat gpurun_out/ops50.log
