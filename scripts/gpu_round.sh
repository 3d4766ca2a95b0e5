cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "filtfilt or notch or causal or butter" > gpurun_out/t11.log 2>&1; tail -3 gpurun_out/t11.log
{
python scripts/prof_ops.py notch,bandpass 256 7200000 5
ECOG_SOS_TPS=768 python scripts/prof_ops.py notch,bandpass 256 7200000 5
} > gpurun_out/t11_prof.log 2>&1; cat gpurun_out/t11_prof.log
