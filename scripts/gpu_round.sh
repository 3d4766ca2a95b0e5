cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python scripts/prof_ops.py pair,pair,pair 256 7200000 10 > gpurun_out/r5_ops1.log 2>&1
ECOG_PROF_FS=3000 timeout 300 python scripts/prof_ops.py pair,pair 128 10800000 10 >> gpurun_out/r5_ops1.log 2>&1
cat gpurun_out/r5_ops1.log
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -s -k "float32_bandpass" > gpurun_out/r5_t1.log 2>&1; grep -n "pair @\|passed\|failed\|Error" gpurun_out/r5_t1.log
