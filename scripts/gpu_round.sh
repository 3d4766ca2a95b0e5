cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for nt in 128 256 128 256; do ECOG_SOS_NT=$nt timeout 120 python scripts/prof_ops.py car,bandpass,notch,bandpass,notch 256 7200000 5 | sed "s/^/nt=$nt /"; done > gpurun_out/ops45.log 2>&1
grep -v car gpurun_out/ops45.log
