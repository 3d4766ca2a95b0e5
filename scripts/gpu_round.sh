cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r4_t2.log 2>&1; tail -3 gpurun_out/r4_t2.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4_smoke.log 2>&1; tail -3 gpurun_out/r4_smoke.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'hilbert_env8|halfband2' --launch-skip 4 -c 2 -o gpurun_out/prof_r4_hb_hil -f python scripts/prof_ops.py fir,hilbert_car 256 7200000 1 > gpurun_out/r4_ncu_full2.log 2>&1; tail -2 gpurun_out/r4_ncu_full2.log
