cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu > gpurun_out/t3_tests.log 2>&1; tail -5 gpurun_out/t3_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2_b.json 2> gpurun_out/bench_c2_b.err; cat gpurun_out/bench_c2_b.json; tail -3 gpurun_out/bench_c2_b.err
