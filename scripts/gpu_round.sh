cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu > gpurun_out/t5_tests.log 2>&1; tail -4 gpurun_out/t5_tests.log
python scripts/prof_e2e.py > gpurun_out/e2e_prof.log 2>&1; cat gpurun_out/e2e_prof.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2_c.json 2> gpurun_out/bench_c2_c.err; cat gpurun_out/bench_c2_c.json; tail -3 gpurun_out/bench_c2_c.err
