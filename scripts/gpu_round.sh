cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu -k "anova or selection or runlength or yaml" > gpurun_out/t9.log 2>&1; tail -3 gpurun_out/t9.log
python bench.py --workload C3 --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c3.json')); print(d['ms_per_step'], d['step_ms'])"; tail -5 gpurun_out/bench_c3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload C3 --steps 1 --warmup 3 > gpurun_out/ncu_c3.log 2>&1
grep anova gpurun_out/launches_c3.csv | tail -6 | awk -F'","' '{print $5, $9, $NF}' | cut -c1-150
