cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r7_bench_c2.json 2> gpurun_out/r7_bench_c2.err; python -c "
import json;d=json.load(open('gpurun_out/r7_bench_c2.json'));print(d['ms_per_step'],d['value']/1e9,d['step_ms'],d['parity']['max_rel'],d['e2e']['ms_per_step'],d['sessions_c5']['s_total'],d['roofline']['traffic'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r7_launches_c2.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu --no-c5 > gpurun_out/r7_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:'halfband2' --launch-skip 2 -c 1 -o gpurun_out/prof_r7_hb -f python scripts/prof_ops.py fir 256 7200000 1 > gpurun_out/r7_ncu_hb.log 2>&1; tail -1 gpurun_out/r7_ncu_hb.log
