cd $GRAFT_REPO_ROOT
python scripts/prof_ops.py notch,bandpass,notch 256 7200000 5 > gpurun_out/t15_prof.log 2>&1; cat gpurun_out/t15_prof.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2_d.json 2> gpurun_out/bench_c2_d.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_d.json')); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['step_ms'], d['clocks'])"
