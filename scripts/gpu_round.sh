cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r3_t7.log 2>&1; tail -3 gpurun_out/r3_t7.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-c5 > gpurun_out/r3_bench1.json 2> gpurun_out/r3_bench1.err; python -c "
import json;d=json.load(open('gpurun_out/r3_bench1.json'));print(d['ms_per_step'],d['value']/1e9,d['step_ms'],d['parity'],d['e2e']['ms_per_step'])"
