cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t43.log 2>&1; tail -2 gpurun_out/t43.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke43.log 2>&1; tail -2 gpurun_out/smoke43.log
timeout 600 python bench.py > gpurun_out/bench43.json 2> gpurun_out/bench43.err; python -c "
import json;d=json.load(open('gpurun_out/bench43.json'));print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['step_ms'],d['gpu_launches'],d['clocks'])"
timeout 600 python bench.py --workload C3 > gpurun_out/bench43_c3.json 2> gpurun_out/bench43_c3.err; tail -c 1200 gpurun_out/bench43_c3.json
