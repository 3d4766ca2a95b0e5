cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t32.log 2>&1; tail -2 gpurun_out/t32.log
timeout 600 python bench.py > gpurun_out/bench32.json 2> gpurun_out/bench32.err; python -c "
import json;d=json.load(open('gpurun_out/bench32.json'));print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['step_ms'],d['gpu_launches'])"
