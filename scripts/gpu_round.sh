cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke52.log 2>&1; tail -1 gpurun_out/smoke52.log
timeout 600 python bench.py > gpurun_out/bench52.json 2> gpurun_out/bench52.err; python -c "
import json;d=json.load(open('gpurun_out/bench52.json'));print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['step_ms'],d['gpu_launches'],d['clocks'])"
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/bench52b.json 2>/dev/null && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/ncu52.log 2>&1
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench52_ref.json 2> gpurun_out/bench52_ref.err; tail -c 700 gpurun_out/bench52_ref.json
