cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t47.log 2>&1; tail -2 gpurun_out/t47.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke47.log 2>&1; tail -1 gpurun_out/smoke47.log
timeout 600 python bench.py > gpurun_out/bench47.json 2> gpurun_out/bench47.err; python -c "
import json;d=json.load(open('gpurun_out/bench47.json'));print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['step_ms'],d['gpu_launches'],d['clocks'])"
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/bench47b.json 2>/dev/null && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/ncu47.log 2>&1
timeout 300 python scripts/prof_ops.py hilbert 64 7200000 2 > gpurun_out/ops47.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hilbert -s 2 -c 1 -o gpurun_out/prof_r01_hilbert8_e -f python scripts/prof_ops.py hilbert 64 7200000 2 > gpurun_out/ncu47b.log 2>&1
tail -1 gpurun_out/ncu47b.log
