# one GPU box call of the development loop: parity tests, smoke, the default bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/final_t.log 2>&1; tail -3 gpurun_out/final_t.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-c5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; python -c "
import json;d=json.load(open('gpurun_out/final_bench.json'));print(d['ms_per_step'],d['value']/1e9,d['step_ms'],d['parity']['max_rel'],d['e2e']['ms_per_step'],d['roofline']['traffic'],d['roofline']['steps']['frequency_filter[butter_bandstop]+frequency_filter[butter_bandpass]'].get('traffic_gb'))"
