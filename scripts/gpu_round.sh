cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for ch in 8 12 16 6; do
ECOG_PIPELINE_CHUNKS=$ch timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench53_$ch.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench53_$ch.json'));print('chunks=$ch',d['ms_per_step'],d['e2e']['ms_per_step'])"
done
