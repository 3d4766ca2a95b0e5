cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
nvidia-smi topo -m 2>/dev/null | head -12
for aff in 0 1; do
if [ $aff = 0 ]; then export ECOG_NO_AFFINITY=1; else unset ECOG_NO_AFFINITY; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$aff bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_c2_n${N}_aff$aff.json 2> gpurun_out/bench_c2_n${N}_aff$aff.err; python -c "
import json;d=json.load(open('gpurun_out/bench_c2_n${N}_aff$aff.json'));print('aff=$aff',d['n_gpus'],d['ms_per_step'],d['value']/1e9,d['e2e']['ms_per_step'],d['e2e']['value']/1e9,d['e2e'].get('host_cpus'))"
done
