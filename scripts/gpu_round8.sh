cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_c2_n8.json 2> gpurun_out/bench_c2_n8.err; python -c "
import json;d=json.load(open('gpurun_out/bench_c2_n8.json'));print(d['n_gpus'],d['ms_per_step'],d['value']/1e9,d['e2e']['ms_per_step'],d['e2e']['value']/1e9,d['e2e'].get('host_cpus'))"
