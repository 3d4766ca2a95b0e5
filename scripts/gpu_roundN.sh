cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
shift
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/r4_bench_n$N.json 2> gpurun_out/r4_bench_n$N.err
python -c "
import json;d=json.load(open('gpurun_out/r4_bench_n$N.json'));s=d.get('sharded') or {}
print(d['n_gpus'],d['ms_per_step'],d['value']/1e9,'e2e',d['e2e']['ms_per_step'],d['e2e']['value']/1e9,'sharded',s.get('ms_per_step'),s.get('value'),s.get('allreduce_ms'),s.get('parity'))"
