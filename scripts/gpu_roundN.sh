cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_c2_n$N.json 2> gpurun_out/bench_c2_n$N.err; python -c "
import json;d=json.load(open('gpurun_out/bench_c2_n$N.json'));print(d['n_gpus'],d['ms_per_step'],d['value']/1e9,d['e2e'])"
