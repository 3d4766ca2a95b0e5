cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_c2_n8.json 2> gpurun_out/bench_c2_n8.err; tail -c 600 gpurun_out/bench_c2_n8.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 --workload C4 > gpurun_out/bench_c4_n8.json 2> gpurun_out/bench_c4_n8.err; tail -c 1500 gpurun_out/bench_c4_n8.json
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/t_multi8.log 2>&1; tail -3 gpurun_out/t_multi8.log
