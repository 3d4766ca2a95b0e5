cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "hilbert" > gpurun_out/t14.log 2>&1; tail -3 gpurun_out/t14.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 2 --warmup 3 --workload C4small > gpurun_out/b2.out 2> gpurun_out/b2.err; wc -l gpurun_out/b2.out; head -c 200 gpurun_out/b2.out
