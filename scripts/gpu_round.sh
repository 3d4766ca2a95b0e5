cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in 0 1; do
if [ $v = 1 ]; then export ECOG_FFT_NARROW=1; fi
timeout 600 python bench.py --steps 8 --warmup 3 --no-c5 --no-e2e --no-cpu > gpurun_out/r5_bench2_$v.json 2> gpurun_out/r5_bench2_$v.err; python -c "
import json;d=json.load(open('gpurun_out/r5_bench2_$v.json'));print($v, d['ms_per_step'],d['step_ms']['downsample'])"
done
ECOG_FFT_NARROW=1 timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "resample or fft" > gpurun_out/r5_t4.log 2>&1; tail -2 gpurun_out/r5_t4.log
