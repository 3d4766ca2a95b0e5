cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "resample or c2c or fir" > gpurun_out/t12.log 2>&1; tail -15 gpurun_out/t12.log
python - <<'PY' > gpurun_out/t12_prof.log 2>&1
import torch, sys, time
sys.path.insert(0, '.')
from decode_tonal_langauge_b200 import ops
x = torch.randn((64, 1_831_054), device='cuda')*30
def t(f, n=2):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): y = f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
print("czt two-stage 64 x 1831054 -> 239999: %.2f ms" % t(lambda: ops.fft_resample(x, 239_999)))
print("czt one-stage 64 x 1831054 -> 239999: %.2f ms" % t(lambda: ops.fft_resample(x, 239_999, two_stage=False)))
PY
cat gpurun_out/t12_prof.log
