cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu > gpurun_out/t7_tests.log 2>&1; tail -6 gpurun_out/t7_tests.log
python - <<'PY' > gpurun_out/t7_prof.log 2>&1
import torch, sys
sys.path.insert(0, '.')
from decode_tonal_langauge_b200 import ops
x = torch.randn((128, 1_200_000), device='cuda')*30
def t(f, n=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): y = f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
print("fir_bank 391 taps  128x1.2M: %.3f ms" % t(lambda: ops.fir_bank(x, 2000.0, 390, [80.,100.,120.])))
print("rolling  W=20000   128x1.2M: %.3f ms" % t(lambda: ops.rolling_zscore(x, 20000)))
PY
cat gpurun_out/t7_prof.log
