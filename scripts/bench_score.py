"""time the score matmul kernel alone (CUDA events, L2 flushed between launches): python scripts/bench_score.py [M]
   DPC_SCORE_NO_TMA_STORE=1 selects the direct-store epilogue"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dpc_b200 import engine as E
from dpc_b200._lib import lib, ptr

M = int(sys.argv[1]) if len(sys.argv) > 1 else 6144
D = 256
L = lib()
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device='cuda').manual_seed(0)
pred = torch.randn(M, D, device='cuda', generator=g) * 0.25
finf = torch.randn(M, D, device='cuda', generator=g)
pp, fp = E._split(pred, st, f16=True), E._split(finf, st, f16=True)
score = torch.empty(M, M, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
ts = []
for i in range(25):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    L.score_matmul_tc(M, M, D, ptr(pp[0]), ptr(pp[1]), ptr(fp[0]), ptr(fp[1]), 1, ptr(score), st)
    b.record()
    torch.cuda.synchronize()
    if i >= 5:
        ts.append(a.elapsed_time(b))
ts.sort()
ref = pred.double() @ finf.double().t()
err = float((score.double() - ref).abs().max() / ref.abs().max())
med = ts[len(ts) // 2]
print('score %dx%d K=%d: median %.1f us (min %.1f)  %.0f GB/s of output, %.0f TFLOP/s algorithmic, rel err %.2e, tma_store=%s'
      % (M, M, D, med * 1e3, ts[0] * 1e3, M * M * 4 / med / 1e6, 2.0 * M * M * D / med / 1e9, err,
         'off' if os.environ.get('DPC_SCORE_NO_TMA_STORE') else 'on'))
