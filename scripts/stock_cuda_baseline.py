"""DEV TOOL (test infrastructure, not product): throughput of the plain PyTorch-CUDA path (the oracle's
functional restatement = the reference's ATen op sequence: cuDNN conv/BN, cuBLAS mm) on the same synthetic
workload, torch defaults as in dpc/main.py:25 (cudnn.benchmark, TF32 convs).  Usage: python scripts/stock_cuda_baseline.py [B]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import dpc_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
torch.backends.cudnn.benchmark = True
net = 'resnet18'
sd = {k: v.cuda() for k, v in O.synthetic_state_dict(net, 0).items()}
params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if not k.startswith('agg.ConvGRUCell_00')}
full = dict(params)
for k in sd:
    if k.startswith('agg.ConvGRUCell_00'):
        full[k] = params[k.replace('agg.ConvGRUCell_00', 'agg.cell_list.0')]
opt = torch.optim.Adam(list(params.values()), lr=1e-3, weight_decay=1e-5)
x = torch.randn(B, 8, 3, 5, 128, 128, device='cuda')
for fmt in ('channels_first', 'channels_last_3d'):
    def step():
        opt.zero_grad(set_to_none=True)
        xx = x
        score, mask = O.dpc_forward(xx, full, net, 3)
        M = score.shape[0] * score.shape[1] * score.shape[2]
        loss = torch.nn.functional.cross_entropy(score.view(M, M), torch.arange(M, device='cuda'))
        loss.backward()
        opt.step()
        return loss
    if fmt == 'channels_last_3d':
        break     # the reference uses the default (contiguous NCDHW) layout only
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 5
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print('stock PyTorch-CUDA (TF32 convs, cudnn.benchmark): B=%d  %.1f ms/step  %.1f clips/s  (peak mem %.1f GB)'
          % (B, ms, B / ms * 1e3, torch.cuda.max_memory_allocated() / 1e9))
