"""the score matmul + NCE criterion alone at BASELINE config 2 / 5 sizes (M = B*P*SQ rows, D = 256), for ncu captures:
   ncu ... python scripts/profile_score.py [M]      launches: split x2, score_gemm_kernel (score fwd), ce_fwd, ce_bwd,
   split (dscore), conv_tc_kernel (dpred), wgrad_tc_kernel (dfinf)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dpc_b200 import engine as E
from dpc_b200._lib import lib, ptr, ConvGeom

M = int(sys.argv[1]) if len(sys.argv) > 1 else 6144
D = 256
L = lib()
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device='cuda').manual_seed(0)
pred = torch.randn(M, D, device='cuda', generator=g) * 0.25
finf = torch.randn(M, D, device='cuda', generator=g)
for _ in range(2):
    pp, fp = E._split(pred, st, f16=True), E._split(finf, st, f16=True)
    score = torch.empty(M, M, device='cuda')
    L.score_matmul_tc(M, M, D, ptr(pp[0]), ptr(pp[1]), ptr(fp[0]), ptr(fp[1]), 1, ptr(score), st)
    out, lse = E.nce_ce_forward(score)
    d = E.nce_ce_backward(score, lse, torch.ones(1, device='cuda'))
    dsp = E._split(d, st)
    ftp = E._split(finf.t().contiguous(), st)
    dpred = torch.empty(M, D, device='cuda')
    L.gemm_nt_split_tc(M, D, M, ptr(dsp[0]), ptr(dsp[1]), ptr(ftp[0]), ptr(ftp[1]), 0, ptr(dpred), 0, st)
    ppb = E._split(pred, st)
    geom = ConvGeom(1, 1, 1, M, D, 1, 1, M, M, 1, 1, 1, 1, 1, 1, 0, 0, 0)
    scratch, dfinf = torch.empty(M, D, device='cuda'), torch.empty(M, D, device='cuda')
    L.conv3d_wgrad_tc(geom, ptr(ppb[0]), ptr(ppb[1]), ptr(dsp[0]), ptr(dsp[1]), ptr(scratch), ptr(dfinf), st)
torch.cuda.synchronize()
print('loss', float(out[0]))
